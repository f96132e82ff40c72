// FP64 DFMA saturation probe: the fp64 roofline denominator for the nf = 8 path.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma(double* out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = 1.1, a2 = 1.2, a3 = 1.3, a4 = 1.4, a5 = 1.5, a6 = 1.6, a7 = 1.7;
    const double b = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200000;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); dfma<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
        printf("fp64 DFMA: %.2f ms  %.2f TFLOP/s\n", ms, flops / ms / 1e9);
    }
    return 0;
}

// Scratch probe (not part of the product): what does the K2 store PATTERN cost without the math?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pattern_probe pattern_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void linear_fill(double2* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) __stcs(out + i, make_double2(1.0, 2.0));
}

// thread <-> k, marches over T rows, 5 var rows each (the K2 pattern); `work` dependent DFMAs per row
template <int SYNC>
__global__ void k2_pattern(double2* out, long nk, long T, int work) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    double x = 1.0 + k * 1e-9, v = 0.5;
    double2* p = out + k;
    for (long n = 0; n < T; ++n) {
        for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
#pragma unroll
        for (int var = 0; var < 5; ++var) __stcs(p + var * nk, make_double2(x, v));
        p += 5 * nk;
        if (SYNC && (n & 63) == 63) __syncthreads();
    }
}

int main(int argc, char** argv) {
    long nk = 65536, T = 7854;
    size_t n = (size_t)nk * T * 5;
    double2* out;
    cudaMalloc(&out, n * sizeof(double2));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn) {
        float best = 1e9;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        printf("%-40s %.3f ms  %.1f GB/s  %s\n", name, best, n * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    timeit("linear fill 148*8 x 512", [&] { linear_fill<<<148 * 8, 512>>>(out, n); });
    int blocks[] = {64, 128, 256, 448, 512};
    for (int b : blocks) for (int work : {0, 8, 24}) {
        char name[96];
        snprintf(name, 96, "k2 pattern block=%d work=%d nosync", b, work);
        timeit(name, [&] { k2_pattern<0><<<(nk + b - 1) / b, b>>>(out, nk, T, work); });
        snprintf(name, 96, "k2 pattern block=%d work=%d sync64", b, work);
        timeit(name, [&] { k2_pattern<1><<<(nk + b - 1) / b, b>>>(out, nk, T, work); });
    }
    return 0;
}

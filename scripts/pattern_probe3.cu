// Scratch probe 3: ways of getting the perturbation rows into a pinned host yresult.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void zero_copy_rows(double2* host_out, long nk, long T, int work) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    double x = 1.0 + k * 1e-9, v = 0.5;
    double2* p = host_out + 3 * nk + k;
    for (long n = 0; n < T; ++n) {
        for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
        p[0] = make_double2(x, v);
        p[nk] = make_double2(v, x);
        p += 5 * nk;
    }
}

int main() {
    long nk = 65536, T = 1500;
    size_t full = (size_t)nk * T * 5 * 16, pert = (size_t)nk * T * 2 * 16;
    double2 *host, *dev;
    if (cudaHostAlloc(&host, full, cudaHostAllocDefault) != cudaSuccess) { printf("hostalloc failed\n"); return 1; }
    cudaMalloc(&dev, pert);
    cudaMemset(dev, 1, pert);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms, size_t bytes) { printf("%-44s %.2f ms  %.1f GB/s  %s\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError())); };
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); cudaMemcpyAsync(host, dev, pert, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("linear memcpy D2H (one call)", ms, pert);
        cudaEventRecord(e0);
        cudaMemcpy2DAsync((char*)host + 3 * nk * 16, 5 * nk * 16, dev, 2 * nk * 16, 2 * nk * 16, T, cudaMemcpyDeviceToHost);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("cudaMemcpy2DAsync pitch 5 rows, width 2 rows", ms, pert);
        cudaEventRecord(e0);
        for (long t = 0; t < T; ++t)
            cudaMemcpyAsync((char*)host + (t * 5 + 3) * nk * 16, (char*)dev + t * 2 * nk * 16, 2 * nk * 16, cudaMemcpyDeviceToHost);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("per-time-step memcpyAsync (2 MB each)", ms, pert);
        for (int b : {128, 448}) for (int work : {0, 8}) {
            cudaEventRecord(e0); zero_copy_rows<<<(nk + b - 1) / b, b>>>(host, nk, T, work); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            char name[96]; snprintf(name, 96, "kernel stores to mapped host, block=%d work=%d", b, work);
            report(name, ms, pert);
        }
    }
    return 0;
}

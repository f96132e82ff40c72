// Scratch probe 2: K2 store pattern vs CTA segment size / alignment / modes-per-thread.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

// CTA b owns modes [b*seg, (b+1)*seg); thread t handles modes t, t+blockDim, ... (KPT = seg/blockDim)
__global__ void k2_pattern(double2* out, long nk, long T, int seg, int work) {
    const long base = (long)blockIdx.x * seg;
    const int kpt = seg / blockDim.x;
    double x = 1.0 + threadIdx.x * 1e-9, v = 0.5;
    for (long n = 0; n < T; ++n) {
        for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
        double2* p = out + n * 5 * nk + base + threadIdx.x;
        for (int m = 0; m < kpt; ++m) {
            long k = base + threadIdx.x + (long)m * blockDim.x;
            if (k < nk) {
#pragma unroll
                for (int var = 0; var < 5; ++var) __stcs(p + var * nk + (long)m * blockDim.x, make_double2(x, v));
            }
        }
    }
}

int main(int argc, char** argv) {
    long nk = argc > 1 ? atol(argv[1]) : 65536, T = argc > 2 ? atol(argv[2]) : 7854;
    size_t n = (size_t)nk * T * 5;
    double2* out;
    if (cudaMalloc(&out, n * sizeof(double2)) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int seg, int threads, int work) {
        float best = 1e9;
        int grid = (int)((nk + seg - 1) / seg);
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0); k2_pattern<<<grid, threads>>>(out, nk, T, seg, work); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        printf("seg=%5d threads=%4d grid=%5d work=%2d  %.3f ms  %.1f GB/s %s\n", seg, threads, grid, work, best, n * 16.0 / best / 1e6,
               cudaGetErrorString(cudaGetLastError()));
    };
    int segs[] = {64, 128, 192, 256, 320, 384, 448, 512, 576, 640, 768, 896, 1024};
    for (int s : segs) run(s, s, 8);
    printf("-- fewer threads than modes per CTA\n");
    run(512, 256, 8); run(512, 128, 8); run(1024, 512, 8); run(1024, 256, 8); run(2048, 512, 8); run(4096, 512, 8);
    run(448, 224, 8); run(896, 448, 8);
    return 0;
}

// DMMA (mma.sync.m8n8k4.f64) vs DFMA on B200: throughput, dependent latency, and whether the two
// overlap.  Decides the design of the nf >= 3 mode kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void mix(double* out, int iters) {
    double acc[NACC > 0 ? NACC : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
    const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999 - threadIdx.x * 1e-9, c = 1e-9;
#pragma unroll
    for (int i = 0; i < (NACC > 0 ? NACC : 1); ++i) { acc[i][0] = i * 1e-3; acc[i][1] = i * 2e-3; }
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) f[i] = 1.0 + i * 0.1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = fma(f[i], b, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NACC > 0 ? NACC : 1); ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int NFMA>
static void run(const char* name, int ctas_per_sm, int threads, double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        mix<NACC, NFMA><<<148 * ctas_per_sm, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double nthreads = 148.0 * ctas_per_sm * threads;
    const double mma_flops = (double)NACC * iters * (nthreads / 32) * 512.0;
    const double fma_flops = (double)NFMA * iters * nthreads * 2.0;
    const double warps_per_smsp = ctas_per_sm * threads / 32.0 / 4.0;
    // cycles per warp-iteration per SMSP at 1.965 GHz
    const double cyc = best * 1e-3 * 1.965e9 / iters / warps_per_smsp;
    printf("%-28s warps/SMSP %4.1f  %.3f ms  DMMA %.2f TF/s  DFMA %.2f TF/s  total %.2f TF/s  (%.1f clk per warp-iter per SMSP: %d dmma + %d dfma)\n",
           name, warps_per_smsp, best, mma_flops / best / 1e9, fma_flops / best / 1e9,
           (mma_flops + fma_flops) / best / 1e9, cyc, NACC, NFMA);
}

int main() {
    double* out; cudaMalloc(&out, 148 * 16 * 1024 * sizeof(double));
    cudaError_t e;
    // throughput
    run<0, 8>("dfma ilp8", 8, 256, out);
    run<0, 8>("dfma ilp8 4w/smsp", 1, 512, out);
    run<0, 8>("dfma ilp8 3w/smsp", 3, 128, out);
    run<0, 4>("dfma ilp4 3w/smsp", 3, 128, out);
    run<0, 2>("dfma ilp2 3w/smsp", 3, 128, out);
    run<8, 0>("dmma ilp8", 8, 256, out);
    run<4, 0>("dmma ilp4", 8, 256, out);
    run<8, 0>("dmma ilp8 4w/smsp", 1, 512, out);
    run<4, 0>("dmma ilp4 4w/smsp", 1, 512, out);
    run<2, 0>("dmma ilp2 4w/smsp", 1, 512, out);
    run<4, 0>("dmma ilp4 2w/smsp", 1, 256, out);
    run<4, 0>("dmma ilp4 1w/smsp", 1, 128, out);
    // latency: one warp per SMSP, one chain
    run<1, 0>("dmma chain 1w/smsp", 1, 128, out);
    run<0, 1>("dfma chain 1w/smsp", 1, 128, out);
    run<2, 0>("dmma 2 chains 1w/smsp", 1, 128, out);
    run<0, 2>("dfma 2 chains 1w/smsp", 1, 128, out);
    // overlap: do DMMA and DFMA share a pipe?
    run<4, 32>("mix 4 dmma + 32 dfma", 8, 256, out);
    run<4, 16>("mix 4 dmma + 16 dfma", 8, 256, out);
    run<2, 14>("mix 2 dmma + 14 dfma", 8, 256, out);
    run<8, 8>("mix 8 dmma + 8 dfma", 8, 256, out);
    e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

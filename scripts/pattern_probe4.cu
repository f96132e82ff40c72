// Scratch probe 4: K2 output via TMA bulk stores from shared memory vs per-thread stores.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// direct: thread <-> k, 5 x STG.128 per time step
__global__ void direct(double2* out, long nk, long T, int work) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nk) return;
    double x = 1.0 + k * 1e-9, v = 0.5;
    double2* p = out + k;
    for (long n = 0; n < T; ++n) {
        for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
#pragma unroll
        for (int var = 0; var < 5; ++var) __stcs(p + var * nk, make_double2(x, v));
        p += 5 * nk;
    }
}

// staged: threads write their 5 values of RT time steps into smem [RT][5][B]; one thread then
// issues RT*5 bulk stores of B*16 bytes (contiguous segment of one row each); double buffered.
template <int RT>
__global__ void staged(double2* out, long nk, long T, int work) {
    extern __shared__ __align__(128) double2 sm[];      // [2][RT][5][B]
    const int B = blockDim.x;
    long k0 = (long)blockIdx.x * B;
    double x = 1.0 + (k0 + threadIdx.x) * 1e-9, v = 0.5;
    int buf = 0;
    for (long n0 = 0; n0 < T; n0 += RT) {
        double2* s = sm + (size_t)buf * RT * 5 * B;
        // make sure the bulk stores that read this buffer two iterations ago are done
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        for (int r = 0; r < RT; ++r) {
            for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
#pragma unroll
            for (int var = 0; var < 5; ++var) s[(r * 5 + var) * B + threadIdx.x] = make_double2(x, v);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int r = 0; r < RT && n0 + r < T; ++r)
                for (int var = 0; var < 5; ++var) {
                    double2* g = out + ((n0 + r) * 5 + var) * nk + k0;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                                 "r"(smem_u32(s + (r * 5 + var) * B)), "r"((uint32_t)(B * 16))
                                 : "memory");
                }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
    long nk = 65536, T = 7854;
    size_t n = (size_t)nk * T * 5;
    double2* out; cudaMalloc(&out, n * sizeof(double2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char* name, auto fn) {
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            cudaEventRecord(e0); fn(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        printf("%-44s %.3f ms  %.1f GB/s  %s\n", name, best, n * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int b : {128, 256, 448, 512}) {
        if (nk % b) { /* 448: 65536 = 146*448 + 128 -> use padded nk below */ }
        char name[96];
        long grid = nk / b;                 // exact tiles only (448 leaves a remainder: skipped modes)
        snprintf(name, 96, "direct block=%d", b);
        timeit(name, [&] { direct<<<grid, b>>>(out, nk, T, 8); });
        snprintf(name, 96, "staged RT=1 block=%d", b);
        cudaFuncSetAttribute(staged<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        timeit(name, [&] { staged<1><<<grid, b, 2 * 1 * 5 * b * 16>>>(out, nk, T, 8); });
        snprintf(name, 96, "staged RT=4 block=%d", b);
        cudaFuncSetAttribute(staged<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        timeit(name, [&] { staged<4><<<grid, b, 2 * 4 * 5 * b * 16>>>(out, nk, T, 8); });
    }
    return 0;
}

// Scratch probe 5: split roles in one launch -- "mode" CTAs write rows 3,4 of every time step with
// the thread-per-k pattern, "background writer" CTAs write rows 0..2 (3 MB contiguous per time
// step) as a linear stream.  Compared with the all-5-rows thread-per-k pattern of probe 1.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void split_roles(double2* out, long nk, long T, int work, int nmode, int mode_threads) {
    if ((int)blockIdx.x < nmode) {
        if ((int)threadIdx.x >= mode_threads) return;
        long k = (long)blockIdx.x * mode_threads + threadIdx.x;
        if (k >= nk) return;
        double x = 1.0 + k * 1e-9, v = 0.5;
        double2* p = out + 3 * nk + k;
        for (long n = 0; n < T; ++n) {
            for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
            __stcs(p, make_double2(x, v));
            __stcs(p + nk, make_double2(v, x));
            p += 5 * nk;
        }
    } else {
        const int w = blockIdx.x - nmode, nw = gridDim.x - nmode;
        for (long n = w; n < T; n += nw) {
            double2* row = out + n * 5 * nk;
            const double2 val = make_double2(1.0 + n, 0.0);
            for (long i = threadIdx.x; i < 3 * nk; i += blockDim.x) __stcs(row + i, val);
        }
    }
}

__global__ void all_rows(double2* out, long nk, long T, int work) {
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
        printf("%-52s %.3f ms  %.1f GB/s  %s\n", name, best, n * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    timeit("all 5 rows, thread-per-k, block 448 (baseline)", [&] { all_rows<<<147, 448>>>(out, nk, T, 24); });
    for (int writers : {74, 148, 296}) for (int wt : {256, 512}) {
        char name[96];
        snprintf(name, 96, "split: 147 mode CTAs(448) + %d writers(%d thr)", writers, wt);
        int block = wt > 448 ? wt : 448;
        timeit(name, [&] { split_roles<<<147 + writers, block>>>(out, nk, T, 24, 147, 448); });
    }
    return 0;
}

// Scratch probe 6: does keeping the CTAs in time lock-step (so that the union of their stores is a
// linear sweep through yresult) recover the linear-fill bandwidth?  Cooperative launch, grid
// barrier every R rows.
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k2_sync(double2* out, long nk, long T, int work, int R) {
    cg::grid_group grid = cg::this_grid();
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double x = 1.0 + k * 1e-9, v = 0.5;
    double2* p = out + k;
    for (long n = 0; n < T; ++n) {
        for (int w = 0; w < work; ++w) { x = fma(x, 0.999999, v); v = fma(v, 1.000001, -1e-7 * x); }
        if (k < nk) {
#pragma unroll
            for (int var = 0; var < 5; ++var) __stcs(p + var * nk, make_double2(x, v));
        }
        p += 5 * nk;
        if (R > 0 && (n % R) == R - 1) grid.sync();
    }
}

int main() {
    long nk = 65536, T = 7854;
    size_t n = (size_t)nk * T * 5;
    double2* out; cudaMalloc(&out, n * sizeof(double2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int block = 448, grid = 147, work = 24;
    for (int R : {0, 256, 64, 16, 4, 1}) {
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            void* args[] = {&out, &nk, &T, &work, &R};
            cudaEventRecord(e0);
            cudaLaunchCooperativeKernel((void*)k2_sync, dim3(grid), dim3(block), args, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        printf("grid barrier every %3d rows: %.3f ms  %.1f GB/s  %s\n", R, best, n * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

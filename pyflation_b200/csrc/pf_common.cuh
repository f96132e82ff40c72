// Shared host/device helpers for the pyflation_b200 C-ABI library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/pyflation_b200.h"

namespace pf {

// thread-local error text behind pf_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);

#define PF_CUDA_CHECK(expr, where)                         \
    do {                                                   \
        cudaError_t _e = (expr);                           \
        if (_e != cudaSuccess) return pf::cuda_fail(_e, where); \
    } while (0)

__host__ __device__ constexpr int nvar_of(int nf) { return 2 * nf * nf + 2 * nf + 1; }
__host__ __device__ constexpr int nbg_of(int nf) { return 2 * nf + 1; }
// record = 4 stage slots + bg row (2nf+1), padded to an even number of doubles.  A slot holds
// (A, R, C[nf][nf]) for the compiled instantiations (nf <= PF_MAX_NFIELDS; the factored nflation
// records of nf >= 5 use a prefix of it), and (A, R, d, U, 1/H^2, 0, p[nf], g[nf]) for the
// runtime-nf kernels above that (pf_rt.cu).
__host__ __device__ constexpr int rec_stage_width(int nf) {
    return nf <= PF_MAX_NFIELDS ? 2 + nf * nf : 6 + 2 * nf;
}
__host__ __device__ constexpr int rec_bg_offset(int nf) { return 4 * rec_stage_width(nf); }
__host__ __device__ constexpr int rec_of(int nf) {
    return ((2 * nf + 1 + 4 * rec_stage_width(nf)) + 1) & ~1;
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// NumPy's float64 add-reduce over n <= 128 contiguous terms (DOUBLE_pairwise_sum): fewer than eight
// terms are added in order; otherwise eight running sums take terms j, j+8, j+16, ..., are combined
// as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), and the last n % 8 terms are added in order.  The sums
// over fields of the reference -- np.sum in cmpotentials.py:837 (nflation's U) and
// cosmomodels.py:569 (sum of phi'^2) -- round this way, and a last-bit difference in the background
// is what the Bunch-Davies phase magnifies, so the device sums follow the same order.
template <class F>
__host__ __device__ __forceinline__ double np_sum(int n, F term) {
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) res += term(i);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = term(j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += term(i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += term(i);
    return res;
}

int check_model(const pf_model* m);

// Flavour of the step records for (potential, nfields): PF_REC_DENSE (A, R, C[nf][nf]) or
// PF_REC_FACTORED (nflation, nf >= 5: the rank-2 factors of C, see pf_records.cuh).  The
// environment variable PF_FO_DENSE forces dense records (A/B measurements).
enum { PF_REC_DENSE = 0, PF_REC_FACTORED = 1 };
int rec_kind(const pf_model* m);

}  // namespace pf

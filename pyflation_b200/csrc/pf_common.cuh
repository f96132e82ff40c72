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

int check_model(const pf_model* m);

// Flavour of the step records for (potential, nfields): PF_REC_DENSE (A, R, C[nf][nf]) or
// PF_REC_FACTORED (nflation, nf >= 5: the rank-2 factors of C, see pf_records.cuh).  The
// environment variable PF_FO_DENSE forces dense records (A/B measurements).
enum { PF_REC_DENSE = 0, PF_REC_FACTORED = 1 };
int rec_kind(const pf_model* m);

}  // namespace pf

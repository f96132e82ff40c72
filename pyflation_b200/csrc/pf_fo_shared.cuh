// Argument block and TMA / mbarrier helpers shared by the first-order kernels.
#pragma once
#include "pf_common.cuh"

namespace pf {

struct FoArgs {
    int64_t nk;
    const double* __restrict__ k;
    const int64_t* __restrict__ tsix;
    const double* __restrict__ ystart;   // c128 [nvar][nk]
    const double* __restrict__ rec;      // [T][REC]
    int64_t T, t0, t1;
    double* __restrict__ state;          // c128 [2 nf^2][nk]
    double* __restrict__ yout;           // c128 [rows][nvar][opitch], this batch at columns [0, nk)
    int64_t opitch;                      // modes per output row: nk, or the width of the assembled
                                         // multi-GPU result this slab is written into
    int64_t out_every;
    int32_t* flags;
    double h;
    int32_t layout;                      // PF_LAYOUT_FULL or PF_LAYOUT_PERT
    int32_t* lockstep;                   // arrival counter of a persistent lock-step launch, or null
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}


// Fused launches: records are produced by another CTA of the same grid; wait (acquire) until
// `need` of them are published, then order the generic-proxy writes before the TMA read.
// A wait that outlasts ~20 s (the producer died, or the CTAs were not co-resident under a
// debugger / profiler) does not hang and does not kill the context: it sets PF_FLAG_TIMEOUT in
// the flags word and lets the CTA run on, so the host raises from check_flags() and can retry
// with PF_FO_NOFUSE (three plain launches, no inter-CTA protocol).
constexpr unsigned PF_SPIN_LIMIT = 1u << 27;

__device__ __forceinline__ void wait_progress(const long long* progress, long long need, int32_t* flags) {
    long long have;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.s64 %0, [%1];" : "=l"(have) : "l"(progress) : "memory");
        if (have >= need) break;
        __nanosleep(128);
        if (++spins > PF_SPIN_LIMIT) {
            if (flags) atomicOr(flags, PF_FLAG_TIMEOUT);
            break;
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}

// Time lock-step of co-resident consumer CTAs (full-trajectory launches): a monotonically
// increasing arrival counter in global memory; thread 0 adds 1 and waits until all consumers of
// this epoch have arrived (`base`: arrivals to discount, 0 for a single pass).
struct LockStep {
    int* counter;                                     // null: no lock-step
    int nconsumers;
    int rows;                                         // barrier period
    int32_t* flags;                                   // PF_FLAG_TIMEOUT goes here
    int base;
};

__device__ __forceinline__ void lockstep_barrier(const LockStep& ls, int64_t row) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const int target = ls.base + (int)((row / ls.rows) + 1) * ls.nconsumers;
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ls.counter) : "memory");
        int have;
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(ls.counter) : "memory");
            if (have >= target) break;
            __nanosleep(64);
            if (++spins > PF_SPIN_LIMIT) {
                if (ls.flags) atomicOr(ls.flags, PF_FLAG_TIMEOUT);
                break;
            }
        }
    }
    __syncthreads();
}

struct PotParams { double v[PF_MAX_PARAMS]; };

int sm_count();                                      // pf_capi.cu
int launch_fo1(const FoArgs& a, cudaStream_t st);   // pf_fo1.cu
bool launch_fo_fused_generic(const pf_model* m, const FoArgs& a, double simtstart, int64_t n0,
                             const double* y0_host, double* rec, double* stagey, int32_t* ws,
                             cudaStream_t st);   // pf_fused.cu (nf = 2)
// runtime-nf path (pf_rt.cu): nflation with PF_MAX_NFIELDS < nfields <= PF_MAX_NFIELDS_RT
int rt_bg_run(const pf_model* m, int64_t T, double h, int64_t n0, const double* y0_host, double* bg,
              double* stagey, cudaStream_t st);
int rt_stage_table(const pf_model* m, int64_t T, double simtstart, double h, int64_t n0,
                   const double* stagey, double* rec, cudaStream_t st);
int rt_fo_run(const pf_model* m, const FoArgs& a, cudaStream_t st);
bool wide_fo_run(int nf, const FoArgs& a, cudaStream_t st);   // pf_fo_wide.cu: nf = 9..16
int rt_potential_eval(const pf_model* m, int64_t n, const double* y, double* U, double* dU, double* d2U,
                      cudaStream_t st);
int rt_derivs(const pf_model* m, int first_order, int64_t ncol, const double* y, double t, const double* k,
              double* dydx, cudaStream_t st);

bool launch_fo1_fused(const pf_model* m, const FoArgs& a, double simtstart, int64_t n0,
                      const double* y0_host, double* rec, int32_t* ws, cudaStream_t st);

}  // namespace pf

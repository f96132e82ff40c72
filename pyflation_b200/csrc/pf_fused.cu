// Fused K1 + K1b + K3 launch for nf = 2 (the nf = 1 version, with its own tuned consumer, is in
// pf_fo1.cu).  Measured on B200 against the three-kernel sequence: nf = 2 full trajectory 2^16
// modes 16.1 vs 17.6 ms, strided 2^20 modes 120.6 vs 125.4 ms; for nf = 8 the fused kernel
// was 15 % SLOWER (the producer's register needs are added to a consumer that is already
// register-bound), so nf >= 3 keeps the separate launches.  One CTA of the grid -- the first to take a ticket -- integrates the background
// and publishes the step records; all other CTAs are fo_consume<NF,NC> consumers that acquire
// the published record count before every TMA chunk.  See pf_fo1.cu for the protocol.
//
//   warp 0, lane 0 : serial RK4 of (phi_i, phi'_i, H); the four stage arguments of every step go
//                    to the global scratch `stagey` (room for all T steps, so the integrator
//                    never waits); s_head is bumped every 8 steps;
//   warps 1-3      : 96 workers sharing the 128 (step, stage) items of a 32-step batch:
//                    stage_record() -> global records; __threadfence, named barrier,
//                    release-store of the count.
#include "pf_common.cuh"
#include "pf_fo_generic.cuh"
#include "pf_fo_shared.cuh"
#include "pf_potentials.cuh"
#include "pf_records.cuh"

namespace pf {

constexpr int FUSED_PROD_THREADS = 128;    // = the consumer CTA size: same register budget as fo_rk4_kernel
constexpr int FUSED_WORKERS = FUSED_PROD_THREADS - 32;

struct ProdArgsN {
    PotParams par;
    double ainit, simtstart;
    double y0[2 * PF_MAX_NFIELDS + 1];
    int64_t n0;
    double* rec;        // [T][REC]
    double* stagey;     // [T][4][NB]
    int32_t* ws;        // [0] ticket, [1] flags, [2..3] published record count (s64)
};

template <int POT, int NF>
__device__ __forceinline__ void fo_produce(const FoArgs& a, const ProdArgsN& f) {
    constexpr int NB = 2 * NF + 1, REC = rec_of(NF);
    __shared__ volatile long long s_head;
    const int tid = threadIdx.x;
    if (tid >= FUSED_PROD_THREADS) return;
    const int64_t T = a.T, n0 = f.n0;
    long long* progress = reinterpret_cast<long long*>(f.ws + 2);
    if (tid == 0) s_head = n0;
    // records of rows before n0 are NaN, exactly as pf_stage_table leaves them
    for (int64_t i = tid; i < n0 * REC; i += FUSED_PROD_THREADS) f.rec[i] = qnan();
    __threadfence();
    asm volatile("bar.sync 2, %0;" ::"n"(FUSED_PROD_THREADS) : "memory");

    if (tid < 32) {
        if (tid != 0) return;
        const double h = a.h, hh = h * 0.5, h6 = h / 6.0;    // rk4.py:31-32
        double y[NB], k1[NB], k2[NB], k3[NB], k4[NB], yt[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) y[c] = f.y0[c];
        for (int64_t n = n0; n < T; ++n) {
            double* srow = f.stagey + n * (4 * NB);
#pragma unroll
            for (int c = 0; c < NB; ++c) srow[c] = y[c];
            bg_rhs<POT, NF>(f.par.v, y, k1);
#pragma unroll
            for (int c = 0; c < NB; ++c) { yt[c] = y[c] + hh * k1[c]; srow[NB + c] = yt[c]; }
            bg_rhs<POT, NF>(f.par.v, yt, k2);
#pragma unroll
            for (int c = 0; c < NB; ++c) { yt[c] = y[c] + hh * k2[c]; srow[2 * NB + c] = yt[c]; }
            bg_rhs<POT, NF>(f.par.v, yt, k3);
#pragma unroll
            for (int c = 0; c < NB; ++c) { yt[c] = y[c] + h * k3[c]; srow[3 * NB + c] = yt[c]; }
            bg_rhs<POT, NF>(f.par.v, yt, k4);
#pragma unroll
            for (int c = 0; c < NB; ++c)                      // rk4.py:46,53
                y[c] = y[c] + h6 * (k1[c] + k4[c] + 2.0 * (k3[c] + k2[c]));
            if ((n & 7) == 7 || n == T - 1) {
                __threadfence_block();
                s_head = n + 1;
            }
        }
        return;
    }

    const int w = tid - 32;                               // 0..95
    for (int64_t b = n0; b < T; b += 32) {
        const int64_t need = (b + 32 < T) ? b + 32 : T;
        unsigned spins = 0;
        while (s_head < need) {
            if (++spins > (1u << 30)) { atomicOr(f.ws + 1, PF_FLAG_TIMEOUT); break; }
        }
        __threadfence_block();                            // stage arguments are read after s_head
        for (int item = w; item < 128; item += FUSED_WORKERS) {
            const int64_t n = b + (item >> 2);
            const int s = item & 3;
            if (n < T) {
                double y[NB];
#pragma unroll
                for (int c = 0; c < NB; ++c) y[c] = __ldcg(f.stagey + (n * 4 + s) * NB + c);
                stage_record<POT, NF>(f.par.v, f.ainit, f.simtstart, a.h, n, s, y, f.rec + n * REC);
            }
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(FUSED_WORKERS) : "memory");
        if (w == 0)
            asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(progress), "l"((long long)need) : "memory");
    }
}

template <int POT, int NF, int NC>
__global__ void __launch_bounds__(FUSED_PROD_THREADS, PF_FO_MINB(NF)) fo_fused_kernel(FoArgs a, ProdArgsN f) {
    __shared__ int s_cta;
    if (threadIdx.x == 0) s_cta = atomicAdd(f.ws, 1);     // ticket: the first CTA to run produces
    __syncthreads();
    const int cta = s_cta;
    if (cta == 0) {
        fo_produce<POT, NF>(a, f);
        return;
    }
    const int64_t bx = (cta - 1) / NF;
    const int j = (cta - 1) % NF;
    fo_consume<NF, NC>(a, bx, j, reinterpret_cast<const long long*>(f.ws + 2));
}

template <int POT, int NF>
static void launch_one(const FoArgs& a, const ProdArgsN& f, cudaStream_t st) {
    if constexpr (NF == 2) {
        constexpr int NC = NF <= 2 ? 2 : 1;
        constexpr int KPB = FoShape<NF>::BLOCK / (3 - NC);
        const int64_t tiles = (a.nk + KPB - 1) / KPB;
        fo_fused_kernel<POT, NF, NC><<<(unsigned)(tiles * NF + 1), FUSED_PROD_THREADS, 0, st>>>(a, f);
    }
}

// Returns false when (potential, nfields) has no fused instantiation here.
bool launch_fo_fused_generic(const pf_model* m, const FoArgs& a, double simtstart, int64_t n0,
                             const double* y0_host, double* rec, double* stagey, int32_t* ws,
                             cudaStream_t st) {
    if (m->nfields != 2) return false;
    ProdArgsN f;
    for (int i = 0; i < PF_MAX_PARAMS; ++i) f.par.v[i] = m->params[i];
    f.ainit = m->ainit;
    f.simtstart = simtstart;
    for (int c = 0; c < 2 * PF_MAX_NFIELDS + 1; ++c) f.y0[c] = c < 2 * m->nfields + 1 ? y0_host[c] : 0.0;
    f.n0 = n0;
    f.rec = rec;
    f.stagey = stagey;
    f.ws = ws;
#define PF_CALL(POT, NF) launch_one<POT, NF>(a, f, st)
    const bool ok = PF_DISPATCH_POT_NF(m->potential_id, m->nfields, PF_CALL);
#undef PF_CALL
    return ok;
}

}  // namespace pf

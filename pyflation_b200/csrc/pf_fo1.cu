// K2 : single-field (nf = 1) first-order RK4 kernel, full-trajectory output -- the headline path
// (BASELINE configs 1, 2, 5).  Same mathematics and output as fo_rk4_kernel<1,2> in pf_fo.cu
// (rk4.py:27-138 + cosmomodels.py:625-675), restructured around what the profile showed
// (profiles/r1_k2_*): the kernel is bound by how fast one SM can retire store-carrying steps,
// so the per-step instruction count and the CTA shape matter more than raw occupancy.
//
//   * a thread owns one mode (4 doubles of state); the 16-double step record is read with
//     eight 128-bit broadcast shared loads;
//   * three row regimes per CTA, all CTA-uniform branches: rows before the CTA's first start
//     (pure NaN stores, no record traffic), the few rows where some modes of the CTA have
//     started and some have not (predicated path), and the steady state (no predicates);
//   * the CTA size is a launch parameter: the host picks it so that the grid is a whole number
//     of CTAs per SM (148 SMs) -- see pf_fo_run.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"
#include "pf_potentials.cuh"

namespace pf {

constexpr int FO1_CH = 64;    // steps per record chunk (64 * 128 B = 8 KB per buffer)
constexpr int FO1_REC = 16;

struct Coef {                 // one step record, nf = 1
    double A[4], R[4], C[4], bg[3];                   // R = 1/(aH)^2
};

__device__ __forceinline__ Coef load_record(const double* __restrict__ r) {
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const double2 q0 = r2[0], q1 = r2[1], q2 = r2[2], q3 = r2[3], q4 = r2[4], q5 = r2[5], q6 = r2[6];
    Coef c;
    c.A[0] = q0.x; c.R[0] = q0.y; c.C[0] = q1.x;
    c.A[1] = q1.y; c.R[1] = q2.x; c.C[1] = q2.y;
    c.A[2] = q3.x; c.R[2] = q3.y; c.C[2] = q4.x;
    c.A[3] = q4.y; c.R[3] = q5.x; c.C[3] = q5.y;
    c.bg[0] = q6.x; c.bg[1] = q6.y; c.bg[2] = r[14];
    return c;
}

// One classical RK4 step of  x' = v, v' = -S,  S = A v + (k^2 R2 + C) x   for one complex mode
// (k2 = k^2; the records hold R2 = 1/(aH)^2).
__device__ __forceinline__ void rk4_mode(const Coef& c, double k2, double h, double hh, double h6,
                                         double (&x)[2], double (&v)[2]) {
    double B[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        B[s] = fma(k2, c.R[s], c.C[s]);               // k^2 / (aH)^2 + C
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double x0 = x[q], v0 = v[q];
        const double S1 = fma(B[0], x0, c.A[0] * v0);
        double xt = fma(hh, v0, x0);
        double vt = fma(-hh, S1, v0);
        const double S2 = fma(B[1], xt, c.A[1] * vt);
        double ax = fma(2.0, vt, v0);
        double aS = fma(2.0, S2, S1);
        xt = fma(hh, vt, x0);
        vt = fma(-hh, S2, v0);
        const double S3 = fma(B[2], xt, c.A[2] * vt);
        ax = fma(2.0, vt, ax);
        aS = fma(2.0, S3, aS);
        xt = fma(h, vt, x0);
        vt = fma(-h, S3, v0);
        const double S4 = fma(B[3], xt, c.A[3] * vt);
        x[q] = fma(h6, ax + vt, x0);
        v[q] = fma(-h6, aS + S4, v0);
    }
}

// Consumer: integrates the modes of logical CTA `cta` (nthreads threads, one mode each; two
// modes per thread were measured 1.8x slower: the kernel needs its >= 14 warps per SM).
// `progress` (may be null) is the producer's published record count in fused launches.
// Time lock-step of the consumer CTAs (full-trajectory launches whose grid is one co-resident
// wave).  yresult[t] is 5 contiguous rows: while all CTAs work on the same few time steps the
// union of their stores is a linear sweep through memory and DRAM pages are written once,
// completely.  CTAs drift apart mainly in the NaN regime, whose length depends on the CTA's
// k range (its rows are pure stores and run faster than integrated rows); once every mode has
// started all CTAs do identical work per row and stay aligned by themselves.  A barrier every
// 512 rows re-aligns them: measured 6.72 -> 6.21 ms per solve together with the record prefetch
// (periods from 256 to 2048 rows are equivalent, 64 and less cost more than they gain;
// scripts/pattern_probe6.cu shows the effect on the bare store pattern: 5.8 -> 6.8 TB/s).
// The barrier is a monotonically increasing arrival counter in global memory: thread 0 adds 1
// and waits until all consumers of this epoch have arrived.
constexpr int FO1_LOCKSTEP_ROWS = 512;

struct Fo1Smem {                                      // one per CTA, owned by the kernel
    alignas(128) double srec[2][FO1_CH * FO1_REC];
    alignas(8) uint64_t bars[2];
    long long tsmin, tsmax;
};

template <bool PERT>
__device__ __forceinline__ void fo1_consume(const FoArgs& a, int64_t cta, int nthreads,
                                            const long long* progress, Fo1Smem& sm,
                                            const LockStep ls = LockStep{nullptr, 0, FO1_LOCKSTEP_ROWS, nullptr, 0}) {
    constexpr int NV = PERT ? 2 : 5;                  // rows per time step in the output
    constexpr int XO = PERT ? 0 : 3;                  // row of dphi within a time step
    double (&srec)[2][FO1_CH * FO1_REC] = sm.srec;
    uint64_t (&bars)[2] = sm.bars;
    long long& s_tsmin = sm.tsmin;
    long long& s_tsmax = sm.tsmax;

    const int tid = threadIdx.x;
    if (tid >= nthreads) return;                      // fused launches may carry spare threads
    const int64_t nk = a.nk;
    const int64_t kk = cta * nthreads + tid;

    if (tid == 0) {
        s_tsmin = INT64_MAX;
        s_tsmax = INT64_MIN;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const bool live = kk < nk;
    const double kval = live ? a.k[kk] : 0.0;
    const double k2 = kval * kval;
    const int64_t ts = live ? a.tsix[kk] : INT64_MAX;
    double x[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
    if (live) {
        atomicMin(&s_tsmin, (long long)ts);
        atomicMax(&s_tsmax, (long long)ts);
        if (a.t0 > ts) {                                      // resume a windowed run
            const double2 sx = reinterpret_cast<const double2*>(a.state)[kk];
            const double2 sv = reinterpret_cast<const double2*>(a.state)[nk + kk];
            x[0] = sx.x; x[1] = sx.y; v[0] = sv.x; v[1] = sv.y;
        }
    }
    __syncthreads();
    const int64_t tsmin = s_tsmin, tsmax = s_tsmax;
    if (tsmin == INT64_MAX) return;                           // CTA past the end of the batch

    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;         // rk4.py:31-32
    const double2 nan2 = make_double2(qnan(), qnan());
    double2* prow = reinterpret_cast<double2*>(a.yout) + kk;  // row t0, var 0, this thread's k
    const int64_t op = a.opitch;                              // modes per output row (>= nk)
    const int64_t rstride = NV * op;

    // ---- regime 1: rows before any mode of this CTA starts: NaN+NaNj, no records needed.
    // Rows are stored when (n - t0) is a multiple of out_every (0: never); next_out tracks it.
    const int64_t every = a.out_every;
    int64_t next_out = (a.yout != nullptr && every > 0) ? a.t0 : INT64_MAX;
    const int64_t nan_end = tsmin < a.t1 ? tsmin : a.t1;
    if (ls.counter) {
        for (int64_t r = a.t0; r < nan_end; ++r) {            // every == 1 in lock-step launches
            if (live) {
#pragma unroll
                for (int var = 0; var < NV; ++var) __stcs(prow + var * op, nan2);
            }
            prow += rstride;
            if ((r - a.t0) % ls.rows == ls.rows - 1) lockstep_barrier(ls, r - a.t0);
        }
        next_out = nan_end > a.t0 ? nan_end : a.t0;
    } else {
        for (; next_out < nan_end; next_out += every) {
            if (live) {
#pragma unroll
                for (int var = 0; var < NV; ++var) __stcs(prow + var * op, nan2);
            }
            prow += rstride;
        }
    }
    // a window that begins after this CTA's first start row resumes at t0 from the carried state
    // (it used to re-integrate from the start row: same bits, wasted steps)
    int64_t n = nan_end > a.t0 ? nan_end : a.t0;
    if (n >= a.t1) return;

    // ---- regimes 2 and 3: records streamed through shared memory from row c0 on
    const int64_t c0 = n;
    const int64_t nchunks = (a.t1 - c0 + FO1_CH - 1) / FO1_CH;
    auto issue = [&](int64_t c) {
        const int buf = (int)(c & 1);
        const int64_t base = c0 + c * FO1_CH;
        const int64_t cnt = (a.t1 - base) < FO1_CH ? (a.t1 - base) : FO1_CH;
        const uint32_t bytes = (uint32_t)(cnt * FO1_REC * sizeof(double));
        if (progress) wait_progress(progress, base + cnt, a.flags);
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(&srec[buf][0], a.rec + base * FO1_REC, bytes, &bars[buf]);
    };
    if (tid == 0) {
        issue(0);
        if (nchunks > 1) issue(1);
    }

    for (int64_t c = 0; c < nchunks; ++c) {
        const int buf = (int)(c & 1);
        mbar_wait(&bars[buf], (uint32_t)((c >> 1) & 1));
        const int64_t base = c0 + c * FO1_CH;
        const int cnt = (int)((a.t1 - base) < FO1_CH ? (a.t1 - base) : FO1_CH);
        // the record of step q+1 is fetched while step q computes (one step of software
        // prefetch: shared loads queue behind the back-pressured stores in the LSU)
        Coef nxt = load_record(&srec[buf][0]);
        for (int q = 0; q < cnt; ++q) {
            n = base + q;
            if (n > tsmax && n != next_out && ls.counter == nullptr) {
                // ---- tight run (strided / state-only launches, which are bound by the fp64 pipe):
                // every mode of the CTA is running and no row is stored before `lim`, so the steps
                // up to there are nothing but record loads and RK4 arithmetic -- no output tests,
                // no register copies of a prefetched record (profiles/r2_c5_*: the general loop
                // issues 136 instructions per step for 56 fp64 ones)
                int64_t lim = next_out - base < (int64_t)cnt ? next_out - base : (int64_t)cnt;
                if (a.T - 1 - base < lim) lim = a.T - 1 - base;
                if (lim > q) {
                    const int qe = (int)lim;
#pragma unroll 2
                    for (; q < qe; ++q) {
                        const Coef ct = load_record(&srec[buf][q * FO1_REC]);
                        rk4_mode(ct, k2, h, hh, h6, x, v);
                    }
                    --q;                                  // the for statement steps to qe
                    if (qe < cnt) nxt = load_record(&srec[buf][qe * FO1_REC]);
                    continue;
                }
            }
            const Coef cf = nxt;
            if (q + 1 < cnt) nxt = load_record(&srec[buf][(q + 1) * FO1_REC]);
            const bool out = (n == next_out);
            if (n > tsmax) {
                // ---- regime 3: every mode of the CTA is running
                if (live && out) {
                    if constexpr (!PERT) {
                        __stcs(prow, make_double2(cf.bg[0], 0.0));
                        __stcs(prow + op, make_double2(cf.bg[1], 0.0));
                        __stcs(prow + 2 * op, make_double2(cf.bg[2], 0.0));
                    }
                    __stcs(prow + XO * op, make_double2(x[0], x[1]));
                    __stcs(prow + (XO + 1) * op, make_double2(v[0], v[1]));
                }
                if (n + 1 < a.T) rk4_mode(cf, k2, h, hh, h6, x, v);
            } else if (live) {
                // ---- regime 2: some modes of the CTA have started, some have not
                if (n < ts) {
                    if (out) {
#pragma unroll
                        for (int var = 0; var < NV; ++var) __stcs(prow + var * op, nan2);   // rk4.py:100
                    }
                } else {
                    double2 b0 = make_double2(cf.bg[0], 0.0), b1 = make_double2(cf.bg[1], 0.0),
                            b2 = make_double2(cf.bg[2], 0.0);
                    if (n == ts) {
                        // the start row is ystart[:, k] itself            rk4.py:106-107
                        const double2* ys = reinterpret_cast<const double2*>(a.ystart) + kk;
                        b0 = ys[0]; b1 = ys[nk]; b2 = ys[2 * nk];
                        const double2 sx = ys[3 * nk], sv = ys[4 * nk];
                        x[0] = sx.x; x[1] = sx.y; v[0] = sv.x; v[1] = sv.y;
                        if (a.flags) {
                            const bool bad = !(fabs(b0.x - cf.bg[0]) <= 1e-9 * fabs(cf.bg[0]) + 1e-300) ||
                                             !(fabs(b1.x - cf.bg[1]) <= 1e-9 * fabs(cf.bg[1]) + 1e-300) ||
                                             !(fabs(b2.x - cf.bg[2]) <= 1e-9 * fabs(cf.bg[2]) + 1e-300);
                            if (bad) atomicOr(a.flags, PF_FLAG_BG_MISMATCH);
                        }
                    }
                    if (out) {
                        if constexpr (!PERT) {
                            __stcs(prow, b0);
                            __stcs(prow + op, b1);
                            __stcs(prow + 2 * op, b2);
                        }
                        __stcs(prow + XO * op, make_double2(x[0], x[1]));
                        __stcs(prow + (XO + 1) * op, make_double2(v[0], v[1]));
                    }
                    if (n + 1 < a.T) rk4_mode(cf, k2, h, hh, h6, x, v);
                }
            }
            if (out) {
                prow += rstride;
                next_out += every;
            }
            if (ls.counter && (n - a.t0) % ls.rows == ls.rows - 1)
                lockstep_barrier(ls, n - a.t0);
        }
        __syncthreads();                              // everyone is done reading srec[buf]
        if (tid == 0 && c + 2 < nchunks) issue(c + 2);
    }

    if (a.state && live && ts < a.t1) {
        reinterpret_cast<double2*>(a.state)[kk] = make_double2(x[0], x[1]);
        reinterpret_cast<double2*>(a.state)[nk + kk] = make_double2(v[0], v[1]);
    }
}

template <bool PERT>
__global__ void __launch_bounds__(512) fo1_kernel(FoArgs a) {
    __shared__ Fo1Smem sm;
    fo1_consume<PERT>(a, blockIdx.x, blockDim.x, nullptr, sm);
}

// Lock-step launch of the plain kernel (full-trajectory windows, and solves whose records already
// exist): cooperative, one 512-thread CTA per k tile, the whole grid co-resident (at most one CTA
// per SM), all CTAs advancing through the rows together.  Batches wider than one such wave do
// NOT use it: persistent CTAs taking several tiles in turn were measured slower than the plain
// multi-wave launch (2^20 modes, 1024-row window: 14.0 vs 12.7 ms; profiles/r2_ab_lockstep.txt).
template <bool PERT>
__global__ void __launch_bounds__(512) fo1_lockstep_kernel(FoArgs a) {
    __shared__ Fo1Smem sm;
    const LockStep ls{a.lockstep, (int)gridDim.x, FO1_LOCKSTEP_ROWS, a.flags, 0};
    fo1_consume<PERT>(a, blockIdx.x, blockDim.x, nullptr, sm, ls);
}

// ------------------------------------------------------------------------------------------
// Producer CTA of the fused launch: K1 (serial background RK4) and K1b (step records) run
// inside the same grid as the mode kernel, so their ~1.7 ms hide behind the trajectory stores.
//   warp 0, lane 0 : the serial RK4 chain; writes the 4 stage arguments of every step into a
//                    shared-memory ring and bumps s_head every 8 steps;
//   warps 1-4      : 128 workers, one (step, stage) item each per 32-step batch: evaluate
//                    A, R, C (potential 2nd derivative, exp) and write the record to global
//                    memory; after a batch: __threadfence, named barrier, release-store of the
//                    published record count that the consumer CTAs acquire.
// ------------------------------------------------------------------------------------------
constexpr int PROD_RING = 128;                        // steps of stage arguments in flight
constexpr int PROD_THREADS = 160;

struct ProdArgs {
    PotParams par;
    double ainit, simtstart;
    double y0[3];
    int64_t n0;
    double* rec;                                      // [T][16]
    int32_t* ws;                                      // [0] ticket, [1] flags, [2..3] progress (s64)
};

template <int POT>
__device__ __forceinline__ void bg1_rhs(const double* __restrict__ p, double phi, double pd, double H,
                                        double& dphi, double& dpd, double& dH) {
    double f[1] = {phi}, dU[1], d2U[1], U;
    potential<POT, 1, false>(p, f, U, dU, d2U);
    dphi = pd;                                        // cosmomodels.py:563
    dpd = -(U * pd + dU[0]) / (H * H);                // :566 (true division, as NumPy's real arrays)
    dH = -0.5 * (pd * pd) * H;                        // :569
}

template <int POT>
__device__ __forceinline__ void fo1_produce(const FoArgs& a, const ProdArgs& f) {
    __shared__ double ring[PROD_RING][12];
    __shared__ volatile long long s_head, s_tail;
    const int tid = threadIdx.x;
    if (tid >= PROD_THREADS) return;
    const int64_t T = a.T, n0 = f.n0;
    long long* progress = reinterpret_cast<long long*>(f.ws + 2);
    if (tid == 0) { s_head = n0; s_tail = n0; }
    // records of rows before n0 are NaN, exactly as pf_stage_table leaves them: a caller whose
    // tsix dips below n0 gets NaN trajectories, not uninitialised memory
    for (int64_t i = tid; i < n0 * FO1_REC; i += PROD_THREADS) f.rec[i] = qnan();
    __threadfence();
    asm volatile("bar.sync 2, %0;" ::"n"(PROD_THREADS) : "memory");

    if (tid < 32) {
        if (tid != 0) return;
        const double h = a.h, hh = h * 0.5, h6 = h / 6.0;
        double y0 = f.y0[0], y1 = f.y0[1], y2 = f.y0[2];
        for (int64_t n = n0; n < T; ++n) {
            if ((n & 7) == 0) {
                unsigned spins = 0;
                while (n - s_tail > PROD_RING - 16) {
                    if (++spins > (1u << 30)) { atomicOr(f.ws + 1, PF_FLAG_TIMEOUT); break; }
                }
            }
            double* slot = ring[n % PROD_RING];
            double a0, a1, a2, b0, b1, b2, c0, c1, c2, d0, d1, d2;
            slot[0] = y0; slot[1] = y1; slot[2] = y2;
            bg1_rhs<POT>(f.par.v, y0, y1, y2, a0, a1, a2);
            double t0 = y0 + hh * a0, t1 = y1 + hh * a1, t2 = y2 + hh * a2;
            slot[3] = t0; slot[4] = t1; slot[5] = t2;
            bg1_rhs<POT>(f.par.v, t0, t1, t2, b0, b1, b2);
            t0 = y0 + hh * b0; t1 = y1 + hh * b1; t2 = y2 + hh * b2;
            slot[6] = t0; slot[7] = t1; slot[8] = t2;
            bg1_rhs<POT>(f.par.v, t0, t1, t2, c0, c1, c2);
            t0 = y0 + h * c0; t1 = y1 + h * c1; t2 = y2 + h * c2;
            slot[9] = t0; slot[10] = t1; slot[11] = t2;
            bg1_rhs<POT>(f.par.v, t0, t1, t2, d0, d1, d2);
            y0 = y0 + h6 * (a0 + d0 + 2.0 * (c0 + b0));      // rk4.py:46,53
            y1 = y1 + h6 * (a1 + d1 + 2.0 * (c1 + b1));
            y2 = y2 + h6 * (a2 + d2 + 2.0 * (c2 + b2));
            if ((n & 7) == 7 || n == T - 1) {
                __threadfence_block();
                s_head = n + 1;
            }
        }
        return;
    }

    const int w = tid - 32;                           // 0..127
    const int dn = w >> 2, s = w & 3;
    for (int64_t b = n0; b < T; b += 32) {
        const int64_t need = (b + 32 < T) ? b + 32 : T;
        {
            unsigned spins = 0;
            while (s_head < need) {
                if (++spins > (1u << 30)) { atomicOr(f.ws + 1, PF_FLAG_TIMEOUT); break; }
            }
            __threadfence_block();                    // ring reads stay behind the s_head read
        }
        const int64_t n = b + dn;
        if (n < T) {
            const double* slot = ring[n % PROD_RING] + 3 * s;
            const double phi = slot[0], pd = slot[1], H = slot[2];
            const double x = f.simtstart + (double)n * a.h;           // rk4.py:118 last_x
            const double t = (s == 0) ? x : ((s == 3) ? x + a.h : x + a.h * 0.5);
            const double sf = f.ainit * exp(t);                       // cosmomodels.py:637
            double ff[1] = {phi}, dU[1], d2U[1], U;
            potential<POT, 1, true>(f.par.v, ff, U, dU, d2U);
            const double invH2 = 1.0 / (H * H);
            double* r = f.rec + n * FO1_REC;
            r[3 * s + 0] = U * invH2;
            const double rr = 1.0 / (sf * H);
            r[3 * s + 1] = rr * rr;
            r[3 * s + 2] = (d2U[0] + pd * dU[0] + dU[0] * pd + pd * pd * U) * invH2;   // :667-674
            if (s == 0) { r[12] = phi; r[13] = pd; r[14] = H; r[15] = 0.0; }
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (w == 0) {
            s_tail = need;
            asm volatile("st.release.gpu.global.s64 [%0], %1;" ::"l"(progress), "l"((long long)need) : "memory");
        }
    }
}

template <int POT>
__global__ void __launch_bounds__(512) fo1_fused_kernel(FoArgs a, ProdArgs f, int cthreads, int lockstep) {
    __shared__ int s_cta;
    if (threadIdx.x == 0) s_cta = atomicAdd(f.ws, 1);   // ticket: the first CTA to run produces
    __syncthreads();
    const int cta = s_cta;
    if (cta == 0) {
        fo1_produce<POT>(a, f);
        return;
    }
    const long long* progress = reinterpret_cast<const long long*>(f.ws + 2);
    __shared__ Fo1Smem sm;
    const LockStep ls{lockstep ? f.ws + 4 : nullptr, (int)gridDim.x - 1, lockstep > 1 ? lockstep : FO1_LOCKSTEP_ROWS, f.ws + 1, 0};
    if (a.layout == PF_LAYOUT_PERT) fo1_consume<true>(a, cta - 1, cthreads, progress, sm, ls);
    else fo1_consume<false>(a, cta - 1, cthreads, progress, sm, ls);
}

// Grid shape: a whole number c of CTAs per SM, c the smallest for which a CTA holds at most
// `max_threads` threads (so every SM gets the same number of equally sized CTAs).
// CTA size.  Full-trajectory launches are bound by the DRAM write pattern: 512-thread CTAs make
// every row segment of a CTA an aligned 8 KB block (measured with lock-step on three different
// B200s: 5.78-5.93 ms vs 6.21-6.26 ms for the balanced 448-thread grid), so they are used
// whenever that grid still fits one wave; otherwise a whole number of equally sized CTAs per SM.
// Strided / state-only launches are bound by the fp64 pipe and run ~10 % faster with 128-thread
// CTAs (scripts/dev_probe.py c5).
static int64_t fo1_block(int64_t nk, int sms, int64_t out_every) {
    int max_threads = out_every == 1 ? 512 : 128;
    if (const char* e = getenv("PF_FO1_MAXT")) max_threads = atoi(e);
    if (max_threads < 32 || max_threads > 512) max_threads = 512;
    int64_t block = 0;
    if (out_every == 1 && max_threads == 512 && nk >= 16384 && (nk + 511) / 512 <= sms) {
        block = 512;
    } else {
        for (int64_t per_sm = 1;; ++per_sm) {
            block = (nk + sms * per_sm - 1) / (sms * per_sm);
            block = (block + 31) / 32 * 32;
            if (block <= max_threads) break;
        }
    }
    if (const char* e = getenv("PF_FO1_BLOCK")) block = atoi(e);
    return block;
}

int launch_fo1(const FoArgs& a, cudaStream_t st) {
    static const bool no_lockstep = getenv("PF_FO1_NOLOCKSTEP") != nullptr;
    const int64_t ntiles = (a.nk + 511) / 512;
    if (a.lockstep && !no_lockstep && a.out_every == 1 && a.nk >= 16384 && ntiles <= sm_count() &&
        a.t1 - a.t0 >= 2 * FO1_LOCKSTEP_ROWS) {
        FoArgs aa = a;
        void* args[] = {&aa};
        void* fn = a.layout == PF_LAYOUT_PERT ? (void*)fo1_lockstep_kernel<true> : (void*)fo1_lockstep_kernel<false>;
        if (cudaMemsetAsync(a.lockstep, 0, sizeof(int32_t), st) == cudaSuccess &&
            cudaLaunchCooperativeKernel(fn, dim3((unsigned)ntiles), dim3(512), args, 0, st) == cudaSuccess)
            return 0;
        (void)cudaGetLastError();
    }
    const int64_t block = fo1_block(a.nk, sm_count(), a.out_every);
    const int64_t grid = (a.nk + block - 1) / block;
    if (a.layout == PF_LAYOUT_PERT) fo1_kernel<true><<<(unsigned)grid, (unsigned)block, 0, st>>>(a);
    else fo1_kernel<false><<<(unsigned)grid, (unsigned)block, 0, st>>>(a);
    return 0;
}

// Fused K1 + K1b + K2 for nf = 1 (single-field potentials and nflation with one field).
// Returns false when (potential, nfields) has no fused instantiation.
bool launch_fo1_fused(const pf_model* m, const FoArgs& a, double simtstart, int64_t n0,
                      const double* y0_host, double* rec, int32_t* ws, cudaStream_t st) {
    ProdArgs f;
    for (int i = 0; i < PF_MAX_PARAMS; ++i) f.par.v[i] = m->params[i];
    f.ainit = m->ainit;
    f.simtstart = simtstart;
    f.y0[0] = y0_host[0]; f.y0[1] = y0_host[1]; f.y0[2] = y0_host[2];
    f.n0 = n0;
    f.rec = rec;
    f.ws = ws;
    const int64_t cthreads = fo1_block(a.nk, sm_count() - 1, a.out_every);   // one SM hosts the producer
    const int64_t block = cthreads > PROD_THREADS ? cthreads : PROD_THREADS;
    const int64_t grid = (a.nk + cthreads - 1) / cthreads + 1;
    // Lock-step needs every CTA resident at once: ask for a cooperative launch (it is refused when
    // the grid does not fit in one wave) and only then switch the barrier on.
    static const bool no_lockstep = getenv("PF_FO1_NOLOCKSTEP") != nullptr;
    const bool want_lockstep = a.out_every == 1 && grid <= 8 * (int64_t)sm_count() && !no_lockstep;
#define PF_CALL(POT, NF)                                                                          \
    do {                                                                                          \
        int ct = (int)cthreads, on = 1;                                                           \
        FoArgs aa = a;                                                                            \
        ProdArgs ff = f;                                                                          \
        void* args[] = {&aa, &ff, &ct, &on};                                                      \
        cudaError_t e = want_lockstep                                                             \
            ? cudaLaunchCooperativeKernel((void*)fo1_fused_kernel<POT>, dim3((unsigned)grid),     \
                                          dim3((unsigned)block), args, 0, st)                     \
            : cudaErrorCooperativeLaunchTooLarge;                                                 \
        if (e != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                             \
            fo1_fused_kernel<POT><<<(unsigned)grid, (unsigned)block, 0, st>>>(a, f, (int)cthreads, 0); \
        }                                                                                         \
    } while (0)
    switch (m->potential_id) {
        case PF_MSQPHISQ: PF_CALL(PF_MSQPHISQ, 1); break;
        case PF_LAMBDAPHI4: PF_CALL(PF_LAMBDAPHI4, 1); break;
        case PF_LINDE: PF_CALL(PF_LINDE, 1); break;
        case PF_HYBRID2AND4: PF_CALL(PF_HYBRID2AND4, 1); break;
        case PF_PHI2OVER3: PF_CALL(PF_PHI2OVER3, 1); break;
        case PF_MSQPHISQ_WITHV0: PF_CALL(PF_MSQPHISQ_WITHV0, 1); break;
        case PF_STEP_POTENTIAL: PF_CALL(PF_STEP_POTENTIAL, 1); break;
        case PF_BUMP_POTENTIAL: PF_CALL(PF_BUMP_POTENTIAL, 1); break;
        case PF_RESONANCE: PF_CALL(PF_RESONANCE, 1); break;
        case PF_BUMP_NOTHIRDDERIV: PF_CALL(PF_BUMP_NOTHIRDDERIV, 1); break;
        case PF_NFLATION: PF_CALL(PF_NFLATION, 1); break;
        default: return false;
    }
#undef PF_CALL
    return true;
}

}  // namespace pf

// K9 : fixed-step RK4 of the second-order perturbation over a batch of k modes.
//
// Replaces rk4.rkdriver_tsix (rk4.py:58-138) + rk4stepks (:27-55) driving
// CanonicalSecondOrder.derivs (cosmomodels.py:708-766), CanonicalRampedSecondOrder.derivs
// (:893-971) and, for single evaluations, CanonicalHomogeneousSecondOrder.derivs (:797-848),
// as SOCanonicalThreeStage.runso sets them up (:1680-1790).
//
// Per mode the state is the real 4-vector (Re d2, Re d2', Im d2, Im d2'); the real and the
// imaginary pair obey the same equation with the real / imaginary part of the source term and
// never mix:
//     y0' = y1,   y1' = -( (3 - eps) y1 + (k/(aH))^2 y0 + (V''/H^2 - 3 phi'^2) y0 + S )
// Everything except k and S is a function of the stage time alone, so the host tabulates it once
// per run ("coefficients": three stage times per step, PF_SO_COEF doubles each) and the kernel
// is what remains: one thread per mode, 32 B of source read and 32 B of result written per mode
// and step -- HBM-bound.  The source values of the next SO_AHEAD steps are copied into shared memory before they are
// needed (they do not depend on the state), which is what keeps enough bytes in flight.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"

namespace pf {

#ifndef PF_SO_BLOCK
#define PF_SO_BLOCK 64
#endif
#ifndef PF_SO_AHEAD
#define PF_SO_AHEAD 4
#endif
#ifndef PF_SO_MINB
#define PF_SO_MINB 8
#endif
constexpr int SO_BLOCK = PF_SO_BLOCK;
constexpr int SO_AHEAD = PF_SO_AHEAD;          // steps of source look-ahead (cp.async ring)
constexpr int SO_C = PF_SO_COEF;

struct SoArgs {
    int64_t nk;
    const double* __restrict__ k;
    const int64_t* __restrict__ tsix;      // second-order start rows
    const int64_t* __restrict__ rtsix;     // rows the ramp compares t/h with (the model's tstartindex)
    const int64_t* __restrict__ fo_tsix;   // first-order start rows (null: never poison)
    const double* __restrict__ ystart;     // [4][nk]
    const double* __restrict__ coef;       // [T][3][SO_C]
    const double* __restrict__ src;        // c128 [T_fo][src_pitch] or null
    int64_t src_pitch;
    const double* __restrict__ tstart;     // [nk] first-order start times (ramp), or null
    double ra, rb, rc, rd, re;             // ramp = (tanh(ra (t - tstart - rb)) + rc) / rd if |.| < re
    int32_t ramp;
    int64_t T, t0, t1;
    double h;
    double* __restrict__ state;            // [4][nk]
    double* __restrict__ yout;             // [rows][4][opitch]
    int64_t opitch, out_every;
};

// One stage slot of the coefficient table (48 bytes, 16-byte aligned); fotix is an int64 stored in
// the slot's fourth 8-byte word (no fp64 -> integer conversion on the device).
struct SoStage {
    double c1, R2, c2;
    int64_t fotix;                        // + t, t/h in the slot's last 16 bytes (ramp only)
};

// p: shared memory (the solver) or global memory (single right-hand sides)
__device__ __forceinline__ SoStage load_stage(const double* __restrict__ p) {
    const double2 a = reinterpret_cast<const double2*>(p)[0];
    const double2 b = reinterpret_cast<const double2*>(p)[1];
    SoStage s;
    s.c1 = a.x; s.R2 = a.y; s.c2 = b.x; s.fotix = __double_as_longlong(b.y);
    return s;
}

// ramp factor of one stage (cosmomodels.py:925-944); 1 outside the ramp window.  Kept out of
// line: it runs for ~1 e-fold per mode, and twelve inlined copies of tanh would bloat the loop.
__device__ __noinline__ double ramp_factor(const double* __restrict__ slot, double tstart, double tsd, double ra,
                                           double rb, double rc, double rd, double re) {
    const double2 tt = reinterpret_cast<const double2*>(slot)[2];
    const double t = tt.x, tdivh = tt.y;
    const double arg = t - tstart - rb;
    if (!(fabs(arg) < re)) return 1.0;
    if (tsd == tdivh) return 0.0;                         // ramp[self.tstartindex == sotix] = 0
    return (tanh(ra * arg) + rc) / rd;
}

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only): completion is tracked by commit
// groups, not by a register scoreboard, so a look-ahead of several steps costs no registers and a
// consumer never waits on younger loads (the register-ring version did: profiles/r2c_so_ncu_notes.txt)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

// y1' = -( c1 y1 + B y0 + S ),  B = k^2 R2 + c2                  cosmomodels.py:754-755, 761-762
__device__ __forceinline__ double so_accel(double c1, double B, double y0, double y1, double s) {
    return -fma(c1, y1, fma(B, y0, s));
}

struct SoState { double yr0, yr1, yi0, yi1; };          // (Re d2, Re d2'), (Im d2, Im d2')

// one classical RK4 step, operation order of rk4.py:27-55, for both pairs
__device__ __forceinline__ void so_step(SoState& y, const SoStage& c0, const SoStage& c1, const SoStage& c2,
                                        double k2, double2 s0, double2 s1, double2 s2, double h, double hh,
                                        double h6) {
    const double B0 = fma(k2, c0.R2, c0.c2), B1 = fma(k2, c1.R2, c1.c2), B2 = fma(k2, c2.R2, c2.c2);
    const double dr = so_accel(c0.c1, B0, y.yr0, y.yr1, s0.x), di = so_accel(c0.c1, B0, y.yi0, y.yi1, s0.y);
    double ur0 = fma(hh, y.yr1, y.yr0), ur1 = fma(hh, dr, y.yr1);
    double ui0 = fma(hh, y.yi1, y.yi0), ui1 = fma(hh, di, y.yi1);
    const double er0 = ur1, er1 = so_accel(c1.c1, B1, ur0, ur1, s1.x);      // dyt
    const double ei0 = ui1, ei1 = so_accel(c1.c1, B1, ui0, ui1, s1.y);
    ur0 = fma(hh, er0, y.yr0); ur1 = fma(hh, er1, y.yr1);
    ui0 = fma(hh, ei0, y.yi0); ui1 = fma(hh, ei1, y.yi1);
    double fr0 = ur1, fr1 = so_accel(c1.c1, B1, ur0, ur1, s1.x);            // dym
    double fi0 = ui1, fi1 = so_accel(c1.c1, B1, ui0, ui1, s1.y);
    ur0 = fma(h, fr0, y.yr0); ur1 = fma(h, fr1, y.yr1);
    ui0 = fma(h, fi0, y.yi0); ui1 = fma(h, fi1, y.yi1);
    fr0 += er0; fr1 += er1; fi0 += ei0; fi1 += ei1;
    const double gr1 = so_accel(c2.c1, B2, ur0, ur1, s2.x), gi1 = so_accel(c2.c1, B2, ui0, ui1, s2.y);
    y.yr0 = fma(h6, y.yr1 + ur1 + 2.0 * fr0, y.yr0);
    y.yi0 = fma(h6, y.yi1 + ui1 + 2.0 * fi0, y.yi0);
    y.yr1 = fma(h6, dr + gr1 + 2.0 * fr1, y.yr1);
    y.yi1 = fma(h6, di + gi1 + 2.0 * fi1, y.yi1);
}

// One thread per mode: both (Re, Im) pairs advance together (two independent dependency chains),
// the stage coefficients, the ramp and the bookkeeping are shared between them.  A warp reads one
// contiguous 512-byte piece of a source row per stage and writes four 256-byte pieces of result
// rows per step.
//
// The coefficient table streams through shared memory in double-buffered chunks (1-D TMA bulk
// copies + mbarrier, as the first-order kernels do with their records): every thread of the CTA
// reads the same slot (broadcast), and -- what matters more -- those reads then wait on the
// short scoreboard instead of sharing a long-scoreboard slot with source loads that are in
// flight to HBM (ncu of the global-load version: two ~800-cycle stalls per step on L1-resident
// coefficients, profiles/r2c_so_ncu_notes.txt).  All threads of a CTA walk the rows in step; a
// mode that has not started yet only writes its NaN rows.
constexpr int SO_CH = (30 / SO_AHEAD) * SO_AHEAD;       // steps per chunk (multiple of SO_AHEAD)
constexpr int SO_CHL = SO_CH + SO_AHEAD;                // + look-ahead slots for the source prefetch
static_assert(SO_CH % SO_AHEAD == 0, "chunk must be a multiple of the look-ahead");

__global__ void __launch_bounds__(SO_BLOCK, PF_SO_MINB) so_rk4_kernel(SoArgs a) {
    __shared__ __align__(128) double scoef[2][SO_CHL * 3 * SO_C];
    __shared__ __align__(16) double2 sring[SO_AHEAD][3][SO_BLOCK];   // per-thread source look-ahead
    __shared__ __align__(8) uint64_t bars[2];

    const int tid = threadIdx.x;
    const int64_t kraw = (int64_t)blockIdx.x * SO_BLOCK + tid;
    const bool live = kraw < a.nk;
    const int64_t kk = live ? kraw : a.nk - 1;            // idle lanes shadow the last mode, store nothing
    const int64_t op = a.opitch, nk = a.nk;
    const int64_t ts = a.tsix[kk];
    const double k2 = a.k[kk] * a.k[kk];
    const double tstart = a.tstart ? a.tstart[kk] : 0.0;
    const double tsd = (double)(a.rtsix ? a.rtsix[kk] : ts);
    const int64_t fo_ts = a.fo_tsix ? a.fo_tsix[kk] : INT64_MIN;
    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;     // rk4.py:31-32
    const double nan = qnan();

    const int64_t nrows = a.t1 - a.t0;
    const int64_t nchunks = (nrows + SO_CH - 1) / SO_CH;
    auto issue = [&](int64_t c) {
        const int buf = (int)(c & 1);
        const int64_t base = a.t0 + c * SO_CH;
        const int64_t cnt = (a.T - base) < SO_CHL ? (a.T - base) : SO_CHL;
        const uint32_t bytes = (uint32_t)(cnt * 3 * SO_C * sizeof(double));
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(&scoef[buf][0], a.coef + base * 3 * SO_C, bytes, &bars[buf]);
    };
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        if (nchunks > 0) issue(0);
        if (nchunks > 1) issue(1);
    }

    int64_t next_out = (a.yout != nullptr && a.out_every > 0 && live) ? a.t0 : INT64_MAX;
    double* orow = a.yout ? a.yout + kk : nullptr;

    SoState y = {0.0, 0.0, 0.0, 0.0};
    if (a.t0 > ts) {                                      // resume a windowed run
        y.yr0 = a.state[kk]; y.yr1 = a.state[nk + kk]; y.yi0 = a.state[2 * nk + kk]; y.yi1 = a.state[3 * nk + kk];
    }
    // the ramp can only be active for steps in [ramp_lo, ramp_hi] (stage times within re of
    // tstart + rb; two steps of slack on either side), checked exactly inside
    int64_t ramp_lo = INT64_MAX, ramp_hi = INT64_MIN;
    if (a.ramp) {
        const double x0 = __ldg(a.coef + 4);              // t of step 0, stage 0 = simtstart
        const double c = tstart + a.rb - x0;
        const double lo = floor((c - a.re) / h) - 2.0, hi = ceil((c + a.re) / h) + 2.0;
        ramp_lo = lo < -9e18 ? INT64_MIN : (lo > 9e18 ? INT64_MAX : (int64_t)lo);
        ramp_hi = hi > 9e18 ? INT64_MAX : (hi < -9e18 ? INT64_MIN : (int64_t)hi);
        if (!(c == c)) { ramp_lo = INT64_MIN; ramp_hi = INT64_MAX; }
    }
    // steps q -> q+1 are taken for q in [max(ts, t0), nstep_end)
    const int64_t nstep_end = (a.t1 < a.T - 1) ? a.t1 : a.T - 1;
    const int64_t first = ts > a.t0 ? ts : a.t0;
    const double2* __restrict__ srcbase = reinterpret_cast<const double2*>(a.src) + kk;
    // source values of step q's three stages -> this thread's ring slot; always one commit group
    auto prefetch = [&](int slot, const double* cslot, int64_t q) {
        if (a.src && q >= first && q < nstep_end) {
#pragma unroll
            for (int s = 0; s < 3; ++s)
                cp_async16(&sring[slot][s][tid], srcbase + __double_as_longlong(cslot[s * SO_C + 3]) * a.src_pitch);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    for (int64_t c = 0; c < nchunks; ++c) {
        const int buf = (int)(c & 1);
        mbar_wait(&bars[buf], (uint32_t)((c >> 1) & 1));
        const double* __restrict__ sc = &scoef[buf][0];
        const int64_t base = a.t0 + c * SO_CH;
        const int cnt = (int)((a.t1 - base) < SO_CH ? (a.t1 - base) : SO_CH);
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < SO_AHEAD; ++j) prefetch(j, sc + j * 3 * SO_C, base + j);
        }
        for (int i = 0; i < cnt; i += SO_AHEAD) {
            // Steady state (almost every step of a full-trajectory run): the mode is running, the
            // ramp window is behind it, every row is stored and the look-ahead stays inside the
            // run -- the same step with none of the per-row decisions.
            const int64_t q0 = base + i;
            if (i + SO_AHEAD <= cnt && q0 > ts && q0 > ramp_hi && q0 >= first && a.out_every == 1 &&
                next_out == q0 && q0 + 2 * SO_AHEAD <= nstep_end && a.src) {
                // the coefficients of step j + 1 are read while step j computes (the asm statements
                // below are compiler barriers: without this every step would start with an exposed
                // shared-memory round trip)
                const double* __restrict__ cb = sc + i * 3 * SO_C;
                SoStage c0 = load_stage(cb), c1 = load_stage(cb + SO_C), c2 = load_stage(cb + 2 * SO_C);
#pragma unroll
                for (int j = 0; j < SO_AHEAD; ++j) {
                    const double* __restrict__ cq = cb + j * 3 * SO_C;
                    const SoStage n0 = load_stage(cq + 3 * SO_C), n1 = load_stage(cq + 4 * SO_C),
                                  n2 = load_stage(cq + 5 * SO_C);
                    __stcs(orow, y.yr0);
                    __stcs(orow + op, y.yr1);
                    __stcs(orow + 2 * op, y.yi0);
                    __stcs(orow + 3 * op, y.yi1);
                    orow += 4 * op;
                    asm volatile("cp.async.wait_group %0;" ::"n"(SO_AHEAD - 1) : "memory");
                    double2 s0 = sring[j][0][tid], s1 = sring[j][1][tid], s2 = sring[j][2][tid];
                    {
                        const double* __restrict__ ca = cq + SO_AHEAD * 3 * SO_C;
#pragma unroll
                        for (int s = 0; s < 3; ++s)
                            cp_async16(&sring[j][s][tid], srcbase + __double_as_longlong(ca[s * SO_C + 3]) * a.src_pitch);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                    if (c0.fotix < fo_ts) {               // see below
                        s0 = make_double2(nan, nan);
                        if (c1.fotix < fo_ts) s1 = s0;
                        if (c2.fotix < fo_ts) s2 = s0;
                    }
                    so_step(y, c0, c1, c2, k2, s0, s1, s2, h, hh, h6);
                    c0 = n0; c1 = n1; c2 = n2;
                }
                next_out += SO_AHEAD;
                continue;
            }
#pragma unroll
            for (int j = 0; j < SO_AHEAD; ++j) {
                const int64_t q = base + i + j;
                if (i + j >= cnt) break;
                const double* __restrict__ cq = sc + (i + j) * 3 * SO_C;
                if (q == ts) {                            // the start row is ystart[:, k] itself
                    y.yr0 = a.ystart[kk]; y.yr1 = a.ystart[nk + kk];
                    y.yi0 = a.ystart[2 * nk + kk]; y.yi1 = a.ystart[3 * nk + kk];
                }
                if (q == next_out) {                      // rows before the start stay NaN, rk4.py:100
                    const bool pre = q < ts;
                    __stcs(orow, pre ? nan : y.yr0);
                    __stcs(orow + op, pre ? nan : y.yr1);
                    __stcs(orow + 2 * op, pre ? nan : y.yi0);
                    __stcs(orow + 3 * op, pre ? nan : y.yi1);
                    next_out += a.out_every;
                    orow += 4 * op;
                }
                const bool stepping = q >= first && q < nstep_end;
                // this step's source values have landed when at most SO_AHEAD - 1 younger groups
                // are still pending
                asm volatile("cp.async.wait_group %0;" ::"n"(SO_AHEAD - 1) : "memory");
                double2 s0 = make_double2(0.0, 0.0), s1 = s0, s2 = s0;
                if (stepping && a.src) { s0 = sring[j][0][tid]; s1 = sring[j][1][tid]; s2 = sring[j][2][tid]; }
                // refill the slot with the source values of step q + SO_AHEAD (its coefficients are
                // inside this chunk's SO_CHL steps)
                prefetch(j, cq + SO_AHEAD * 3 * SO_C, q + SO_AHEAD);
                if (stepping) {
                    const SoStage c0 = load_stage(cq), c1 = load_stage(cq + SO_C), c2 = load_stage(cq + 2 * SO_C);
                    if (q >= ramp_lo && q <= ramp_hi) {
                        const double r0 = ramp_factor(cq, tstart, tsd, a.ra, a.rb, a.rc, a.rd, a.re);
                        const double r1 = ramp_factor(cq + SO_C, tstart, tsd, a.ra, a.rb, a.rc, a.rd, a.re);
                        const double r2 = ramp_factor(cq + 2 * SO_C, tstart, tsd, a.ra, a.rb, a.rc, a.rd, a.re);
                        if (r0 != 1.0) { s0.x = r0 * s0.x; s0.y = r0 * s0.y; }
                        if (r1 != 1.0) { s1.x = r1 * s1.x; s1.y = r1 * s1.y; }
                        if (r2 != 1.0) { s2.x = r2 * s2.x; s2.y = r2 * s2.y; }
                    }
                    // a stage that reads the first-order result before this mode's first-order
                    // start sees NaN there in the reference
                    if (c0.fotix < fo_ts) {
                        s0 = make_double2(nan, nan);
                        if (c1.fotix < fo_ts) s1 = s0;
                        if (c2.fotix < fo_ts) s2 = s0;
                    }
                    so_step(y, c0, c1, c2, k2, s0, s1, s2, h, hh, h6);
                }
            }
        }
        __syncthreads();                                  // everyone is done reading scoef[buf]
        if (tid == 0 && c + 2 < nchunks) issue(c + 2);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (live && a.state && ts < a.t1) {
        a.state[kk] = y.yr0; a.state[nk + kk] = y.yr1; a.state[2 * nk + kk] = y.yi0; a.state[3 * nk + kk] = y.yi1;
    }
}

// one right-hand side: dydx[4][nk] = derivs(y[4][nk], t) with the stage coefficients in a.coef
__global__ void so_derivs_kernel(SoArgs a, const double* __restrict__ y, const double* __restrict__ srcrow,
                                 double* __restrict__ dydx) {
    const int64_t kk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= a.nk) return;
    const int64_t nk = a.nk;
    const SoStage c = load_stage(a.coef);
    const double k2 = a.k[kk] * a.k[kk];
    double2 s = srcrow ? reinterpret_cast<const double2*>(srcrow)[kk] : make_double2(0.0, 0.0);
    if (a.ramp) {
        const double r = ramp_factor(a.coef, a.tstart[kk], (double)a.rtsix[kk], a.ra, a.rb, a.rc, a.rd, a.re);
        if (r != 1.0) { s.x = r * s.x; s.y = r * s.y; }
    }
    if (a.fo_tsix && c.fotix < a.fo_tsix[kk]) s = make_double2(qnan(), qnan());
    const double B = fma(k2, c.R2, c.c2);
    dydx[kk] = y[nk + kk];                                               // cosmomodels.py:751
    dydx[nk + kk] = so_accel(c.c1, B, y[kk], y[nk + kk], s.x);
    dydx[2 * nk + kk] = y[3 * nk + kk];                                  // :758
    dydx[3 * nk + kk] = so_accel(c.c1, B, y[2 * nk + kk], y[3 * nk + kk], s.y);
}

static int fill_ramp(SoArgs& a, const double* ramp_host, const double* tstart_dev, const char* who) {
    a.ramp = 0; a.ra = a.rb = a.rc = a.rd = a.re = 0.0; a.tstart = nullptr;
    if (!ramp_host) return PF_OK;
    if (!tstart_dev) { set_error("%s: a ramp needs tstart_dev", who); return PF_E_ARG; }
    a.ramp = 1;
    a.ra = ramp_host[0]; a.rb = ramp_host[1]; a.rc = ramp_host[2]; a.rd = ramp_host[3]; a.re = ramp_host[4];
    a.tstart = tstart_dev;
    return PF_OK;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_so_run(int64_t nk, const double* k_dev, const int64_t* tsix_dev, const int64_t* fo_tsix_dev,
                         const double* ystart_dev, const double* coef_dev, const double* source_dev,
                         int64_t source_rows, int64_t source_pitch, const double* tstart_dev,
                         const int64_t* ramp_tsix_dev, const double* ramp_host, int64_t T, double h,
                         int64_t t0, int64_t t1,
                         double* state_dev, double* yout_dev, int64_t out_every, int64_t out_pitch,
                         void* stream) {
    if (nk < 0 || T < 1 || t0 < 0 || t1 > T || t0 > t1 || out_every < 0) {
        set_error("pf_so_run: bad sizes (nk=%lld T=%lld window=[%lld,%lld) out_every=%lld)",
                  (long long)nk, (long long)T, (long long)t0, (long long)t1, (long long)out_every);
        return PF_E_ARG;
    }
    if (!(h == h) || h == 0.0) { set_error("pf_so_run: Need to specify h."); return PF_E_STEP; }
    if (nk == 0 || t0 == t1) return PF_OK;
    if (!k_dev || !tsix_dev || !ystart_dev || !coef_dev) { set_error("pf_so_run: null buffer"); return PF_E_ARG; }
    if (source_dev && (source_rows < 1 || source_pitch < nk)) {
        set_error("pf_so_run: source is [%lld][%lld] for %lld modes", (long long)source_rows,
                  (long long)source_pitch, (long long)nk);
        return PF_E_ARG;
    }
    if (t0 > 0 && !state_dev) { set_error("pf_so_run: a window with t0 > 0 needs state_dev"); return PF_E_ARG; }
    if (out_every > 0 && !yout_dev) { set_error("pf_so_run: out_every > 0 needs yout_dev"); return PF_E_ARG; }
    if (out_pitch != 0 && out_pitch < nk) { set_error("pf_so_run: out_pitch=%lld is smaller than nk=%lld", (long long)out_pitch, (long long)nk); return PF_E_ARG; }
    SoArgs a;
    a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.fo_tsix = fo_tsix_dev; a.ystart = ystart_dev;
    a.rtsix = ramp_tsix_dev ? ramp_tsix_dev : tsix_dev;
    a.coef = coef_dev; a.src = source_dev; a.src_pitch = source_pitch;
    int rc = fill_ramp(a, ramp_host, tstart_dev, "pf_so_run");
    if (rc) return rc;
    a.T = T; a.t0 = t0; a.t1 = t1; a.h = h; a.state = state_dev;
    a.yout = out_every > 0 ? yout_dev : nullptr; a.out_every = out_every;
    a.opitch = out_pitch ? out_pitch : nk;
    so_rk4_kernel<<<(unsigned)((nk + SO_BLOCK - 1) / SO_BLOCK), SO_BLOCK, 0, (cudaStream_t)stream>>>(a);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_so_run launch");
    return PF_OK;
}

extern "C" int pf_so_derivs(int64_t nk, const double* k_dev, const int64_t* tsix_dev,
                            const int64_t* fo_tsix_dev, const double* y_dev, const double* coef_dev,
                            const double* source_row_dev, const double* tstart_dev,
                            const double* ramp_host, double* dydx_dev, void* stream) {
    if (nk < 0) { set_error("pf_so_derivs: bad sizes"); return PF_E_ARG; }
    if (nk == 0) return PF_OK;
    if (!k_dev || !y_dev || !coef_dev || !dydx_dev) { set_error("pf_so_derivs: null buffer"); return PF_E_ARG; }
    SoArgs a{};
    a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.rtsix = tsix_dev; a.fo_tsix = fo_tsix_dev; a.coef = coef_dev;
    int rc = fill_ramp(a, ramp_host, tstart_dev, "pf_so_derivs");
    if (rc) return rc;
    if (a.ramp && !tsix_dev) { set_error("pf_so_derivs: a ramp needs tsix_dev"); return PF_E_ARG; }
    const int block = 128;
    so_derivs_kernel<<<(unsigned)((nk + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
        a, y_dev, source_row_dev, dydx_dev);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_so_derivs launch");
    return PF_OK;
}

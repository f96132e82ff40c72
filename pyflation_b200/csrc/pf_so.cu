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
// is what remains: one thread per (k, re|im), 16 B of source read and 16 B of result written per
// thread and step -- HBM-bound.  Consecutive lanes hold (re, im) of consecutive k: a warp reads
// one contiguous 256-byte piece of a source row and writes two 128-byte pieces of result rows.
// The source values of the next PF_SO_AHEAD steps are loaded into registers before they are
// needed (they do not depend on the state), which is what keeps enough bytes in flight.
#include "pf_common.cuh"

namespace pf {

constexpr int SO_BLOCK = 128;
constexpr int SO_AHEAD = 4;          // steps of source values held in registers ahead of the RK4
constexpr int SO_C = PF_SO_COEF;

struct SoArgs {
    int64_t nk;
    const double* __restrict__ k;
    const int64_t* __restrict__ tsix;      // second-order start rows
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

struct SoMode {                            // what a thread knows about its mode
    double k2, tstart, tsd;
    int64_t fo_ts;
};

// source value of one stage as the right-hand side sees it: ramped (cosmomodels.py:925-944) and
// NaN when the first-order result the reference would read there has not started yet
__device__ __forceinline__ double stage_source(const SoArgs& a, const SoMode& m, const double* __restrict__ c,
                                               double raw) {
    double s = raw;
    if (a.ramp) {
        const double arg = c[3] - m.tstart - a.rb;
        if (fabs(arg) < a.re) {
            double r = (tanh(a.ra * arg) + a.rc) / a.rd;
            if (m.tsd == c[5]) r = 0.0;                   // ramp[self.tstartindex == sotix] = 0
            s = r * s;
        }
    }
    if ((int64_t)c[4] < m.fo_ts) s = qnan();
    return s;
}

// -( (3-eps) y1 + k^2 R2 y0 + c2 y0 + S )                       cosmomodels.py:754-755, 761-762
__device__ __forceinline__ double so_accel(const double* __restrict__ c, double k2, double y0, double y1,
                                           double s) {
    double t = fma(c[2], y0, s);
    t = fma(k2 * c[1], y0, t);
    t = fma(c[0], y1, t);
    return -t;
}

__device__ __forceinline__ double load_src(const SoArgs& a, const double* __restrict__ c, int64_t kk, int part) {
    if (!a.src) return 0.0;
    const int64_t fotix = (int64_t)c[4];
    return __ldg(a.src + 2 * (fotix * a.src_pitch + kk) + part);
}

__global__ void __launch_bounds__(SO_BLOCK) so_rk4_kernel(SoArgs a) {
    const int tid = threadIdx.x;
    const int part = tid & 1;
    const int64_t kk = (int64_t)blockIdx.x * (SO_BLOCK / 2) + (tid >> 1);
    if (kk >= a.nk) return;
    const int64_t op = a.opitch;
    const int64_t ts = a.tsix[kk];
    SoMode m;
    m.k2 = a.k[kk] * a.k[kk];
    m.tstart = a.tstart ? a.tstart[kk] : 0.0;
    m.tsd = (double)ts;
    m.fo_ts = a.fo_tsix ? a.fo_tsix[kk] : INT64_MIN;
    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;     // rk4.py:31-32
    const double nan = qnan();

    int64_t next_out = (a.yout != nullptr && a.out_every > 0) ? a.t0 : INT64_MAX;
    double* orow = a.yout ? a.yout + (2 * part) * op + kk : nullptr;   // this thread's two rows
    auto advance_out = [&]() { next_out += a.out_every; orow += 4 * op; };

    // rows before this mode's start stay NaN                       rk4.py:100
    int64_t n = a.t0;
    const int64_t pre_end = ts < a.t1 ? ts : a.t1;
    if (next_out != INT64_MAX) {
        // first output row at or after n is next_out (= t0): walk the output rows only
        for (; next_out < pre_end; advance_out()) {
            __stcs(orow, nan);
            __stcs(orow + op, nan);
        }
    }
    if (ts >= a.t1) return;
    n = ts > a.t0 ? ts : a.t0;

    double y0, y1;
    if (n == ts) {                                        // the start row is ystart[:, k] itself
        y0 = a.ystart[(2 * part) * a.nk + kk];
        y1 = a.ystart[(2 * part + 1) * a.nk + kk];
    } else {                                              // resume a windowed run
        y0 = a.state[(2 * part) * a.nk + kk];
        y1 = a.state[(2 * part + 1) * a.nk + kk];
    }

    // steps n -> n+1 are taken for n in [n, nstep_end)
    const int64_t nstep_end = (a.t1 < a.T - 1) ? a.t1 : a.T - 1;
    double ring[SO_AHEAD][3];
#pragma unroll
    for (int j = 0; j < SO_AHEAD; ++j) {
        const int64_t q = n + j;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            ring[j][s] = (q < nstep_end) ? load_src(a, a.coef + (q * 3 + s) * SO_C, kk, part) : 0.0;
    }

    for (; n < a.t1; n += SO_AHEAD) {
#pragma unroll
        for (int j = 0; j < SO_AHEAD; ++j) {
            const int64_t q = n + j;
            if (q >= a.t1) break;
            if (q == next_out) {
                __stcs(orow, y0);
                __stcs(orow + op, y1);
                advance_out();
            }
            if (q < nstep_end) {
                const double* __restrict__ c = a.coef + q * 3 * SO_C;
                const double s0 = stage_source(a, m, c, ring[j][0]);
                const double s1 = stage_source(a, m, c + SO_C, ring[j][1]);
                const double s2 = stage_source(a, m, c + 2 * SO_C, ring[j][2]);
                // refill this slot with the values of step q + SO_AHEAD
                const int64_t qa = q + SO_AHEAD;
#pragma unroll
                for (int s = 0; s < 3; ++s)
                    ring[j][s] = (qa < nstep_end) ? load_src(a, a.coef + (qa * 3 + s) * SO_C, kk, part) : 0.0;
                // one classical RK4 step, operation order of rk4.py:27-55
                const double d0 = y1, d1 = so_accel(c, m.k2, y0, y1, s0);
                double u0 = fma(hh, d0, y0), u1 = fma(hh, d1, y1);
                const double e0 = u1, e1 = so_accel(c + SO_C, m.k2, u0, u1, s1);          // dyt
                u0 = fma(hh, e0, y0); u1 = fma(hh, e1, y1);
                double f0 = u1, f1 = so_accel(c + SO_C, m.k2, u0, u1, s1);                // dym
                u0 = fma(h, f0, y0); u1 = fma(h, f1, y1);
                f0 += e0; f1 += e1;
                const double g0 = u1, g1 = so_accel(c + 2 * SO_C, m.k2, u0, u1, s2);
                y0 = fma(h6, d0 + g0 + 2.0 * f0, y0);
                y1 = fma(h6, d1 + g1 + 2.0 * f1, y1);
            }
        }
    }
    if (a.state) {
        a.state[(2 * part) * a.nk + kk] = y0;
        a.state[(2 * part + 1) * a.nk + kk] = y1;
    }
}

// one right-hand side: dydx[4][nk] = derivs(y[4][nk], t) with the stage coefficients in c
__global__ void so_derivs_kernel(SoArgs a, const double* __restrict__ y, const double* __restrict__ srcrow,
                                 double* __restrict__ dydx) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int part = (int)(gid & 1);
    const int64_t kk = gid >> 1;
    if (kk >= a.nk) return;
    SoMode m;
    m.k2 = a.k[kk] * a.k[kk];
    m.tstart = a.tstart ? a.tstart[kk] : 0.0;
    m.tsd = a.tsix ? (double)a.tsix[kk] : -1.0;
    m.fo_ts = a.fo_tsix ? a.fo_tsix[kk] : INT64_MIN;
    const double* c = a.coef;
    const double raw = srcrow ? srcrow[2 * kk + part] : 0.0;
    const double s = stage_source(a, m, c, raw);
    const double y0 = y[(2 * part) * a.nk + kk], y1 = y[(2 * part + 1) * a.nk + kk];
    dydx[(2 * part) * a.nk + kk] = y1;                                   // cosmomodels.py:751, 758
    dydx[(2 * part + 1) * a.nk + kk] = so_accel(c, m.k2, y0, y1, s);
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
                         const double* ramp_host, int64_t T, double h, int64_t t0, int64_t t1,
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
    a.coef = coef_dev; a.src = source_dev; a.src_pitch = source_pitch;
    int rc = fill_ramp(a, ramp_host, tstart_dev, "pf_so_run");
    if (rc) return rc;
    a.T = T; a.t0 = t0; a.t1 = t1; a.h = h; a.state = state_dev;
    a.yout = out_every > 0 ? yout_dev : nullptr; a.out_every = out_every;
    a.opitch = out_pitch ? out_pitch : nk;
    const int kpb = SO_BLOCK / 2;
    so_rk4_kernel<<<(unsigned)((nk + kpb - 1) / kpb), SO_BLOCK, 0, (cudaStream_t)stream>>>(a);
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
    a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.fo_tsix = fo_tsix_dev; a.coef = coef_dev;
    int rc = fill_ramp(a, ramp_host, tstart_dev, "pf_so_derivs");
    if (rc) return rc;
    if (a.ramp && !tsix_dev) { set_error("pf_so_derivs: a ramp needs tsix_dev"); return PF_E_ARG; }
    const int block = 128;
    so_derivs_kernel<<<(unsigned)((2 * nk + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
        a, y_dev, source_row_dev, dydx_dev);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_so_derivs launch");
    return PF_OK;
}

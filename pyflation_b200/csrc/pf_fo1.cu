// K2 : single-field (nf = 1) first-order RK4 kernel, full-trajectory output -- the headline path
// (BASELINE configs 1, 2, 5).  Same mathematics and output as fo_rk4_kernel<1,2> in pf_fo.cu
// (rk4.py:27-138 + cosmomodels.py:625-675), restructured around what the profile showed
// (profiles/r1_k2_*): the kernel is bound by how fast one SM can retire store-carrying steps,
// so the per-step instruction count and the CTA shape matter more than raw occupancy.
//
//   * a thread owns KPT adjacent modes (4*KPT doubles of state); the 16-double step record is
//     read once per thread with eight 128-bit broadcast shared loads and reused for KPT modes;
//   * three row regimes per CTA, all CTA-uniform branches: rows before the CTA's first start
//     (pure NaN stores, no record traffic), the few rows where some modes of the CTA have
//     started and some have not (predicated path), and the steady state (no predicates);
//   * the CTA size is a launch parameter: the host picks it so that the grid is a whole number
//     of CTAs per SM (148 SMs) -- see pf_fo_run.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"

namespace pf {

constexpr int FO1_CH = 64;    // steps per record chunk (64 * 128 B = 8 KB per buffer)
constexpr int FO1_REC = 16;

struct Coef {                 // one step record, nf = 1
    double A[4], R[4], C[4], bg[3];
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

// One classical RK4 step of  x' = v, v' = -S,  S = A v + ((kR)^2 + C) x   for one complex mode.
__device__ __forceinline__ void rk4_mode(const Coef& c, double kval, double h, double hh, double h6,
                                         double (&x)[2], double (&v)[2]) {
    double B[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double kr = kval * c.R[s];
        B[s] = fma(kr, kr, c.C[s]);
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

template <int KPT>
__global__ void __launch_bounds__(512) fo1_kernel(FoArgs a) {
    __shared__ __align__(128) double srec[2][FO1_CH * FO1_REC];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ long long s_tsmin, s_tsmax;

    const int tid = threadIdx.x;
    const int64_t nk = a.nk;
    const int64_t kk0 = ((int64_t)blockIdx.x * blockDim.x + tid) * KPT;

    if (tid == 0) {
        s_tsmin = INT64_MAX;
        s_tsmax = INT64_MIN;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    double kval[KPT], x[KPT][2], v[KPT][2];
    int64_t ts[KPT];
    bool live[KPT];
#pragma unroll
    for (int m = 0; m < KPT; ++m) {
        live[m] = (kk0 + m) < nk;
        kval[m] = live[m] ? a.k[kk0 + m] : 0.0;
        ts[m] = live[m] ? a.tsix[kk0 + m] : INT64_MAX;
        x[m][0] = x[m][1] = v[m][0] = v[m][1] = 0.0;
        if (live[m]) {
            atomicMin(&s_tsmin, (long long)ts[m]);
            atomicMax(&s_tsmax, (long long)ts[m]);
            if (a.t0 > ts[m]) {                               // resume a windowed run
                const double2 sx = reinterpret_cast<const double2*>(a.state)[kk0 + m];
                const double2 sv = reinterpret_cast<const double2*>(a.state)[nk + kk0 + m];
                x[m][0] = sx.x; x[m][1] = sx.y; v[m][0] = sv.x; v[m][1] = sv.y;
            }
        }
    }
    __syncthreads();
    const int64_t tsmin = s_tsmin, tsmax = s_tsmax;
    if (tsmin == INT64_MAX) return;                           // CTA past the end of the batch

    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;         // rk4.py:31-32
    const double nan = qnan();
    double2* prow = reinterpret_cast<double2*>(a.yout) + kk0; // row t0, var 0, this thread's k
    const int64_t rstride = 5 * nk;

    // ---- regime 1: rows before any mode of this CTA starts: NaN+NaNj, no records needed
    int64_t n = a.t0;
    const int64_t nan_end = tsmin < a.t1 ? tsmin : a.t1;
    for (; n < nan_end; ++n) {
#pragma unroll
        for (int m = 0; m < KPT; ++m)
            if (live[m]) {
#pragma unroll
                for (int var = 0; var < 5; ++var) __stcs(prow + var * nk + m, make_double2(nan, nan));
            }
        prow += rstride;
    }
    if (n >= a.t1) return;

    // ---- regimes 2 and 3: records streamed through shared memory from row c0 on
    const int64_t c0 = n;
    const int64_t nchunks = (a.t1 - c0 + FO1_CH - 1) / FO1_CH;
    auto issue = [&](int64_t c) {
        const int buf = (int)(c & 1);
        const int64_t base = c0 + c * FO1_CH;
        const int64_t cnt = (a.t1 - base) < FO1_CH ? (a.t1 - base) : FO1_CH;
        const uint32_t bytes = (uint32_t)(cnt * FO1_REC * sizeof(double));
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
        for (int q = 0; q < cnt; ++q) {
            n = base + q;
            const Coef cf = load_record(&srec[buf][q * FO1_REC]);
            if (n > tsmax) {
                // ---- regime 3: every mode of the CTA is running
#pragma unroll
                for (int m = 0; m < KPT; ++m)
                    if (live[m]) {
                        __stcs(prow + m, make_double2(cf.bg[0], 0.0));
                        __stcs(prow + nk + m, make_double2(cf.bg[1], 0.0));
                        __stcs(prow + 2 * nk + m, make_double2(cf.bg[2], 0.0));
                        __stcs(prow + 3 * nk + m, make_double2(x[m][0], x[m][1]));
                        __stcs(prow + 4 * nk + m, make_double2(v[m][0], v[m][1]));
                    }
                if (n + 1 < a.T) {
#pragma unroll
                    for (int m = 0; m < KPT; ++m) rk4_mode(cf, kval[m], h, hh, h6, x[m], v[m]);
                }
            } else {
                // ---- regime 2: some modes of the CTA have started, some have not
#pragma unroll
                for (int m = 0; m < KPT; ++m) {
                    if (!live[m]) continue;
                    if (n < ts[m]) {
#pragma unroll
                        for (int var = 0; var < 5; ++var)
                            __stcs(prow + var * nk + m, make_double2(nan, nan));   // rk4.py:100
                        continue;
                    }
                    double2 b0 = make_double2(cf.bg[0], 0.0), b1 = make_double2(cf.bg[1], 0.0),
                            b2 = make_double2(cf.bg[2], 0.0);
                    if (n == ts[m]) {
                        // the start row is ystart[:, k] itself            rk4.py:106-107
                        const double2* ys = reinterpret_cast<const double2*>(a.ystart) + kk0 + m;
                        b0 = ys[0]; b1 = ys[nk]; b2 = ys[2 * nk];
                        const double2 sx = ys[3 * nk], sv = ys[4 * nk];
                        x[m][0] = sx.x; x[m][1] = sx.y; v[m][0] = sv.x; v[m][1] = sv.y;
                        if (a.flags) {
                            const bool bad = !(fabs(b0.x - cf.bg[0]) <= 1e-9 * fabs(cf.bg[0]) + 1e-300) ||
                                             !(fabs(b1.x - cf.bg[1]) <= 1e-9 * fabs(cf.bg[1]) + 1e-300) ||
                                             !(fabs(b2.x - cf.bg[2]) <= 1e-9 * fabs(cf.bg[2]) + 1e-300);
                            if (bad) atomicOr(a.flags, PF_FLAG_BG_MISMATCH);
                        }
                    }
                    __stcs(prow + m, b0);
                    __stcs(prow + nk + m, b1);
                    __stcs(prow + 2 * nk + m, b2);
                    __stcs(prow + 3 * nk + m, make_double2(x[m][0], x[m][1]));
                    __stcs(prow + 4 * nk + m, make_double2(v[m][0], v[m][1]));
                    if (n + 1 < a.T) rk4_mode(cf, kval[m], h, hh, h6, x[m], v[m]);
                }
            }
            prow += rstride;
        }
        __syncthreads();                              // everyone is done reading srec[buf]
        if (tid == 0 && c + 2 < nchunks) issue(c + 2);
    }

    if (a.state) {
#pragma unroll
        for (int m = 0; m < KPT; ++m)
            if (live[m] && ts[m] < a.t1) {
                reinterpret_cast<double2*>(a.state)[kk0 + m] = make_double2(x[m][0], x[m][1]);
                reinterpret_cast<double2*>(a.state)[nk + kk0 + m] = make_double2(v[m][0], v[m][1]);
            }
    }
}

// Grid shape: a whole number c of CTAs per SM, c the smallest for which a CTA holds at most
// `max_threads` threads (so every SM gets the same number of equally sized CTAs).
int launch_fo1(const FoArgs& a, cudaStream_t st) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    int kpt = 1, max_threads = 512;
    if (const char* e = getenv("PF_FO1_KPT")) kpt = atoi(e);
    if (const char* e = getenv("PF_FO1_MAXT")) max_threads = atoi(e);
    if (kpt != 2) kpt = 1;
    if (max_threads < 32 || max_threads > 512) max_threads = 512;
    int64_t threads_needed = (a.nk + kpt - 1) / kpt;
    int64_t per_sm = 1;
    int64_t block = 0;
    for (;; ++per_sm) {
        block = (threads_needed + sms * per_sm - 1) / (sms * per_sm);
        block = (block + 31) / 32 * 32;
        if (block <= max_threads) break;
    }
    if (const char* e = getenv("PF_FO1_BLOCK")) block = atoi(e);
    const int64_t grid = (threads_needed + block - 1) / block;
    if (kpt == 2) fo1_kernel<2><<<(unsigned)grid, (unsigned)block, 0, st>>>(a);
    else fo1_kernel<1><<<(unsigned)grid, (unsigned)block, 0, st>>>(a);
    return 0;
}

}  // namespace pf

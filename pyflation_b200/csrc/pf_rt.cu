// Runtime-nfields kernels: nflation with nfields > 8 (cmpotentials.py:791-846 accepts any nfields;
// it is the only potential of the reference defined for more than two fields).
//
// nf = 1..8 run the compiled instantiations (pf_bg.cu, pf_fo*.cu: state in registers, records
// streamed through shared memory by TMA).  Above that the register file cannot hold a column of
// the mode matrix, so these kernels take nf as an argument, keep the per-thread state in local
// memory and read the (L2-resident, factored) step records straight from global memory: slower
// per mode-step, same mathematics, same output layout.
//
//   bg_rt_kernel          K1   serial RK4 of (phi_i, phi'_i, H), NumPy's rounding (-fmad=false)
//   stage_table_rt_kernel K1b  factored records: A, R2, d, U, 1/H^2, 0, p[nf], g[nf] per stage
//   fo_rt_kernel          K3   thread <-> (k, j, re|im): x' = v, v' = -(A v + ((kR)^2 + d) x + p w + g z)
//   potential_rt / derivs_rt   K0 / K8 for API completeness
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"

namespace pf {

constexpr int RT = PF_MAX_NFIELDS_RT;

struct BgStateRt { double v[2 * RT + 1]; };

// nflation: U = sum 1/2 m^2 phi_i^2, dU_i = m^2 phi_i, d2U = m^2 I      cmpotentials.py:830-843
__device__ __forceinline__ void nflation_rt(double mass, int nf, const double* y, double& U, double* dU) {
    const double m2 = mass * mass;
    for (int i = 0; i < nf; ++i) dU[i] = m2 * y[2 * i];
    U = np_sum(nf, [&](int i) { return 0.5 * m2 * (y[2 * i] * y[2 * i]); });   // np.sum, cmpotentials.py:837
}

__device__ __forceinline__ void bg_rhs_rt(double mass, int nf, const double* y, double* dy, double* dU) {
    double U;
    nflation_rt(mass, nf, y, U, dU);
    const double H = y[2 * nf];
    const double H2 = H * H;
    for (int i = 0; i < nf; ++i) {
        const double pd = y[2 * i + 1];
        dy[2 * i] = pd;                                   // cosmomodels.py:563
        dy[2 * i + 1] = -(U * pd + dU[i]) / H2;           // :566
    }
    const double s = np_sum(nf, [&](int i) { return y[2 * i + 1] * y[2 * i + 1]; });
    dy[2 * nf] = -0.5 * s * H;                            // :569
}

__global__ void __launch_bounds__(32) bg_rt_kernel(double mass, int nf, int64_t T, double h, int64_t n0,
                                                   BgStateRt y0, double* __restrict__ bg,
                                                   double* __restrict__ stagey) {
    const int NB = 2 * nf + 1;
    if (threadIdx.x != 0) {
        const double nan = qnan();
        for (int64_t i = threadIdx.x - 1; i < n0 * NB; i += 31) bg[i] = nan;
        if (stagey)
            for (int64_t i = threadIdx.x - 1; i < n0 * NB * 4; i += 31) stagey[i] = nan;
        return;
    }
    const double hh = h * 0.5, h6 = h / 6.0;              // rk4.py:31-32
    double y[2 * RT + 1], k1[2 * RT + 1], k2[2 * RT + 1], k3[2 * RT + 1], k4[2 * RT + 1], yt[2 * RT + 1], dU[RT];
    for (int c = 0; c < NB; ++c) y[c] = y0.v[c];
    for (int64_t n = n0; n < T; ++n) {
        double* srow = stagey ? stagey + n * (4 * NB) : nullptr;
        for (int c = 0; c < NB; ++c) bg[n * NB + c] = y[c];
        if (srow) for (int c = 0; c < NB; ++c) srow[c] = y[c];
        bg_rhs_rt(mass, nf, y, k1, dU);
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + hh * k1[c];
        if (srow) for (int c = 0; c < NB; ++c) srow[NB + c] = yt[c];
        bg_rhs_rt(mass, nf, yt, k2, dU);
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + hh * k2[c];
        if (srow) for (int c = 0; c < NB; ++c) srow[2 * NB + c] = yt[c];
        bg_rhs_rt(mass, nf, yt, k3, dU);
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + h * k3[c];
        if (srow) for (int c = 0; c < NB; ++c) srow[3 * NB + c] = yt[c];
        bg_rhs_rt(mass, nf, yt, k4, dU);
        for (int c = 0; c < NB; ++c)                       // rk4.py:46,53
            y[c] = y[c] + h6 * (k1[c] + k4[c] + 2.0 * (k3[c] + k2[c]));
    }
}

// thread <-> (step n, stage s); record layout: rec_of / rec_stage_width / rec_bg_offset (pf_common.cuh)
__global__ void stage_table_rt_kernel(double mass, int nf, double ainit, int64_t T, double simtstart,
                                      double h, int64_t n0, const double* __restrict__ stagey,
                                      double* __restrict__ rec) {
    const int NB = 2 * nf + 1, SW = rec_stage_width(nf), REC = rec_of(nf), BG = 4 * SW;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = gid >> 2;
    const int s = (int)(gid & 3);
    if (n >= T) return;
    double* r = rec + n * REC;
    double* o = r + s * SW;
    if (n < n0) {
        const double nan = qnan();
        if (s == 0)
            for (int c = 0; c < NB; ++c) r[BG + c] = nan;
        for (int c = 0; c < SW; ++c) o[c] = nan;
        if (s == 3 && (NB + 4 * SW) < REC) r[REC - 1] = 0.0;
        return;
    }
    const double* y = stagey + (n * 4 + s) * NB;
    const double x = simtstart + (double)n * h;           // rk4.py:118 last_x
    const double t = (s == 0) ? x : ((s == 3) ? x + h : x + h * 0.5);
    const double a = ainit * exp(t);                      // cosmomodels.py:637
    const double m2 = mass * mass;
    const double U = np_sum(nf, [&](int i) { return 0.5 * m2 * (y[2 * i] * y[2 * i]); });
    const double H = y[2 * nf];
    const double invH2 = 1.0 / (H * H);
    o[0] = U * invH2;                                     // A
    const double rr = 1.0 / (a * H);
    o[1] = rr * rr;                                       // R2 = 1/(aH)^2
    o[2] = m2 * invH2;                                    // d
    o[3] = U;
    o[4] = invH2;
    o[5] = 0.0;
    for (int i = 0; i < nf; ++i) { o[6 + i] = y[2 * i + 1]; o[6 + nf + i] = m2 * y[2 * i]; }
    if (s == 0)
        for (int c = 0; c < NB; ++c) r[BG + c] = y[c];
    if (s == 3 && (NB + 4 * SW) < REC) r[REC - 1] = 0.0;
}

// One stage: kv_i from the stage argument (xin, vin); see mode_rhs / FactStage in pf_fo_generic.cuh.
struct StageRt {
    double A, Bd, z, w;
    const double* p;
    const double* g;
    __device__ __forceinline__ void prepare(const double* __restrict__ cs, int nf, double kval, const double* xin) {
        A = __ldg(cs);
        Bd = fma(kval, __ldg(cs + 1), __ldg(cs + 2));   // kval = k^2, record holds 1/(aH)^2
        const double U = __ldg(cs + 3), invH2 = __ldg(cs + 4);
        p = cs + 6;
        g = cs + 6 + nf;
        double sp0 = 0.0, sp1 = 0.0, sg0 = 0.0, sg1 = 0.0;
        int l = 0;
        for (; l + 1 < nf; l += 2) {
            sp0 = fma(__ldg(p + l), xin[l], sp0);
            sp1 = fma(__ldg(p + l + 1), xin[l + 1], sp1);
            sg0 = fma(__ldg(g + l), xin[l], sg0);
            sg1 = fma(__ldg(g + l + 1), xin[l + 1], sg1);
        }
        if (l < nf) { sp0 = fma(__ldg(p + l), xin[l], sp0); sg0 = fma(__ldg(g + l), xin[l], sg0); }
        z = (sp0 + sp1) * invH2;
        w = fma(U, z, (sg0 + sg1) * invH2);
    }
    __device__ __forceinline__ double kv(int i, double xi, double vi) const {
        double t = fma(Bd, xi, A * vi);
        t = fma(__ldg(p + i), w, t);
        t = fma(__ldg(g + i), z, t);
        return -t;
    }
};

// Mode kernel, runtime nf.  Row logic identical to fo_consume<NF, 1> (pf_fo_generic.cuh).
// MAXF (16 / 32 / 64) sizes the per-thread local arrays; with the CTA size chosen by the launcher
// (128 / 64 / 32 threads) a CTA's 6 * MAXF * 8 bytes per thread stay at 96 KB, i.e. inside L1.
template <int MAXF>
__global__ void __launch_bounds__(128) fo_rt_kernel(FoArgs a, int nf) {
    const int NB = 2 * nf + 1, SW = rec_stage_width(nf), REC = rec_of(nf), BG_OFF = 4 * SW;
    const bool pert = a.layout == PF_LAYOUT_PERT;
    const int NV = pert ? 2 * nf * nf : nvar_of(nf);
    const int RO = pert ? -NB : 0;
    const int tid = threadIdx.x;
    const int part = tid & 1, kl = tid >> 1;
    const int j = blockIdx.y;
    const int64_t kk = (int64_t)blockIdx.x * (blockDim.x >> 1) + kl;
    if (kk >= a.nk) return;
    const int64_t nk = a.nk, op = a.opitch;
    const double kval = a.k[kk] * a.k[kk];
    const int64_t ts = a.tsix[kk];
    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;     // rk4.py:31-32
    const double nan = qnan();

    double x[MAXF], v[MAXF], ax[MAXF], av[MAXF], xt[MAXF], vt[MAXF];
    for (int i = 0; i < nf; ++i) { x[i] = 0.0; v[i] = 0.0; }
    if (a.t0 > ts) {                                      // resume a windowed run
        for (int i = 0; i < nf; ++i) {
            x[i] = a.state[2 * ((int64_t)(2 * (i * nf + j)) * nk + kk) + part];
            v[i] = a.state[2 * ((int64_t)(2 * (i * nf + j) + 1) * nk + kk) + part];
        }
    }
    int64_t next_out = (a.yout != nullptr && a.out_every > 0) ? a.t0 : INT64_MAX;
    double* orow = a.yout;
    auto put = [&](int64_t row, double val) { __stcs(orow + 2 * (row * op + kk) + part, val); };

    for (int64_t n = a.t0; n < a.t1; ++n) {
        const double* __restrict__ r = a.rec + n * REC;
        const bool out = (n == next_out);
        if (n < ts) {
            if (out) {                                    // rows before the start stay NaN+NaNj
                if (!pert) {
                    put(2 * j, nan);
                    put(2 * j + 1, nan);
                    if (j == 0) put(2 * nf, nan);
                }
                for (int i = 0; i < nf; ++i) {
                    const int64_t row = RO + NB + 2 * (i * nf + j);
                    put(row, nan);
                    put(row + 1, nan);
                }
            }
        } else {
            double b0, b1, b2;                            // phi_j, phi'_j, H as this thread stores them
            if (n == ts) {                                // the start row is ystart[:, k] itself
                const double* ys = a.ystart;
                for (int i = 0; i < nf; ++i) {
                    const int64_t row = NB + 2 * (i * nf + j);
                    x[i] = ys[2 * (row * nk + kk) + part];
                    v[i] = ys[2 * ((row + 1) * nk + kk) + part];
                }
                b0 = ys[2 * ((int64_t)(2 * j) * nk + kk) + part];
                b1 = ys[2 * ((int64_t)(2 * j + 1) * nk + kk) + part];
                b2 = ys[2 * ((int64_t)(2 * nf) * nk + kk) + part];
                if (part == 0 && a.flags) {
                    const double ref[3] = {__ldg(r + BG_OFF + 2 * j), __ldg(r + BG_OFF + 2 * j + 1), __ldg(r + BG_OFF + 2 * nf)};
                    const double got[3] = {b0, b1, b2};
                    for (int q = 0; q < 3; ++q)
                        if (!(fabs(got[q] - ref[q]) <= 1e-9 * fabs(ref[q]) + 1e-300))
                            atomicOr(a.flags, PF_FLAG_BG_MISMATCH);
                }
            } else {
                b0 = part == 0 ? __ldg(r + BG_OFF + 2 * j) : 0.0;
                b1 = part == 0 ? __ldg(r + BG_OFF + 2 * j + 1) : 0.0;
                b2 = part == 0 ? __ldg(r + BG_OFF + 2 * nf) : 0.0;
            }
            if (out) {
                if (!pert) {
                    put(2 * j, b0);
                    put(2 * j + 1, b1);
                    if (j == 0) put(2 * nf, b2);
                }
                for (int i = 0; i < nf; ++i) {
                    const int64_t row = RO + NB + 2 * (i * nf + j);
                    put(row, x[i]);
                    put(row + 1, v[i]);
                }
            }
            if (n + 1 < a.T) {                            // one classical RK4 step, rk4.py:27-55
                StageRt st;
                st.prepare(r, nf, kval, x);
                for (int i = 0; i < nf; ++i) {
                    const double k1 = st.kv(i, x[i], v[i]);
                    av[i] = k1;
                    xt[i] = fma(hh, v[i], x[i]);
                    vt[i] = fma(hh, k1, v[i]);
                }
                st.prepare(r + SW, nf, kval, xt);
                for (int i = 0; i < nf; ++i) {
                    const double k2 = st.kv(i, xt[i], vt[i]);
                    ax[i] = fma(2.0, vt[i], v[i]);
                    av[i] = fma(2.0, k2, av[i]);
                    xt[i] = fma(hh, vt[i], x[i]);
                    vt[i] = fma(hh, k2, v[i]);
                }
                st.prepare(r + 2 * SW, nf, kval, xt);
                for (int i = 0; i < nf; ++i) {
                    const double k3 = st.kv(i, xt[i], vt[i]);
                    ax[i] = fma(2.0, vt[i], ax[i]);
                    av[i] = fma(2.0, k3, av[i]);
                    xt[i] = fma(h, vt[i], x[i]);
                    vt[i] = fma(h, k3, v[i]);
                }
                st.prepare(r + 3 * SW, nf, kval, xt);
                for (int i = 0; i < nf; ++i) {
                    const double k4 = st.kv(i, xt[i], vt[i]);
                    x[i] = fma(h6, ax[i] + vt[i], x[i]);
                    v[i] = fma(h6, av[i] + k4, v[i]);
                }
            }
        }
        if (out) {
            next_out += a.out_every;
            orow += 2 * (int64_t)NV * op;
        }
    }
    if (a.state && ts < a.t1) {
        for (int i = 0; i < nf; ++i) {
            a.state[2 * ((int64_t)(2 * (i * nf + j)) * nk + kk) + part] = x[i];
            a.state[2 * ((int64_t)(2 * (i * nf + j) + 1) * nk + kk) + part] = v[i];
        }
    }
}

__global__ void potential_rt_kernel(double mass, int nf, int64_t n, const double* __restrict__ y,
                                    double* __restrict__ U, double* __restrict__ dU, double* __restrict__ d2U) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int NB = 2 * nf + 1;
    double g[RT], u;
    nflation_rt(mass, nf, y + i * NB, u, g);
    U[i] = u;
    for (int f = 0; f < nf; ++f) dU[i * nf + f] = g[f];
    if (d2U) {
        const double m2 = mass * mass;
        for (int f = 0; f < nf; ++f)
            for (int q = 0; q < nf; ++q) d2U[(i * nf + f) * nf + q] = (f == q) ? m2 : 0.0;
    }
}

// K8, runtime nf: complex arithmetic throughout, potential on the real part of column 0
// (CanonicalFirstOrder.derivs cosmomodels.py:625-675, CanonicalBackground.derivs :551-571).
struct cx { double re, im; };
__device__ __forceinline__ cx cxmul(cx a, cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cx cxadd(cx a, cx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cx cxs(double s, cx a) { return {s * a.re, s * a.im}; }
__device__ __forceinline__ cx cxinv(cx a) { const double d = a.re * a.re + a.im * a.im; return {a.re / d, -a.im / d}; }

__global__ void derivs_rt_kernel(double mass, int nf, double ainit, int first_order, int64_t ncol, double t,
                                 const cx* __restrict__ y, const double* __restrict__ k, cx* __restrict__ dydx) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    const int NB = 2 * nf + 1;
    const double m2 = mass * mass;
    const double U = np_sum(nf, [&](int i) {
        const double phi = y[(int64_t)(2 * i) * ncol].re;                       // column 0
        return 0.5 * m2 * (phi * phi);
    });
    const cx H = y[(int64_t)(2 * nf) * ncol + c];
    const cx invH2 = cxinv(cxmul(H, H));
    cx s = {0.0, 0.0};
    for (int i = 0; i < nf; ++i) {
        const cx pd = y[(int64_t)(2 * i + 1) * ncol + c];
        const double dUi = m2 * y[(int64_t)(2 * i) * ncol].re;
        dydx[(int64_t)(2 * i) * ncol + c] = pd;
        cx num = cxs(U, pd);
        num.re += dUi;
        const cx q = cxmul(num, invH2);
        dydx[(int64_t)(2 * i + 1) * ncol + c] = {-q.re, -q.im};
        s = cxadd(s, cxmul(pd, pd));
    }
    dydx[(int64_t)(2 * nf) * ncol + c] = cxmul(cxs(-0.5, s), H);
    if (!first_order) return;
    const cx kr = cxs(k[c], cxinv(cxs(ainit * exp(t), H)));
    const cx kr2 = cxmul(kr, kr);
    for (int i = 0; i < nf; ++i) {
        const cx pdi = y[(int64_t)(2 * i + 1) * ncol + c];
        const double dUi = m2 * y[(int64_t)(2 * i) * ncol].re;
        for (int j = 0; j < nf; ++j) {
            const int64_t row = NB + 2 * (i * nf + j);
            const cx dp = y[row * ncol + c], dpd = y[(row + 1) * ncol + c];
            cx inner = {0.0, 0.0};
            for (int l = 0; l < nf; ++l) {
                const cx pdl = y[(int64_t)(2 * l + 1) * ncol + c];
                const double dUl = m2 * y[(int64_t)(2 * l) * ncol].re;
                cx M = cxadd(cxs(dUl, pdi), cxs(dUi, pdl));
                M = cxadd(M, cxs(U, cxmul(pdi, pdl)));
                if (i == l) M.re += m2;
                inner = cxadd(inner, cxmul(M, y[(int64_t)(NB + 2 * (l * nf + j)) * ncol + c]));
            }
            cx tot = cxadd(cxmul(cxs(U, dpd), invH2), cxmul(kr2, dp));
            tot = cxadd(tot, cxmul(inner, invH2));
            dydx[row * ncol + c] = dpd;
            dydx[(row + 1) * ncol + c] = {-tot.re, -tot.im};
        }
    }
}

// ---- host-side launchers used by the C entry points when nfields > PF_MAX_NFIELDS ----
int rt_bg_run(const pf_model* m, int64_t T, double h, int64_t n0, const double* y0_host, double* bg,
              double* stagey, cudaStream_t st) {
    BgStateRt y0;
    for (int c = 0; c < 2 * RT + 1; ++c) y0.v[c] = c < 2 * m->nfields + 1 ? y0_host[c] : 0.0;
    bg_rt_kernel<<<1, 32, 0, st>>>(m->params[0], m->nfields, T, h, n0, y0, bg, stagey);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_bg_run launch (runtime nf)");
    return PF_OK;
}

int rt_stage_table(const pf_model* m, int64_t T, double simtstart, double h, int64_t n0,
                   const double* stagey, double* rec, cudaStream_t st) {
    const int block = 128;
    const int64_t grid = (T * 4 + block - 1) / block;
    stage_table_rt_kernel<<<(unsigned)grid, block, 0, st>>>(m->params[0], m->nfields, m->ainit, T, simtstart, h,
                                                            n0, stagey, rec);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_stage_table launch (runtime nf)");
    return PF_OK;
}

int rt_fo_run(const pf_model* m, const FoArgs& a, cudaStream_t st) {
    const int nf = m->nfields;
    // 9..16 fields: register-resident factored instantiations (pf_fo_wide.cu); PF_FO_RT forces the
    // local-memory kernel below (A/B measurements, and the path every nf > 16 takes)
    const bool rt_only = getenv("PF_FO_RT") != nullptr;   // read per call: the tests flip it
    if (!rt_only && wide_fo_run(nf, a, st)) {
        PF_CUDA_CHECK(cudaGetLastError(), "pf_fo_run launch (wide nf)");
        return PF_OK;
    }
    const int block = nf <= 16 ? 128 : (nf <= 32 ? 64 : 32);
    dim3 grid((unsigned)((a.nk + block / 2 - 1) / (block / 2)), (unsigned)nf, 1);
    static bool carved = false;
    if (!carved) {                                        // no shared memory used: all of it to L1
        cudaFuncSetAttribute(fo_rt_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(fo_rt_kernel<32>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(fo_rt_kernel<64>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        carved = true;
    }
    if (nf <= 16) fo_rt_kernel<16><<<grid, block, 0, st>>>(a, nf);
    else if (nf <= 32) fo_rt_kernel<32><<<grid, block, 0, st>>>(a, nf);
    else fo_rt_kernel<64><<<grid, block, 0, st>>>(a, nf);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_fo_run launch (runtime nf)");
    return PF_OK;
}

int rt_potential_eval(const pf_model* m, int64_t n, const double* y, double* U, double* dU, double* d2U,
                      cudaStream_t st) {
    const int block = 128;
    potential_rt_kernel<<<(unsigned)((n + block - 1) / block), block, 0, st>>>(m->params[0], m->nfields, n, y, U,
                                                                                 dU, d2U);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_potential_eval launch (runtime nf)");
    return PF_OK;
}

int rt_derivs(const pf_model* m, int first_order, int64_t ncol, const double* y, double t, const double* k,
              double* dydx, cudaStream_t st) {
    const int block = 64;
    derivs_rt_kernel<<<(unsigned)((ncol + block - 1) / block), block, 0, st>>>(
        m->params[0], m->nfields, m->ainit, first_order, ncol, t, (const cx*)y, k, (cx*)dydx);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_derivs launch (runtime nf)");
    return PF_OK;
}

}  // namespace pf

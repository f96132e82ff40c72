// K1  : serial fp64 RK4 of the shared background (phi_i, phi'_i, H)   -- pf_bg_run
// K1b : per-step coefficient records for the perturbation kernel        -- pf_stage_table
// K0  : potential evaluation at n points                                -- pf_potential_eval
//
// Reference semantics: rk4.py:27-138 driving CanonicalBackground.derivs (cosmomodels.py:551-571);
// the stage coefficients are the k-independent factors of CanonicalFirstOrder.derivs
// (cosmomodels.py:637-674) evaluated at the RK4 stage arguments of the background.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"
#include "pf_potentials.cuh"
#include "pf_records.cuh"

namespace pf {

struct BgState { double v[2 * PF_MAX_NFIELDS + 1]; };

// One thread integrates; the rest of the warp NaN-fills rows before n0 and leaves.
template <int POT, int NF>
__global__ void __launch_bounds__(32) bg_rk4_kernel(PotParams par, int64_t T, double h, int64_t n0,
                                                    BgState y0, double* __restrict__ bg,
                                                    double* __restrict__ stagey) {
    constexpr int NB = 2 * NF + 1;
    if (threadIdx.x != 0) {
        const double nan = qnan();
        for (int64_t i = threadIdx.x - 1; i < n0 * NB; i += 31) bg[i] = nan;
        if (stagey)
            for (int64_t i = threadIdx.x - 1; i < n0 * NB * 4; i += 31) stagey[i] = nan;
        return;
    }
    const double hh = h * 0.5, h6 = h / 6.0;              // rk4.py:31-32
    double y[NB], k1[NB], k2[NB], k3[NB], k4[NB], yt[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) y[c] = y0.v[c];
    for (int64_t n = n0; n < T; ++n) {
        double* srow = stagey ? stagey + n * (4 * NB) : nullptr;
#pragma unroll
        for (int c = 0; c < NB; ++c) bg[n * NB + c] = y[c];
        if (srow) {
#pragma unroll
            for (int c = 0; c < NB; ++c) srow[c] = y[c];
        }
        bg_rhs<POT, NF>(par.v, y, k1);
#pragma unroll
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + hh * k1[c];
        if (srow) {
#pragma unroll
            for (int c = 0; c < NB; ++c) srow[NB + c] = yt[c];
        }
        bg_rhs<POT, NF>(par.v, yt, k2);
#pragma unroll
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + hh * k2[c];
        if (srow) {
#pragma unroll
            for (int c = 0; c < NB; ++c) srow[2 * NB + c] = yt[c];
        }
        bg_rhs<POT, NF>(par.v, yt, k3);
#pragma unroll
        for (int c = 0; c < NB; ++c) yt[c] = y[c] + h * k3[c];
        if (srow) {
#pragma unroll
            for (int c = 0; c < NB; ++c) srow[3 * NB + c] = yt[c];
        }
        bg_rhs<POT, NF>(par.v, yt, k4);
#pragma unroll
        for (int c = 0; c < NB; ++c)                       // rk4.py:46,53
            y[c] = y[c] + h6 * (k1[c] + k4[c] + 2.0 * (k3[c] + k2[c]));
    }
}

// thread <-> (step n, stage s)
template <int POT, int NF>
__global__ void stage_table_kernel(PotParams par, double ainit, int64_t T, double simtstart,
                                   double h, int64_t n0, const double* __restrict__ stagey,
                                   double* __restrict__ rec, int factored) {
    constexpr int NB = 2 * NF + 1, REC = rec_of(NF), SW = 2 + NF * NF, BG = 4 * SW;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = gid >> 2;
    const int s = (int)(gid & 3);
    if (n >= T) return;
    double* r = rec + n * REC;
    if (n < n0) {
        const double nan = qnan();
        if (s == 0)
            for (int c = 0; c < NB; ++c) r[BG + c] = nan;
        for (int c = 0; c < SW; ++c) r[s * SW + c] = nan;
        if (s == 3 && (NB + 4 * SW) < REC) r[REC - 1] = 0.0;
        return;
    }
    double y[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) y[c] = stagey[(n * 4 + s) * NB + c];
    stage_record<POT, NF>(par.v, ainit, simtstart, h, n, s, y, r, factored != 0);
}

template <int POT, int NF>
__global__ void potential_eval_kernel(PotParams par, int64_t n, const double* __restrict__ y,
                                      double* __restrict__ U, double* __restrict__ dU,
                                      double* __restrict__ d2U) {
    constexpr int NB = 2 * NF + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double phi[NF], g[NF], hess[NF * NF], u;
#pragma unroll
    for (int f = 0; f < NF; ++f) phi[f] = y[i * NB + 2 * f];
    potential<POT, NF, true>(par.v, phi, u, g, hess);
    U[i] = u;
#pragma unroll
    for (int f = 0; f < NF; ++f) dU[i * NF + f] = g[f];
    if (d2U) {
#pragma unroll
        for (int f = 0; f < NF * NF; ++f) d2U[i * NF * NF + f] = hess[f];
    }
}

// K8: one evaluation of the right-hand side, dydx = derivs(y, t), for API completeness
// (CanonicalFirstOrder.derivs cosmomodels.py:625-675, CanonicalBackground.derivs :551-571).
// Complex arithmetic throughout, each column with its own background rows, the potential
// evaluated on the (real part of the) background of column 0 -- the reference's semantics.
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return {s * a.re, s * a.im}; }
__device__ __forceinline__ cplx cinv(cplx a) {
    const double d = a.re * a.re + a.im * a.im;
    return {a.re / d, -a.im / d};
}

template <int POT, int NF>
__global__ void derivs_kernel(PotParams par, double ainit, int first_order, int64_t ncol, double t,
                              const cplx* __restrict__ y, const double* __restrict__ k,
                              cplx* __restrict__ dydx) {
    constexpr int NB = 2 * NF + 1;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    double phi0[NF], dU[NF], d2U[NF * NF], U;
#pragma unroll
    for (int i = 0; i < NF; ++i) phi0[i] = y[(int64_t)(2 * i) * ncol].re;       // column 0
    potential<POT, NF, true>(par.v, phi0, U, dU, d2U);
    cplx pd[NF];
#pragma unroll
    for (int i = 0; i < NF; ++i) pd[i] = y[(int64_t)(2 * i + 1) * ncol + c];
    const cplx H = y[(int64_t)(2 * NF) * ncol + c];
    const cplx invH2 = cinv(cmul(H, H));
    cplx s = {0.0, 0.0};
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        dydx[(int64_t)(2 * i) * ncol + c] = pd[i];
        cplx num = cscale(U, pd[i]);
        num.re += dU[i];
        const cplx q = cmul(num, invH2);
        dydx[(int64_t)(2 * i + 1) * ncol + c] = {-q.re, -q.im};
        s = cadd(s, cmul(pd[i], pd[i]));
    }
    const cplx dH = cmul(cscale(-0.5, s), H);
    dydx[(int64_t)(2 * NF) * ncol + c] = dH;
    if (!first_order) return;
    const cplx aH = cscale(ainit * exp(t), H);
    const cplx kr = cscale(k[c], cinv(aH));
    const cplx kr2 = cmul(kr, kr);
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int j = 0; j < NF; ++j) {
            const int64_t row = NB + 2 * (i * NF + j);
            const cplx dp = y[row * ncol + c], dpd = y[(row + 1) * ncol + c];
            cplx inner = {0.0, 0.0};
#pragma unroll
            for (int l = 0; l < NF; ++l) {
                cplx M = cadd(cscale(dU[l], pd[i]), cscale(dU[i], pd[l]));
                M = cadd(M, cscale(U, cmul(pd[i], pd[l])));
                M.re += d2U[i * NF + l];
                inner = cadd(inner, cmul(M, y[(int64_t)(NB + 2 * (l * NF + j)) * ncol + c]));
            }
            cplx tot = cadd(cmul(cscale(U, dpd), invH2), cmul(kr2, dp));
            tot = cadd(tot, cmul(inner, invH2));
            dydx[row * ncol + c] = dpd;
            dydx[(row + 1) * ncol + c] = {-tot.re, -tot.im};
        }
}

static PotParams pack_params(const pf_model* m) {
    PotParams p;
    for (int i = 0; i < PF_MAX_PARAMS; ++i) p.v[i] = m->params[i];
    return p;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_bg_run(const pf_model* m, int64_t T, double h, int64_t n0, const double* y0_host,
                         double* bg_dev, double* stagey_dev, void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (!y0_host || !bg_dev || T < 1) { set_error("pf_bg_run: null buffer or T < 1"); return PF_E_ARG; }
    if (!(h == h) || h == 0.0) { set_error("pf_bg_run: Need to specify h."); return PF_E_STEP; }
    if (n0 < 0 || n0 > T) { set_error("pf_bg_run: Start times outside range of steps."); return PF_E_START; }
    if (m->nfields > PF_MAX_NFIELDS) return rt_bg_run(m, T, h, n0, y0_host, bg_dev, stagey_dev, (cudaStream_t)stream);
    BgState y0;
    for (int c = 0; c < 2 * PF_MAX_NFIELDS + 1; ++c) y0.v[c] = c < 2 * m->nfields + 1 ? y0_host[c] : 0.0;
    PotParams par = pack_params(m);
    cudaStream_t st = (cudaStream_t)stream;
#define PF_CALL(POT, NF) bg_rk4_kernel<POT, NF><<<1, 32, 0, st>>>(par, T, h, n0, y0, bg_dev, stagey_dev)
    bool ok = PF_DISPATCH_POT_NF(m->potential_id, m->nfields, PF_CALL);
#undef PF_CALL
    if (!ok) { set_error("pf_bg_run: potential %d is not defined for nfields=%d", m->potential_id, m->nfields); return PF_E_POTENTIAL; }
    PF_CUDA_CHECK(cudaGetLastError(), "pf_bg_run launch");
    return PF_OK;
}

extern "C" int pf_stage_table(const pf_model* m, int64_t T, double simtstart, double h, int64_t n0,
                              const double* bg_dev, const double* stagey_dev, double* rec_dev,
                              void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    (void)bg_dev;
    if (!stagey_dev || !rec_dev || T < 1) { set_error("pf_stage_table: null buffer or T < 1"); return PF_E_ARG; }
    if (!(h == h) || h == 0.0) { set_error("pf_stage_table: Need to specify h."); return PF_E_STEP; }
    if (n0 < 0 || n0 > T) { set_error("pf_stage_table: Start times outside range of steps."); return PF_E_START; }
    if (m->nfields > PF_MAX_NFIELDS) return rt_stage_table(m, T, simtstart, h, n0, stagey_dev, rec_dev, (cudaStream_t)stream);
    PotParams par = pack_params(m);
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 128;
    const int64_t grid = (T * 4 + block - 1) / block;
    const int factored = rec_kind(m) == PF_REC_FACTORED;
#define PF_CALL(POT, NF) \
    stage_table_kernel<POT, NF><<<(unsigned)grid, block, 0, st>>>(par, m->ainit, T, simtstart, h, n0, stagey_dev, rec_dev, factored)
    bool ok = PF_DISPATCH_POT_NF(m->potential_id, m->nfields, PF_CALL);
#undef PF_CALL
    if (!ok) { set_error("pf_stage_table: potential %d is not defined for nfields=%d", m->potential_id, m->nfields); return PF_E_POTENTIAL; }
    PF_CUDA_CHECK(cudaGetLastError(), "pf_stage_table launch");
    return PF_OK;
}

extern "C" int pf_potential_eval(const pf_model* m, int64_t n, const double* y_dev, double* U_dev,
                                 double* dU_dev, double* d2U_dev, void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!y_dev || !U_dev || !dU_dev))) { set_error("pf_potential_eval: null buffer"); return PF_E_ARG; }
    if (n == 0) return PF_OK;
    if (m->nfields > PF_MAX_NFIELDS) return rt_potential_eval(m, n, y_dev, U_dev, dU_dev, d2U_dev, (cudaStream_t)stream);
    PotParams par = pack_params(m);
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 128;
    const int64_t grid = (n + block - 1) / block;
#define PF_CALL(POT, NF) \
    potential_eval_kernel<POT, NF><<<(unsigned)grid, block, 0, st>>>(par, n, y_dev, U_dev, dU_dev, d2U_dev)
    bool ok = PF_DISPATCH_POT_NF(m->potential_id, m->nfields, PF_CALL);
#undef PF_CALL
    if (!ok) { set_error("pf_potential_eval: potential %d is not defined for nfields=%d", m->potential_id, m->nfields); return PF_E_POTENTIAL; }
    PF_CUDA_CHECK(cudaGetLastError(), "pf_potential_eval launch");
    return PF_OK;
}

extern "C" int pf_derivs(const pf_model* m, int first_order, int64_t ncol, const double* y_dev, double t,
                         const double* k_dev, double* dydx_dev, void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (ncol < 0) { set_error("pf_derivs: negative size"); return PF_E_ARG; }
    if (ncol == 0) return PF_OK;
    if (!y_dev || !dydx_dev || (first_order && !k_dev)) { set_error("pf_derivs: null buffer"); return PF_E_ARG; }
    if (m->nfields > PF_MAX_NFIELDS) return rt_derivs(m, first_order, ncol, y_dev, t, k_dev, dydx_dev, (cudaStream_t)stream);
    PotParams par = pack_params(m);
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 64;
    const int64_t grid = (ncol + block - 1) / block;
#define PF_CALL(POT, NF) \
    derivs_kernel<POT, NF><<<(unsigned)grid, block, 0, st>>>(par, m->ainit, first_order, ncol, t, (const cplx*)y_dev, k_dev, (cplx*)dydx_dev)
    bool ok = PF_DISPATCH_POT_NF(m->potential_id, m->nfields, PF_CALL);
#undef PF_CALL
    if (!ok) { set_error("pf_derivs: potential %d is not defined for nfields=%d", m->potential_id, m->nfields); return PF_E_POTENTIAL; }
    PF_CUDA_CHECK(cudaGetLastError(), "pf_derivs launch");
    return PF_OK;
}

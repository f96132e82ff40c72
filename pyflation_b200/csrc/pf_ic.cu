// K6 / K7 : the O(nk) host glue of FOCanonicalTwoStage.setfoics on the device, for batches
// where the reference's per-mode Python loops (cosmomodels.py:1074-1131) or even vectorised
// NumPy would dominate the wall time (2^20 .. 2^24 modes).
//   K6 pf_crossing_rows  : first row with k < cq*a*H            (findkcrossing, :1108)
//   K7 pf_bunch_davies   : start vectors foystart[:, k]          (getfoystart, :1443-1492;
//                          FONoPhase :2083-2132; FOSuppressOneField :2144-2197)
#include "pf_common.cuh"

namespace pf {

// cum[t] = running maximum of cq*a*H (non-decreasing), so "first t with cq*a*H > k" is an
// upper bound search; rows[k] = T when the mode never crosses.
__global__ void crossing_kernel(int64_t nk, const double* __restrict__ k, int64_t T,
                                const double* __restrict__ cum, int64_t* __restrict__ rows,
                                int32_t* __restrict__ nbad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const double kk = k[i];
    int64_t lo = 0, hi = T;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cum[mid] > kk) hi = mid; else lo = mid + 1;
    }
    rows[i] = lo;
    if ((lo >= T || lo == 0) && nbad) atomicAdd(nbad + (lo == 0 ? 1 : 0), 1);
}

__global__ void bunch_davies_kernel(int nf, int64_t nk, const double* __restrict__ k,
                                    const int64_t* __restrict__ rows, const double* __restrict__ bg,
                                    double simtstart, double h, double ainit, double etainit,
                                    int phase, int suppress_ix, double2* __restrict__ ystart) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const int nb = nbg_of(nf), nvar = nvar_of(nf);
    const int64_t row = rows[i];
    const double* y = bg + row * nb;
    const double kk = k[i];
    const double ts = simtstart + (double)row * h;        // tresult[row], rk4.py:134
    const double astar = ainit * exp(ts);
    const double Hstar = y[2 * nf];
    for (int f = 0; f < nf; ++f) {
        ystart[(int64_t)(2 * f) * nk + i] = make_double2(y[2 * f], 0.0);
        ystart[(int64_t)(2 * f + 1) * nk + i] = make_double2(y[2 * f + 1], 0.0);
    }
    // epsilon = 0.5 * np.sum(phidots**2, axis=1) (cosmomodels.py:521), NumPy's summation order
    const double eps = 0.5 * np_sum(nf, [&](int f) { return y[2 * f + 1] * y[2 * f + 1]; });
    ystart[(int64_t)(2 * nf) * nk + i] = make_double2(Hstar, 0.0);
    const double amp = 1.0 / (astar * sqrt(2.0 * kk));    // 1/(a sqrt(2k))
    double c = 1.0, s = 0.0;
    if (phase) {
        const double etastar = -1.0 / (astar * Hstar * (1.0 - eps));
        sincos(-(kk * (etastar - etainit)), &s, &c);      // exp(-i k (eta_* - eta_init))
    }
    const double w = kk / (astar * Hstar);
    const double2 dp = make_double2(amp * c, amp * s);
    // -(amp e^{-i keta}) (1 + i w)
    const double2 dpdot = make_double2(-(dp.x - dp.y * w), -(dp.y + dp.x * w));
    const double2 zero = make_double2(0.0, 0.0);
    for (int a = 0; a < nf; ++a)
        for (int b = 0; b < nf; ++b) {
            const bool diag = (a == b) && (a != suppress_ix);
            const int64_t r = nb + 2 * (a * nf + b);
            ystart[r * nk + i] = diag ? dp : zero;
            ystart[(r + 1) * nk + i] = diag ? dpdot : zero;
        }
    (void)nvar;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_crossing_rows(int64_t nk, const double* k_dev, int64_t T, const double* cum_dev,
                                int64_t* rows_dev, int32_t* nbad_dev, void* stream) {
    if (nk < 0 || T < 1) { set_error("pf_crossing_rows: bad sizes"); return PF_E_ARG; }
    if (nk == 0) return PF_OK;
    if (!k_dev || !cum_dev || !rows_dev) { set_error("pf_crossing_rows: null buffer"); return PF_E_ARG; }
    const int block = 256;
    crossing_kernel<<<(unsigned)((nk + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
        nk, k_dev, T, cum_dev, rows_dev, nbad_dev);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_crossing_rows launch");
    return PF_OK;
}

extern "C" int pf_bunch_davies(int nfields, int64_t nk, const double* k_dev, const int64_t* rows_dev,
                               const double* bg_dev, double simtstart, double h, double ainit,
                               double etainit, int phase, int suppress_ix, double* ystart_dev,
                               void* stream) {
    if (nfields < 1 || nfields > PF_MAX_NFIELDS_RT) { set_error("pf_bunch_davies: nfields=%d", nfields); return PF_E_NFIELDS; }
    if (nk < 0) { set_error("pf_bunch_davies: bad sizes"); return PF_E_ARG; }
    if (nk == 0) return PF_OK;
    if (!k_dev || !rows_dev || !bg_dev || !ystart_dev) { set_error("pf_bunch_davies: null buffer"); return PF_E_ARG; }
    const int block = 256;
    bunch_davies_kernel<<<(unsigned)((nk + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
        nfields, nk, k_dev, rows_dev, bg_dev, simtstart, h, ainit, etainit, phase, suppress_ix,
        (double2*)ystart_dev);
    PF_CUDA_CHECK(cudaGetLastError(), "pf_bunch_davies launch");
    return PF_OK;
}

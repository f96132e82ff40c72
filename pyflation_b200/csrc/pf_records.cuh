// Background right-hand side and step-record arithmetic shared by the stand-alone kernels
// (pf_bg.cu: K1, K1b) and the producer CTA of the fused launch (pf_fused.cu).
#pragma once
#include "pf_common.cuh"
#include "pf_potentials.cuh"

namespace pf {

template <int POT, int NF>
__device__ __forceinline__ void bg_rhs(const double* __restrict__ p, const double (&y)[2 * NF + 1],
                                       double (&dy)[2 * NF + 1]) {
    double phi[NF], dU[NF], d2U[NF * NF], U;
#pragma unroll
    for (int i = 0; i < NF; ++i) phi[i] = y[2 * i];
    potential<POT, NF, false>(p, phi, U, dU, d2U);
    const double H = y[2 * NF];
    const double H2 = H * H;
#pragma unroll
    for (int i = 0; i < NF; ++i) {
        const double pd = y[2 * i + 1];
        dy[2 * i] = pd;                                   // cosmomodels.py:563
        dy[2 * i + 1] = -(U * pd + dU[i]) / H2;           // :566 (true division, as NumPy's real arrays)
    }
    const double s = np_sum(NF, [&](int i) { return y[2 * i + 1] * y[2 * i + 1]; });   // np.sum, :569
    dy[2 * NF] = -0.5 * s * H;                            // :569
}

// One (step n, stage s) entry of a step record from the RK4 stage argument `y` of the background:
// A = U/H^2, R2 = 1/(aH)^2, C = M/H^2 (cosmomodels.py:637-674); stage 0 also stores the background
// row itself, stage 3 the padding word.  Layout: include/pyflation_b200.h (pf_stage_table).
//
// `factored` (PF_REC_FACTORED, nflation with nf >= 5 only): nflation's Hessian is m^2 * identity
// (cmpotentials.py:840), so  M = m^2 I + p g^T + g p^T + U p p^T  with p = phi', g = dU: a scalar
// diagonal plus a symmetric rank-2 term.  The stage slot then holds the factors instead of the
// dense matrix:  A, R2, d = m^2/H^2, U, 1/H^2, 0, p[NF], g[NF]  (6 + 2 NF <= 2 + NF^2 doubles), and the
// mode kernel forms  C x = d x + p (g.x + U p.x)/H^2 + g (p.x)/H^2  in 4 NF + 3 FMAs instead of NF^2.
template <int POT, int NF>
__device__ __forceinline__ void stage_record(const double* __restrict__ par, double ainit,
                                             double simtstart, double h, int64_t n, int s,
                                             const double (&y)[2 * NF + 1], double* __restrict__ r,
                                             bool factored = false) {
    constexpr int NB = 2 * NF + 1, REC = rec_of(NF), SW = 2 + NF * NF, BG = 4 * SW;
    const double x = simtstart + (double)n * h;           // rk4.py:118 last_x
    const double t = (s == 0) ? x : ((s == 3) ? x + h : x + h * 0.5);
    const double a = ainit * exp(t);                      // cosmomodels.py:637
    double phi[NF], pd[NF], dU[NF], d2U[NF * NF], U;
#pragma unroll
    for (int i = 0; i < NF; ++i) { phi[i] = y[2 * i]; pd[i] = y[2 * i + 1]; }
    potential<POT, NF, true>(par, phi, U, dU, d2U);
    const double H = y[2 * NF];
    const double invH2 = 1.0 / (H * H);
    double* o = r + s * SW;
    o[0] = U * invH2;                                     // A  (:673  U*y'/H**2)
    const double rr = 1.0 / (a * H);
    o[1] = rr * rr;                                       // R2 = 1/(aH)^2  (:673  (k/(a*H))**2)
    bool dense = true;
    if constexpr (POT == PF_NFLATION && NF >= 5) {
        if (factored) {
            dense = false;
            o[2] = d2U[0] * invH2;
            o[3] = U;
            o[4] = invH2;
            o[5] = 0.0;
#pragma unroll
            for (int i = 0; i < NF; ++i) { o[6 + i] = pd[i]; o[6 + NF + i] = dU[i]; }
#pragma unroll
            for (int i = 6 + 2 * NF; i < SW; ++i) o[i] = 0.0;
        }
    }
    if (dense) {
#pragma unroll
        for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int l = 0; l < NF; ++l)                  // :667-669 and /H**2 at :674
                o[2 + i * NF + l] =
                    (d2U[i * NF + l] + pd[i] * dU[l] + dU[i] * pd[l] + pd[i] * pd[l] * U) * invH2;
    }
    if (s == 0) {
#pragma unroll
        for (int c = 0; c < NB; ++c) r[BG + c] = y[c];
    }
    if (s == 3 && (NB + 4 * SW) < REC) r[REC - 1] = 0.0;
}

}  // namespace pf

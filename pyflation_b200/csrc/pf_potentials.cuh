// fp64 device versions of pyflation/cmpotentials.py (reference file:line given per case).
// Only (U, dU, d2U) exist here: the first-order path takes `potentials(...)[0:3]`
// (cosmomodels.py:641) and never touches d3U.
//
// phi[] holds the field values only (y[0], y[2], ... of the reference's background vector).
// p[] is pf_model.params in the order documented in include/pyflation_b200.h.
#pragma once
#include "../../include/pyflation_b200.h"

namespace pf {

// x^3 and x^4 rounded once (double-double product, then one rounding), like the correctly rounded
// pow() behind NumPy's  y[0]**3, y[0]**4  (cmpotentials.py:116-118 etc.); (x*x)*(x*x) rounds
// twice and is off by an ulp in a quarter of the cases, which the Bunch-Davies phase of the
// first-order start vectors magnifies by 1e8..1e11.
__device__ __forceinline__ double pow4_cr(double x) {
    const double p = x * x, e = fma(x, x, -p);          // x^2 = p + e exactly
    const double hi = p * p;
    const double lo = fma(p, p, -hi) + 2.0 * (p * e);
    return hi + lo;
}
__device__ __forceinline__ double pow3_cr(double x) {
    const double p = x * x, e = fma(x, x, -p);
    const double hi = p * x;
    const double lo = fma(p, x, -hi) + e * x;
    return hi + lo;
}

template <int POT>
struct PotNF {  // number of fields a potential is defined for (0 = any)
    static constexpr int value =
        (POT == PF_NFLATION) ? 0 : (POT >= PF_HYBRIDQUADRATIC ? 2 : 1);
};

// D2 = false skips the second derivatives (the background right-hand side needs U, dU only).
template <int POT, int NF, bool D2>
__device__ __forceinline__ void potential(const double* __restrict__ p, const double (&phi)[NF],
                                          double& U, double (&dU)[NF], double (&d2U)[NF * NF]) {
    if constexpr (D2) {
#pragma unroll
        for (int i = 0; i < NF * NF; ++i) d2U[i] = 0.0;
    }
    if constexpr (POT == PF_MSQPHISQ) {                       // cmpotentials.py:16-72
        const double m2 = p[0] * p[0];
        U = 0.5 * m2 * (phi[0] * phi[0]);
        dU[0] = m2 * phi[0];
        if constexpr (D2) d2U[0] = m2;
    } else if constexpr (POT == PF_LAMBDAPHI4) {              // :74-129
        const double l = p[0], f2 = phi[0] * phi[0];
        U = 0.25 * l * pow4_cr(phi[0]);
        dU[0] = l * pow3_cr(phi[0]);
        if constexpr (D2) d2U[0] = 3.0 * l * f2;
    } else if constexpr (POT == PF_LINDE) {                   // :131-199
        const double m2 = p[0] * p[0], l = p[1], f2 = phi[0] * phi[0];
        U = -0.5 * m2 * f2 + 0.25 * l * pow4_cr(phi[0]) + (m2 * m2) / (4.0 * l);
        dU[0] = -m2 * phi[0] + l * pow3_cr(phi[0]);
        if constexpr (D2) d2U[0] = -m2 + 3.0 * l * f2;
    } else if constexpr (POT == PF_HYBRID2AND4) {             // :201-269
        const double m2 = p[0] * p[0], l = p[1], f2 = phi[0] * phi[0];
        U = 0.5 * m2 * f2 + 0.25 * l * pow4_cr(phi[0]);
        dU[0] = m2 * phi[0] + l * pow3_cr(phi[0]);
        if constexpr (D2) d2U[0] = m2 + 3.0 * l * f2;
    } else if constexpr (POT == PF_PHI2OVER3) {               // :271-326
        const double s = p[0];
        U = s * pow(phi[0], 2.0 / 3);
        dU[0] = (2.0 / 3) * s * pow(phi[0], -1.0 / 3);
        if constexpr (D2) d2U[0] = -(2.0 / 9) * s * pow(phi[0], -4.0 / 3);
    } else if constexpr (POT == PF_MSQPHISQ_WITHV0) {         // :328-389
        const double m2 = p[0] * p[0];
        U = 0.5 * m2 * (phi[0] * phi[0]) + p[1];
        dU[0] = m2 * phi[0];
        if constexpr (D2) d2U[0] = m2;
    } else if constexpr (POT == PF_STEP_POTENTIAL) {          // :391-466 (note t-1 as coded)
        const double m2 = p[0] * p[0], c = p[1], d = p[2], phis = p[3];
        const double f = phi[0], f2 = f * f;
        const double arg = (f - phis) / d;
        const double s = 1.0 / cosh(arg), t = tanh(arg), s2 = s * s;
        const double w = 1.0 + c * (t - 1.0);
        U = 0.5 * m2 * f2 * w;
        dU[0] = m2 * f * w + c * m2 * f2 * s2 / (2.0 * d);
        if constexpr (D2)
            d2U[0] = 0.5 * m2 * (4.0 * c * f * s2 / d - 2.0 * c * f2 * s2 * t / (d * d) + 2.0 * w);
    } else if constexpr (POT == PF_BUMP_POTENTIAL || POT == PF_BUMP_NOTHIRDDERIV) {  // :468-543, :621-694
        const double m2 = p[0] * p[0], c = p[1], d = p[2], phib = p[3];
        const double f = phi[0], f2 = f * f;
        const double arg = (f - phib) / d;
        const double s = 1.0 / cosh(arg), t = tanh(arg);
        const double w = 1.0 + c * s;
        U = 0.5 * m2 * f2 * w;
        dU[0] = m2 * f * w - c * m2 * f2 * s * t / (2.0 * d);
        if constexpr (D2)
            d2U[0] = 0.5 * m2 * (-4.0 * c * f * s * t / d
                                 + c * f2 * (-(s * s * s) / (d * d) + s * (t * t) / (d * d))
                                 + 2.0 * w);
    } else if constexpr (POT == PF_RESONANCE) {               // :545-619 (d2U as coded at :615)
        const double m2 = p[0] * p[0], c = p[1], d = p[2];
        const double f = phi[0], f2 = f * f;
        double sp, cp;
        sincos(f / d, &sp, &cp);
        const double w = 1.0 + c * sp;
        U = 0.5 * m2 * f2 * w;
        dU[0] = m2 * f * w + c * m2 * f2 * cp / (2.0 * d);
        if constexpr (D2) d2U[0] = m2 * (w + 2.0 * c / d * cp * f);
    } else if constexpr (POT == PF_HYBRIDQUADRATIC) {         // :696-742
        const double a = p[0] * p[0], b = p[1] * p[1];
        U = 0.5 * (a * (phi[0] * phi[0]) + b * (phi[1] * phi[1]));
        dU[0] = a * phi[0];
        dU[1] = b * phi[1];
        if constexpr (D2) { d2U[0] = a; d2U[3] = b; }
    } else if constexpr (POT == PF_RIDGE_TWOFIELD) {          // :744-789
        const double g = p[0], m2 = p[1] * p[1], V0 = p[2];
        U = V0 - g * phi[0] - 0.5 * m2 * (phi[1] * phi[1]);
        dU[0] = -g;
        dU[1] = -m2 * phi[1];
        if constexpr (D2) d2U[3] = -m2;
    } else if constexpr (POT == PF_NFLATION) {                // :791-846
        const double m2 = p[0] * p[0];
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            dU[i] = m2 * phi[i];
            if constexpr (D2) d2U[i * NF + i] = m2;
        }
        U = np_sum(NF, [&](int i) { return 0.5 * m2 * (phi[i] * phi[i]); });    // np.sum, :837
    } else if constexpr (POT == PF_QUARTICTWOFIELD) {         // :848-894
        const double a = p[0] * p[0], b = p[1] * p[1], l1 = p[2], l2 = p[3];
        const double f2 = phi[0] * phi[0], c2 = phi[1] * phi[1];
        U = 0.5 * (a * f2 + 0.5 * l1 * pow4_cr(phi[0]) + b * c2 + 0.5 * l2 * pow4_cr(phi[1]));
        dU[0] = a * phi[0] + l1 * pow3_cr(phi[0]);
        dU[1] = b * phi[1] + l2 * pow3_cr(phi[1]);
        if constexpr (D2) { d2U[0] = a + 3.0 * l1 * f2; d2U[3] = b + 3.0 * l2 * c2; }
    } else if constexpr (POT == PF_HYBRIDQUARTIC) {           // :896-961
        const double l = p[0], v = p[1], mu = p[2], pc = p[3];
        const double f = phi[0], x = phi[1];
        const double l4 = (l * l) * (l * l), v2 = v * v, mu2 = mu * mu;
        const double pcv2 = (pc * v) * (pc * v);
        const double q = 1.0 - x * x / v2;
        U = l4 * (q * q + f * f / mu2 + 2.0 * ((f * x) * (f * x)) / pcv2);
        dU[0] = l4 * (2.0 * f / mu2 + 4.0 * f * (x * x) / pcv2);
        dU[1] = l4 * (-4.0 * x / v2 * q + 4.0 * (f * f) * x / pcv2);
        if constexpr (D2) {
            const double off = l4 * (8.0 * f * x / pcv2);
            d2U[0] = l4 * (2.0 / mu2 + 4.0 * (x * x) / pcv2);
            d2U[1] = off;
            d2U[2] = off;
            d2U[3] = l4 * (-4.0 / v2 * (1.0 - 3.0 * (x * x) / v2) + 4.0 * (f * f) / pcv2);
        }
    } else if constexpr (POT == PF_INFLECTION) {              // :963-1024 (V_0 as coded at :1007)
        const double l = p[0], g = p[1], r = p[2], m2 = p[3] * p[3];
        const double f = phi[0], x = phi[1], r3 = r * r * r, x2 = x * x;
        const double V0 = 0.75 * g * r + l / 24.0 * r3;
        U = V0 + 0.5 * m2 * (f * f) + g * x + 1.0 / 6.0 * l * (x2 * x)
            + (g / (4.0 * r3) + l / (8.0 * r)) * (x2 * x2);
        dU[0] = m2 * f;
        dU[1] = g + 0.5 * l * x2 + (g / (2.0 * r3) + l / (2.0 * r)) * (x2 * x);
        if constexpr (D2) { d2U[0] = m2; d2U[3] = l * x + 1.5 * (g / r3 + l / r) * x2; }
    } else if constexpr (POT == PF_HILLTOPAXION) {            // :1026-1086
        const double L = p[0], fa = p[1], m2 = p[2] * p[2];
        const double L4 = (L * L) * (L * L);
        const double w = 2.0 * 3.141592653589793 / fa;
        double sx, cx;
        sincos(w * phi[1], &sx, &cx);
        U = 0.5 * m2 * (phi[0] * phi[0]) + L4 * (1.0 - cx);
        dU[0] = m2 * phi[0];
        dU[1] = L4 * w * sx;
        if constexpr (D2) { d2U[0] = m2; d2U[3] = L4 * (w * w) * cx; }
    } else if constexpr (POT == PF_PRODUCTEXPONENTIAL) {      // :1088-1144 ((1-2 l chi) as coded)
        const double l = p[0], V0 = p[1];
        const double f = phi[0], x = phi[1];
        const double e = exp(-l * (x * x));
        U = V0 * (f * f) * e;
        dU[0] = 2.0 * V0 * f * e;
        dU[1] = -2.0 * l * x * V0 * (f * f) * e;
        if constexpr (D2) {
            const double off = -4.0 * l * x * V0 * f * e;
            d2U[0] = 2.0 * V0 * e;
            d2U[1] = off;
            d2U[2] = off;
            d2U[3] = -2.0 * l * V0 * (f * f) * e * (1.0 - 2.0 * l * x);
        }
    } else {
        U = 0.0;
    }
}

}  // namespace pf

// Calls `F<POT, NF>(args...)` for the runtime (pot, nf) pair; evaluates to false when the
// pair is not a compiled combination.  Single-field potentials need nf == 1, two-field ones
// nf == 2, nflation accepts 1..PF_MAX_NFIELDS.
#define PF_DISPATCH_POT_NF(pot, nf, CALL)                                            \
    [&]() -> bool {                                                                  \
        switch (pot) {                                                               \
            case PF_MSQPHISQ: if (nf == 1) { CALL(PF_MSQPHISQ, 1); return true; } break; \
            case PF_LAMBDAPHI4: if (nf == 1) { CALL(PF_LAMBDAPHI4, 1); return true; } break; \
            case PF_LINDE: if (nf == 1) { CALL(PF_LINDE, 1); return true; } break;   \
            case PF_HYBRID2AND4: if (nf == 1) { CALL(PF_HYBRID2AND4, 1); return true; } break; \
            case PF_PHI2OVER3: if (nf == 1) { CALL(PF_PHI2OVER3, 1); return true; } break; \
            case PF_MSQPHISQ_WITHV0: if (nf == 1) { CALL(PF_MSQPHISQ_WITHV0, 1); return true; } break; \
            case PF_STEP_POTENTIAL: if (nf == 1) { CALL(PF_STEP_POTENTIAL, 1); return true; } break; \
            case PF_BUMP_POTENTIAL: if (nf == 1) { CALL(PF_BUMP_POTENTIAL, 1); return true; } break; \
            case PF_RESONANCE: if (nf == 1) { CALL(PF_RESONANCE, 1); return true; } break; \
            case PF_BUMP_NOTHIRDDERIV: if (nf == 1) { CALL(PF_BUMP_NOTHIRDDERIV, 1); return true; } break; \
            case PF_HYBRIDQUADRATIC: if (nf == 2) { CALL(PF_HYBRIDQUADRATIC, 2); return true; } break; \
            case PF_RIDGE_TWOFIELD: if (nf == 2) { CALL(PF_RIDGE_TWOFIELD, 2); return true; } break; \
            case PF_QUARTICTWOFIELD: if (nf == 2) { CALL(PF_QUARTICTWOFIELD, 2); return true; } break; \
            case PF_HYBRIDQUARTIC: if (nf == 2) { CALL(PF_HYBRIDQUARTIC, 2); return true; } break; \
            case PF_INFLECTION: if (nf == 2) { CALL(PF_INFLECTION, 2); return true; } break; \
            case PF_HILLTOPAXION: if (nf == 2) { CALL(PF_HILLTOPAXION, 2); return true; } break; \
            case PF_PRODUCTEXPONENTIAL: if (nf == 2) { CALL(PF_PRODUCTEXPONENTIAL, 2); return true; } break; \
            case PF_NFLATION:                                                        \
                switch (nf) {                                                        \
                    case 1: CALL(PF_NFLATION, 1); return true;                       \
                    case 2: CALL(PF_NFLATION, 2); return true;                       \
                    case 3: CALL(PF_NFLATION, 3); return true;                       \
                    case 4: CALL(PF_NFLATION, 4); return true;                       \
                    case 5: CALL(PF_NFLATION, 5); return true;                       \
                    case 6: CALL(PF_NFLATION, 6); return true;                       \
                    case 7: CALL(PF_NFLATION, 7); return true;                       \
                    case 8: CALL(PF_NFLATION, 8); return true;                       \
                    default: break;                                                  \
                }                                                                    \
                break;                                                               \
            default: break;                                                          \
        }                                                                            \
        return false;                                                                \
    }()

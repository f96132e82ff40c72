// Generic first-order consumer (any nf, any output stride): shared by the plain kernel
// (pf_fo.cu) and the fused producer/consumer launch (pf_fused.cu).  See pf_fo.cu for the design.
#pragma once
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"

namespace pf {

template <int NF>
struct FoShape {
    static constexpr int NB = 2 * NF + 1;
    static constexpr int SW = rec_stage_width(NF);  // doubles per stage: A, R, C[NF][NF] (NF > 8: factored slot)
    static constexpr int REC = rec_of(NF);
    static constexpr int BG_OFF = 4 * SW;           // background row sits after the 4 stages
    static constexpr int CH_RAW = 16384 / (REC * 8);
    static constexpr int CH = CH_RAW > 64 ? 64 : (CH_RAW < 8 ? 8 : CH_RAW);  // steps per chunk
#ifndef PF_FO_BLOCK
#define PF_FO_BLOCK 128
#endif
    static constexpr int BLOCK = PF_FO_BLOCK;
};

// one complex (NC = 2) or one real component (NC = 1) per row, per thread
template <int NC>
__device__ __forceinline__ void store_row(double* rowbase, int64_t kk, int part, const double (&c)[NC]) {
#ifndef PF_STORE
#define PF_STORE __stcs
#endif
    if constexpr (NC == 2) {
        PF_STORE(reinterpret_cast<double2*>(rowbase) + kk, make_double2(c[0], c[1]));
    } else {
        PF_STORE(rowbase + 2 * kk + part, c[0]);
    }
}

// kv = -(A v + k^2 R2 x + C x)        cosmomodels.py:663-674   (callers pass kval = k^2)
// Every coefficient is a broadcast shared-memory load.  With one load per DFMA the LSU (one
// broadcast wavefront per clock per SM) would cap the fp64 pipe (two warp-DFMAs per clock per
// SM) at ~50 %, so for even NF the 16-byte aligned rows of C are read as 128-bit pairs.
template <int NF, int NC>
__device__ __forceinline__ void mode_rhs(const double* __restrict__ cs, double kval,
                                         const double (&x)[NF][NC], const double (&v)[NF][NC],
                                         double (&kv)[NF][NC]) {
    // odd NF: SW = 2 + NF*NF is odd, so stages 1 and 3 are only 8-byte aligned -> scalar loads
    double A, R;
    if constexpr (NF % 2 == 0) {
        const double2 ar = *reinterpret_cast<const double2*>(cs);
        A = ar.x; R = ar.y;
    } else {
        A = cs[0]; R = cs[1];
    }
    const double B = kval * R;                        // kval = k^2, R = 1/(aH)^2
    // column-by-column accumulation: the NF row sums are independent dependency chains and are
    // advanced together (ILP = NF*NC), instead of finishing one row's NF-long chain at a time.
    // k^2 R2 is folded into the diagonal of C (one add per row) instead of one FMA per element.
    double acc[NF][NC];
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[i][c] = A * v[i][c];
    if constexpr (NF % 2 == 0) {
#pragma unroll
        for (int l = 0; l < NF / 2; ++l) {
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                double2 cc = reinterpret_cast<const double2*>(cs + 2 + i * NF)[l];
                if (2 * l == i) cc.x += B;
                if (2 * l + 1 == i) cc.y += B;
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[i][c] = fma(cc.x, x[2 * l][c], acc[i][c]);
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[i][c] = fma(cc.y, x[2 * l + 1][c], acc[i][c]);
            }
        }
    } else {
#pragma unroll
        for (int l = 0; l < NF; ++l)
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                double cc = cs[2 + i * NF + l];
                if (l == i) cc += B;
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[i][c] = fma(cc, x[l][c], acc[i][c]);
            }
    }
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) kv[i][c] = -acc[i][c];
}

// Factored records (PF_REC_FACTORED, pf_records.cuh): C = d I + (p g^T + g p^T + U p p^T)/H^2, so
//   kv_i = -( A v_i + ((kR)^2 + d) x_i + p_i w + g_i z ),   z = (p.x)/H^2,  w = (g.x)/H^2 + U z
// 4 NF + 3 FMAs for the coupling instead of NF^2, and 6 + 2 NF broadcast loads instead of 2 + NF^2.
// The whole RK4 step is written out so that kv_i is consumed as soon as it exists: six state
// arrays stay live (x, v, ax, av, xt, vt) instead of seven, and p, g are re-read from shared
// memory for the update pass instead of being held across the dot products.
__device__ __forceinline__ double2 lds_pair(const double* p) {       // never merged with other loads
    double2 r;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(smem_u32(p)));
    return r;
}
__device__ __forceinline__ double lds_one(const double* p) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(smem_u32(p)));
    return r;
}

template <int NF, int NC>
struct FactStage {
    double A, Bd, z[NC], w[NC];
    // scalars of one stage: dot products of the stage argument xin with p and g
    __device__ __forceinline__ void prepare(const double* __restrict__ cs, double kval,
                                            const double (&xin)[NF][NC]) {
        double hd[5];
        if constexpr (NF % 2 == 0) {                  // slots are 16-byte aligned for even NF
            const double2 q0 = lds_pair(cs), q1 = lds_pair(cs + 2), q2 = lds_pair(cs + 4);
            hd[0] = q0.x; hd[1] = q0.y; hd[2] = q1.x; hd[3] = q1.y; hd[4] = q2.x;
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i) hd[i] = lds_one(cs + i);
        }
        A = hd[0];
        Bd = fma(kval, hd[1], hd[2]);                 // kval = k^2, hd[1] = 1/(aH)^2
        double sp[2][NC], sg[2][NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { sp[0][c] = sp[1][c] = sg[0][c] = sg[1][c] = 0.0; }
        if constexpr (NF % 2 == 0) {
#pragma unroll
            for (int l = 0; l < NF / 2; ++l) {
                const double2 pp = lds_pair(cs + 6 + 2 * l), gg = lds_pair(cs + 6 + NF + 2 * l);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    sp[0][c] = fma(pp.x, xin[2 * l][c], sp[0][c]);
                    sp[1][c] = fma(pp.y, xin[2 * l + 1][c], sp[1][c]);
                    sg[0][c] = fma(gg.x, xin[2 * l][c], sg[0][c]);
                    sg[1][c] = fma(gg.y, xin[2 * l + 1][c], sg[1][c]);
                }
            }
        } else {
#pragma unroll
            for (int l = 0; l < NF; ++l) {
                const double pp = lds_one(cs + 6 + l), gg = lds_one(cs + 6 + NF + l);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    sp[l & 1][c] = fma(pp, xin[l][c], sp[l & 1][c]);
                    sg[l & 1][c] = fma(gg, xin[l][c], sg[l & 1][c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            z[c] = (sp[0][c] + sp[1][c]) * hd[4];
            w[c] = fma(hd[3], z[c], (sg[0][c] + sg[1][c]) * hd[4]);
        }
    }
    // kv_i of this stage for field i (pi, gi = p_i, g_i)
    __device__ __forceinline__ double kv(double pi, double gi, double xi, double vi, int c) const {
        double t = fma(Bd, xi, A * vi);
        t = fma(pi, w[c], t);
        t = fma(gi, z[c], t);
        return -t;
    }
};

// visit(i, p_i, g_i) for every field, p and g re-read from the stage slot
template <int NF, class F>
__device__ __forceinline__ void for_each_pg(const double* __restrict__ cs, F&& visit) {
    if constexpr (NF % 2 == 0) {
#pragma unroll
        for (int l = 0; l < NF / 2; ++l) {
            const double2 pp = lds_pair(cs + 6 + 2 * l), gg = lds_pair(cs + 6 + NF + 2 * l);
            visit(2 * l, pp.x, gg.x);
            visit(2 * l + 1, pp.y, gg.y);
        }
    } else {
#pragma unroll
        for (int l = 0; l < NF; ++l) visit(l, lds_one(cs + 6 + l), lds_one(cs + 6 + NF + l));
    }
}

// one classical RK4 step n -> n+1 (rk4.py:27-55) from the factored record r
template <int NF, int NC>
__device__ __forceinline__ void rk4_step_factored(const double* __restrict__ r, double kval, double h,
                                                  double hh, double h6, double (&x)[NF][NC],
                                                  double (&v)[NF][NC]) {
    constexpr int SW = FoShape<NF>::SW;
    double ax[NF][NC], av[NF][NC], xt[NF][NC], vt[NF][NC];
    FactStage<NF, NC> st;
    st.prepare(r, kval, x);
    for_each_pg<NF>(r, [&](int i, double pi, double gi) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double k1 = st.kv(pi, gi, x[i][c], v[i][c], c);
            av[i][c] = k1;
            xt[i][c] = fma(hh, v[i][c], x[i][c]);
            vt[i][c] = fma(hh, k1, v[i][c]);
        }
    });
    st.prepare(r + SW, kval, xt);
    for_each_pg<NF>(r + SW, [&](int i, double pi, double gi) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double k2 = st.kv(pi, gi, xt[i][c], vt[i][c], c);
            ax[i][c] = fma(2.0, vt[i][c], v[i][c]);
            av[i][c] = fma(2.0, k2, av[i][c]);
            xt[i][c] = fma(hh, vt[i][c], x[i][c]);
            vt[i][c] = fma(hh, k2, v[i][c]);
        }
    });
    st.prepare(r + 2 * SW, kval, xt);
    for_each_pg<NF>(r + 2 * SW, [&](int i, double pi, double gi) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double k3 = st.kv(pi, gi, xt[i][c], vt[i][c], c);
            ax[i][c] = fma(2.0, vt[i][c], ax[i][c]);
            av[i][c] = fma(2.0, k3, av[i][c]);
            xt[i][c] = fma(h, vt[i][c], x[i][c]);
            vt[i][c] = fma(h, k3, v[i][c]);
        }
    });
    st.prepare(r + 3 * SW, kval, xt);
    for_each_pg<NF>(r + 3 * SW, [&](int i, double pi, double gi) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double k4 = st.kv(pi, gi, xt[i][c], vt[i][c], c);
            x[i][c] = fma(h6, ax[i][c] + vt[i][c], x[i][c]);
            v[i][c] = fma(h6, av[i][c] + k4, v[i][c]);
        }
    });
}

// one classical RK4 step n -> n+1 (rk4.py:27-55) from the dense record r
template <int NF, int NC>
__device__ __forceinline__ void rk4_step_dense(const double* __restrict__ r, double kval, double h,
                                               double hh, double h6, double (&x)[NF][NC],
                                               double (&v)[NF][NC]) {
    constexpr int SW = FoShape<NF>::SW;
    double kv[NF][NC], ax[NF][NC], av[NF][NC], xt[NF][NC], vt[NF][NC];
    mode_rhs<NF, NC>(r, kval, x, v, kv);
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            ax[i][cc] = v[i][cc];
            av[i][cc] = kv[i][cc];
            xt[i][cc] = fma(hh, v[i][cc], x[i][cc]);
            vt[i][cc] = fma(hh, kv[i][cc], v[i][cc]);
        }
    mode_rhs<NF, NC>(r + SW, kval, xt, vt, kv);
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            ax[i][cc] = fma(2.0, vt[i][cc], ax[i][cc]);
            av[i][cc] = fma(2.0, kv[i][cc], av[i][cc]);
            xt[i][cc] = fma(hh, vt[i][cc], x[i][cc]);
            vt[i][cc] = fma(hh, kv[i][cc], v[i][cc]);
        }
    mode_rhs<NF, NC>(r + 2 * SW, kval, xt, vt, kv);
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            ax[i][cc] = fma(2.0, vt[i][cc], ax[i][cc]);
            av[i][cc] = fma(2.0, kv[i][cc], av[i][cc]);
            xt[i][cc] = fma(h, vt[i][cc], x[i][cc]);
            vt[i][cc] = fma(h, kv[i][cc], v[i][cc]);
        }
    mode_rhs<NF, NC>(r + 3 * SW, kval, xt, vt, kv);
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            x[i][cc] = fma(h6, ax[i][cc] + vt[i][cc], x[i][cc]);
            v[i][cc] = fma(h6, av[i][cc] + kv[i][cc], v[i][cc]);
        }
}

template <int NF, int NC, bool FACT>
__device__ __forceinline__ void rk4_step(const double* __restrict__ r, double kval, double h, double hh,
                                         double h6, double (&x)[NF][NC], double (&v)[NF][NC]) {
    if constexpr (FACT) rk4_step_factored<NF, NC>(r, kval, h, hh, h6, x, v);
    else rk4_step_dense<NF, NC>(r, kval, h, hh, h6, x, v);
}

// nf >= 5: cap the registers at 168 so three 128-thread CTAs fit an SM (12 warps instead of 8;
// measured +3 % at nf = 8; a cap of 128 registers spills and is slower)
#ifndef PF_FO_MINB
#define PF_FO_MINB(NF) ((NF) >= 5 ? 3 : 1)
#endif
// factored records: six state arrays (96 registers at nf = 8) and no kv array.  Measured at nf = 8:
// three CTAs per SM (165 registers, no spills) 2.50e9 mode-steps/s, four (128 registers, 228 B of
// spills) 1.97e9 -- A/B switch PF_FO_FACT_MINB

#ifndef PF_FO_MINB_FACT
#define PF_FO_MINB_FACT 3
#endif
// Consumer: integrates column j of the mode matrix for the k tile bx.  `progress` (may be null)
// is the producer's published record count in fused launches (pf_fused.cu).
// (A persistent, time-lock-stepped variant of this consumer -- one co-resident wave of CTAs looping
// over (k tile, column) work items, as pf_fo1.cu does for nf = 1 -- was measured on B200 and is
// NOT used: nf = 2 full trajectory 20.4 vs 15.6 ms, nf = 8 40.1 vs 31.6 ms.  With several CTAs per
// SM the hardware's own back-filling of finished CTAs beats fixed passes whose last one is
// partly empty, and the 13 / 145 rows per time step already spread the stores evenly;
// profiles/r2_ab_lockstep.txt.)
template <int NF, int NC, bool FACT = false>
__device__ __forceinline__ void fo_consume(const FoArgs& a, int64_t bx, int j, const long long* progress) {
    using S = FoShape<NF>;
    constexpr int NB = S::NB, REC = S::REC, CH = S::CH;
    const bool pert = a.layout == PF_LAYOUT_PERT;     // perturbation rows only
    const int NV = pert ? 2 * NF * NF : nvar_of(NF);
    const int RO = pert ? -NB : 0;                    // row offset of the perturbation block
    constexpr int KPB = S::BLOCK / (3 - NC);        // k modes per CTA

    __shared__ __align__(128) double srec[2][CH * REC];
    __shared__ __align__(8) uint64_t bars[2];

    const int tid = threadIdx.x;
    if (tid >= S::BLOCK) return;                      // fused launches may carry spare threads
    const int part = (NC == 2) ? 0 : (tid & 1);
    const int kl = (NC == 2) ? tid : (tid >> 1);
    const int64_t kk = bx * KPB + kl;
    const bool live = kk < a.nk;
    const int64_t nk = a.nk;
    const int64_t op = a.opitch;                      // modes per output row (>= nk)

    const int64_t nsteps = a.t1 - a.t0;
    const int64_t nchunks = (nsteps + CH - 1) / CH;

    auto issue = [&](int64_t c) {
        const int buf = (int)(c & 1);
        const int64_t base = a.t0 + c * CH;
        const int64_t cnt = (a.t1 - base) < CH ? (a.t1 - base) : CH;
        const uint32_t bytes = (uint32_t)(cnt * REC * sizeof(double));
        if (progress) wait_progress(progress, base + cnt, a.flags);
        mbar_expect_tx(&bars[buf], bytes);
        bulk_g2s(&srec[buf][0], a.rec + base * REC, bytes, &bars[buf]);
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

    const double kmode = live ? a.k[kk] : 0.0;
    const double kval = kmode * kmode;                // k^2: the records hold 1/(aH)^2
    const int64_t ts = live ? a.tsix[kk] : INT64_MAX;
    const double h = a.h, hh = h * 0.5, h6 = h / 6.0;   // rk4.py:31-32
    const double nan = qnan();

    double x[NF][NC], v[NF][NC];
#pragma unroll
    for (int i = 0; i < NF; ++i)
#pragma unroll
        for (int c = 0; c < NC; ++c) { x[i][c] = 0.0; v[i][c] = 0.0; }

    if (live && a.t0 > ts) {                          // resume a windowed run
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            const double* sx = a.state + 2 * ((int64_t)(2 * (i * NF + j)) * nk + kk);
            const double* sv = a.state + 2 * ((int64_t)(2 * (i * NF + j) + 1) * nk + kk);
#pragma unroll
            for (int c = 0; c < NC; ++c) { x[i][c] = sx[part + c]; v[i][c] = sv[part + c]; }
        }
    }

    int64_t next_out = (a.yout != nullptr && a.out_every > 0) ? a.t0 : INT64_MAX;
    double* orow = a.yout;                            // start of the current output row block

    for (int64_t c = 0; c < nchunks; ++c) {
        const int buf = (int)(c & 1);
        mbar_wait(&bars[buf], (uint32_t)((c >> 1) & 1));
        const int64_t base = a.t0 + c * CH;
        const int cnt = (int)((a.t1 - base) < CH ? (a.t1 - base) : CH);
        for (int q = 0; q < cnt; ++q) {
            const int64_t n = base + q;
            if (n > ts && n != next_out) {
                // tight run (strided / state-only launches, bound by the fp64 pipe): this mode is
                // running and no row is stored before `lim`: nothing but record loads and RK4
                int64_t lim = next_out - base < (int64_t)cnt ? next_out - base : (int64_t)cnt;
                if (a.T - 1 - base < lim) lim = a.T - 1 - base;
                if (lim > q) {
                    const int qe = (int)lim;
                    for (; q < qe; ++q) rk4_step<NF, NC, FACT>(&srec[buf][q * REC], kval, h, hh, h6, x, v);
                    --q;                              // the for statement steps to qe
                    continue;
                }
            }
            const double* __restrict__ r = &srec[buf][q * REC];
            const bool out = (n == next_out);
            if (n < ts) {
                // rows before this mode's start stay NaN+NaNj        rk4.py:100
                if (out && live) {
                    double nn[NC];
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) nn[cc] = nan;
                    if (!pert) {
                        store_row<NC>(orow + 2 * (int64_t)(2 * j) * op, kk, part, nn);
                        store_row<NC>(orow + 2 * (int64_t)(2 * j + 1) * op, kk, part, nn);
                        if (j == 0) store_row<NC>(orow + 2 * (int64_t)(2 * NF) * op, kk, part, nn);
                    }
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        const int64_t row = RO + NB + 2 * (i * NF + j);
                        store_row<NC>(orow + 2 * row * op, kk, part, nn);
                        store_row<NC>(orow + 2 * (row + 1) * op, kk, part, nn);
                    }
                }
            } else {
                double bgv[3][NC];                    // phi_j, phi'_j, H as this thread stores them
                if (n == ts) {
                    // the start row is ystart[:, k] itself                rk4.py:106-107
                    const double* ys = a.ystart;
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        const int64_t row = NB + 2 * (i * NF + j);
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc) {
                            x[i][cc] = ys[2 * (row * nk + kk) + part + cc];
                            v[i][cc] = ys[2 * ((row + 1) * nk + kk) + part + cc];
                        }
                    }
                    const int rows3[3] = {2 * j, 2 * j + 1, 2 * NF};
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
#pragma unroll
                        for (int cc = 0; cc < NC; ++cc)
                            bgv[b][cc] = ys[2 * ((int64_t)rows3[b] * nk + kk) + part + cc];
                        if (part == 0 && a.flags) {
                            const double ref = r[S::BG_OFF + rows3[b]];
                            const double got = bgv[b][0];
                            if (!(fabs(got - ref) <= 1e-9 * fabs(ref) + 1e-300))
                                atomicOr(a.flags, PF_FLAG_BG_MISMATCH);
                        }
                    }
                } else {
                    const int rows3[3] = {2 * j, 2 * j + 1, 2 * NF};
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const double val = r[S::BG_OFF + rows3[b]];
                        if constexpr (NC == 2) { bgv[b][0] = val; bgv[b][1] = 0.0; }
                        else { bgv[b][0] = part == 0 ? val : 0.0; }
                    }
                }
                if (out) {
                    if (!pert) {
                        store_row<NC>(orow + 2 * (int64_t)(2 * j) * op, kk, part, bgv[0]);
                        store_row<NC>(orow + 2 * (int64_t)(2 * j + 1) * op, kk, part, bgv[1]);
                        if (j == 0) store_row<NC>(orow + 2 * (int64_t)(2 * NF) * op, kk, part, bgv[2]);
                    }
#pragma unroll
                    for (int i = 0; i < NF; ++i) {
                        const int64_t row = RO + NB + 2 * (i * NF + j);
                        store_row<NC>(orow + 2 * row * op, kk, part, x[i]);
                        store_row<NC>(orow + 2 * (row + 1) * op, kk, part, v[i]);
                    }
                }
                if (n + 1 < a.T) {
                    // one classical RK4 step n -> n+1                      rk4.py:27-55
                    rk4_step<NF, NC, FACT>(r, kval, h, hh, h6, x, v);
                }
            }
            if (out) {
                next_out += a.out_every;
                orow += 2 * (int64_t)NV * op;
            }
        }
        __syncthreads();                              // everyone is done reading srec[buf]
        if (tid == 0 && c + 2 < nchunks) issue(c + 2);
    }

    if (live && a.state && ts < a.t1) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
            double* sx = a.state + 2 * ((int64_t)(2 * (i * NF + j)) * nk + kk);
            double* sv = a.state + 2 * ((int64_t)(2 * (i * NF + j) + 1) * nk + kk);
#pragma unroll
            for (int c = 0; c < NC; ++c) { sx[part + c] = x[i][c]; sv[part + c] = v[i][c]; }
        }
    }
}


template <int NF, int NC>
__global__ void __launch_bounds__(FoShape<NF>::BLOCK, PF_FO_MINB(NF)) fo_rk4_kernel(FoArgs a) {
    fo_consume<NF, NC>(a, blockIdx.x, blockIdx.y, nullptr);
}

template <int NF, int NC, int MINB>
__global__ void __launch_bounds__(FoShape<NF>::BLOCK, MINB) fo_rk4_fact_kernel(FoArgs a) {
    fo_consume<NF, NC, true>(a, blockIdx.x, blockIdx.y, nullptr);
}

}  // namespace pf

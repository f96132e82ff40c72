// K2/K3 : fixed-step RK4 of the complex first-order mode matrix over a batch of k modes,
//         each mode starting at its own row, writing the reference's trajectory layout.
//
// Replaces rk4.rkdriver_tsix (rk4.py:58-138) + rk4stepks (:27-55) driving
// CanonicalFirstOrder.derivs (cosmomodels.py:625-675).
//
// Formulation (DESIGN.md "K2/K3"): the background is shared by every mode, so its rows and the
// k-independent factors of the right-hand side come from the per-step records built by
// pf_stage_table (A = U/H^2, R = 1/(aH), C = M/H^2 at each RK4 stage).  What is left per mode
// and per column j of the mode matrix is the real-coefficient linear system
//     x' = v,   v' = -(A v + (k R)^2 x + C x),      x = dphi_{.j},  v = dphi'_{.j}
// whose real and imaginary parts never mix.  One thread owns one (k, j) pair (NC = 2, complex
// in registers) or one (k, j, re|im) triple (NC = 1, used for nf >= 3 to fit the register
// file); all RK4 stage vectors stay in registers.  Consecutive lanes hold consecutive k, so
// every row store of a warp is one contiguous 512-byte (NC = 2: 128-bit per lane) or
// 256-byte (NC = 1: 64-bit per lane) segment of yresult[t][var][:].
//
// The records are streamed through shared memory in double-buffered chunks with 1-D TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx); every thread of the CTA reads the same record,
// i.e. conflict-free broadcast loads.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"
#include "pf_fo_generic.cuh"

namespace pf {

template <int NF, int NC>
static int launch_fo(const FoArgs& a, cudaStream_t st, bool factored = false) {
    using S = FoShape<NF>;
    constexpr int KPB = S::BLOCK / (3 - NC);
    dim3 grid((unsigned)((a.nk + KPB - 1) / KPB), NF, 1);
    if constexpr (NF >= 5) {
        if (factored) {
            // CTAs per SM (register cap 128 / 168): A/B switch PF_FO_FACT_MINB=3|4
            static const int minb = getenv("PF_FO_FACT_MINB") ? atoi(getenv("PF_FO_FACT_MINB")) : PF_FO_MINB_FACT;
            if (minb == 3) fo_rk4_fact_kernel<NF, NC, 3><<<grid, S::BLOCK, 0, st>>>(a);
            else fo_rk4_fact_kernel<NF, NC, 4><<<grid, S::BLOCK, 0, st>>>(a);
            return 0;
        }
    }
    fo_rk4_kernel<NF, NC><<<grid, S::BLOCK, 0, st>>>(a);
    return 0;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_fo_run(const pf_model* m, int64_t nk, const double* k_dev, const int64_t* tsix_dev,
                         const double* ystart_dev, const double* rec_dev, int64_t T, double h,
                         int64_t t0, int64_t t1, double* state_dev, double* yout_dev,
                         int64_t out_every, int32_t layout, int64_t out_pitch, int32_t* flags_dev,
                         void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (nk < 0 || T < 1 || t0 < 0 || t1 > T || t0 > t1 || out_every < 0) {
        set_error("pf_fo_run: bad sizes (nk=%lld T=%lld window=[%lld,%lld) out_every=%lld)",
                  (long long)nk, (long long)T, (long long)t0, (long long)t1, (long long)out_every);
        return PF_E_ARG;
    }
    if (!(h == h) || h == 0.0) { set_error("pf_fo_run: Need to specify h."); return PF_E_STEP; }
    if (nk == 0 || t0 == t1) return PF_OK;
    if (!k_dev || !tsix_dev || !ystart_dev || !rec_dev) { set_error("pf_fo_run: null buffer"); return PF_E_ARG; }
    if (t0 > 0 && !state_dev) { set_error("pf_fo_run: a window with t0 > 0 needs state_dev"); return PF_E_ARG; }
    if (out_every > 0 && !yout_dev) { set_error("pf_fo_run: out_every > 0 needs yout_dev"); return PF_E_ARG; }
    if (layout != PF_LAYOUT_FULL && layout != PF_LAYOUT_PERT) { set_error("pf_fo_run: unknown layout %d", layout); return PF_E_ARG; }
    if (out_pitch != 0 && out_pitch < nk) { set_error("pf_fo_run: out_pitch=%lld is smaller than nk=%lld", (long long)out_pitch, (long long)nk); return PF_E_ARG; }
    FoArgs a;
    a.layout = layout;
    a.opitch = out_pitch ? out_pitch : nk;
    a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.ystart = ystart_dev; a.rec = rec_dev;
    a.T = T; a.t0 = t0; a.t1 = t1; a.state = state_dev;
    a.yout = out_every > 0 ? yout_dev : nullptr; a.out_every = out_every; a.flags = flags_dev;
    a.h = h;
    a.lockstep = flags_dev ? flags_dev + 3 : nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool generic_only = getenv("PF_FO_GENERIC") != nullptr;
    const bool fact = rec_kind(m) == PF_REC_FACTORED;   // must match the records pf_stage_table wrote
    if (m->nfields > PF_MAX_NFIELDS) return rt_fo_run(m, a, st);
    switch (m->nfields) {
        case 1:
            if (!generic_only) launch_fo1(a, st);                     // single-field fast path
            else launch_fo<1, 2>(a, st);
            break;
        // nf = 2: thread = (k, j), complex.  A (k, j, re|im) split for strided launches (half the
        // registers, up to 8 CTAs per SM) measured 15-22 % slower on B200 (profiles/r2_ab_*).
        case 2: launch_fo<2, 2>(a, st); break;
        case 3: launch_fo<3, 1>(a, st); break;
        case 4: launch_fo<4, 1>(a, st); break;
        case 5: launch_fo<5, 1>(a, st, fact); break;
        case 6: launch_fo<6, 1>(a, st, fact); break;
        case 7: launch_fo<7, 1>(a, st, fact); break;
        case 8: launch_fo<8, 1>(a, st, fact); break;
        default: set_error("pf_fo_run: nfields=%d not compiled", m->nfields); return PF_E_NFIELDS;
    }
    PF_CUDA_CHECK(cudaGetLastError(), "pf_fo_run launch");
    return PF_OK;
}

extern "C" int64_t pf_solve_scratch_doubles(int nfields, int64_t T) {
    if (nfields < 1 || nfields > PF_MAX_NFIELDS_RT || T < 0) return -1;
    return T * (int64_t)(rec_of(nfields) + 5 * nbg_of(nfields));
}

extern "C" int pf_fo_solve(const pf_model* m, int64_t nk, const double* k_dev, const int64_t* tsix_dev,
                           const double* ystart_dev, int64_t n0, const double* y0_host, int64_t T,
                           double simtstart, double h, int64_t t1, double* scratch_dev,
                           double* state_dev, double* yout_dev, int64_t out_every, int32_t layout,
                           int64_t out_pitch, int32_t* ws_dev, void* stream) {
    int rc = check_model(m);
    if (rc) return rc;
    if (nk < 0 || T < 1 || t1 < 0 || t1 > T || out_every < 0) { set_error("pf_fo_solve: bad sizes"); return PF_E_ARG; }
    if (!(h == h) || h == 0.0) { set_error("pf_fo_solve: Need to specify h."); return PF_E_STEP; }
    if (n0 < 0 || n0 >= T) { set_error("pf_fo_solve: Start times outside range of steps."); return PF_E_START; }
    if (!y0_host || !scratch_dev || !ws_dev) { set_error("pf_fo_solve: null buffer"); return PF_E_ARG; }
    if (nk > 0 && (!k_dev || !tsix_dev || !ystart_dev)) { set_error("pf_fo_solve: null buffer"); return PF_E_ARG; }
    if (out_every > 0 && !yout_dev) { set_error("pf_fo_solve: out_every > 0 needs yout_dev"); return PF_E_ARG; }
    if (layout != PF_LAYOUT_FULL && layout != PF_LAYOUT_PERT) { set_error("pf_fo_solve: unknown layout %d", layout); return PF_E_ARG; }
    if (out_pitch != 0 && out_pitch < nk) { set_error("pf_fo_solve: out_pitch=%lld is smaller than nk=%lld", (long long)out_pitch, (long long)nk); return PF_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const int nf = m->nfields;
    double* rec = scratch_dev;
    double* stagey = rec + T * (int64_t)rec_of(nf);
    double* bg = stagey + T * (int64_t)(4 * nbg_of(nf));
    PF_CUDA_CHECK(cudaMemsetAsync(ws_dev, 0, 8 * sizeof(int32_t), st), "pf_fo_solve memset");
    static const bool no_fuse = getenv("PF_FO_NOFUSE") != nullptr;
    if (nf == 1 && nk > 0 && t1 > 0 && !no_fuse) {
        FoArgs a;
        a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.ystart = ystart_dev; a.rec = rec;
        a.T = T; a.t0 = 0; a.t1 = t1; a.state = state_dev;
        a.yout = out_every > 0 ? yout_dev : nullptr; a.out_every = out_every;
        a.flags = ws_dev + 1; a.h = h; a.layout = layout; a.opitch = out_pitch ? out_pitch : nk; a.lockstep = nullptr;
        if (launch_fo1_fused(m, a, simtstart, n0, y0_host, rec, ws_dev, st)) {
            PF_CUDA_CHECK(cudaGetLastError(), "pf_fo_solve fused launch");
            return PF_OK;
        }
    }
    if (nf == 2 && nk > 0 && t1 > 0 && !no_fuse) {
        FoArgs a;
        a.nk = nk; a.k = k_dev; a.tsix = tsix_dev; a.ystart = ystart_dev; a.rec = rec;
        a.T = T; a.t0 = 0; a.t1 = t1; a.state = state_dev;
        a.yout = out_every > 0 ? yout_dev : nullptr; a.out_every = out_every;
        a.flags = ws_dev + 1; a.h = h; a.layout = layout; a.opitch = out_pitch ? out_pitch : nk; a.lockstep = nullptr;
        if (launch_fo_fused_generic(m, a, simtstart, n0, y0_host, rec, stagey, ws_dev, st)) {
            PF_CUDA_CHECK(cudaGetLastError(), "pf_fo_solve fused launch");
            return PF_OK;
        }
    }
    rc = pf_bg_run(m, T, h, n0, y0_host, bg, stagey, stream);
    if (rc) return rc;
    rc = pf_stage_table(m, T, simtstart, h, n0, bg, stagey, rec, stream);
    if (rc) return rc;
    return pf_fo_run(m, nk, k_dev, tsix_dev, ystart_dev, rec, T, h, 0, t1, state_dev, yout_dev, out_every,
                     layout, out_pitch, ws_dev + 1, stream);
}

// Register-resident mode kernels for nflation with 9..16 fields.
//
// cmpotentials.py:791-846 accepts any nfields.  Up to 8 fields pf_fo.cu runs compiled
// instantiations; above that pf_rt.cu keeps the per-thread state in local memory.  For 9..16 fields
// a column of the mode matrix (six live arrays of NF doubles: x, v, ax, av, xt, vt) still fits the
// 255-register budget of a thread, so the factored consumer of pf_fo_generic.cuh is instantiated
// for those counts too: same record layout as pf_rt.cu's stage table writes
// (A, R2, d, U, 1/H^2, 0, p[nf], g[nf] per stage; rec_stage_width), records streamed through shared
// memory by TMA, 4 nf + 3 FMAs per stage for the coupling.  Two 128-thread CTAs per SM.
#include "pf_common.cuh"
#include "pf_fo_shared.cuh"
#include "pf_fo_generic.cuh"

namespace pf {

template <int NF>
static void launch_wide(const FoArgs& a, cudaStream_t st) {
    using S = FoShape<NF>;
    constexpr int KPB = S::BLOCK / 2;                 // thread <-> (k, j, re|im)
    dim3 grid((unsigned)((a.nk + KPB - 1) / KPB), NF, 1);
    fo_rk4_fact_kernel<NF, 1, 2><<<grid, S::BLOCK, 0, st>>>(a);
}

// returns false when nf has no register-resident instantiation (caller falls back to fo_rt_kernel)
bool wide_fo_run(int nf, const FoArgs& a, cudaStream_t st) {
    switch (nf) {
        case 9: launch_wide<9>(a, st); return true;
        case 10: launch_wide<10>(a, st); return true;
        case 11: launch_wide<11>(a, st); return true;
        case 12: launch_wide<12>(a, st); return true;
        case 13: launch_wide<13>(a, st); return true;
        case 14: launch_wide<14>(a, st); return true;
        case 15: launch_wide<15>(a, st); return true;
        case 16: launch_wide<16>(a, st); return true;
        default: return false;
    }
}

}  // namespace pf

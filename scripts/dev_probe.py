"""Development probe: times K1 / K1b / K2 separately on one GPU (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import cosmomodels, engine

def main(pot="lambdaphi4", nf=1, nk=2**16, bgy=(25.0, 0.0, 0.0), params=None):
    params = params or {"nfields": nf}
    k = np.linspace(0.85e-60, 8.2165e-58, nk)
    m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func=pot, pot_params=params, nfields=nf,
                                        bgystart=None if bgy is None else np.array(bgy), cq=50)
    t = time.time(); m.runbg(); t1 = time.time(); m.setfoics(); t2 = time.time()
    print("runbg %.3fs setfoics %.3fs" % (t1 - t, t2 - t1))
    T = int(np.around((m.fotend - 0) / 0.01) + 1)
    pm = engine.make_model(pot, params, nf, m.ainit)
    prob = engine.FirstOrderProblem(pm, k, m.fotstartindex, m.foystart, 0.0, T, 0.01)
    n0 = int(m.fotstartindex.min())
    y0 = np.real(m.foystart[0:2 * nf + 1, 0])
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(); prob.build_tables(y0, n0); e1.record()
        torch.cuda.synchronize()
        print("K1+K1b %.3f ms" % e0.elapsed_time(e1))
    if os.environ.get("PF_WINDOW"):
        # one full-output time window [t0, t1) of a windowed run (state carried in from row t0)
        t0w, t1w = (int(v) for v in os.environ["PF_WINDOW"].split(","))
        prob.build_tables(y0, n0)
        prob.launch(0, t0w, None, 0)
        state0 = prob.state.clone()
        outw = torch.empty((t1w - t0w, prob.nvar, nk), dtype=torch.complex128, device="cuda")
        for rep in range(4):
            prob.state.copy_(state0)
            e0, e1 = ev(), ev()
            e0.record(); prob.launch(t0w, t1w, outw, 1); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("WINDOW [%d,%d) %.3f ms  write %.1f GB/s" % (t0w, t1w, ms, outw.numel() * 16 / ms / 1e6))
        prob.check_flags()
        return
    every = int(os.environ.get("PF_EVERY", "1"))
    nrows = (T + every - 1) // every
    out = torch.empty((nrows, prob.nvar, nk), dtype=torch.complex128, device="cuda")
    steps = float(np.sum(T - 1 - m.fotstartindex))
    flops = {1: 104, 2: 456, 8: 13320}.get(nf, 0)
    for rep in range(4):
        e0, e1 = ev(), ev()
        e0.record(); prob.launch(0, T, out, every); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("K2 %.3f ms  %.3e mode-steps/s  write %.1f GB/s (all rows) %.1f GB/s (active)  %.2f TFLOP/s" % (
            ms, steps / ms * 1e3, out.numel() * 16 / ms / 1e6, steps * 16 * prob.nvar / every / ms / 1e6, steps * flops / ms / 1e9))
    for rep in range(3):
        e0, e1 = ev(), ev()
        e0.record(); prob.solve(y0, n0, T, out, every); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("SOLVE every=%d (K1+K1b+K2) %.3f ms  %.3e mode-steps/s" % (every, ms, steps / ms * 1e3))
    if every != 1 or nf != 1:
        prob.check_flags(); return
    if os.environ.get("PF_FILL_PROBE"):
        flat = out.view(torch.float64).view(-1)
        for rep in range(3):
            e0, e1 = ev(), ev()
            e0.record(); flat.fill_(1.5); e1.record(); torch.cuda.synchronize()
            print("torch fill %.3f ms %.1f GB/s" % (e0.elapsed_time(e1), flat.numel() * 8 / e0.elapsed_time(e1) / 1e6))
        half = flat.numel() // 2
        for rep in range(3):
            e0, e1 = ev(), ev()
            e0.record(); flat[:half].copy_(flat[half:2 * half]); e1.record(); torch.cuda.synchronize()
            print("torch copy %.3f ms %.1f GB/s (r+w)" % (e0.elapsed_time(e1), half * 16 / e0.elapsed_time(e1) / 1e6))
    for rep in range(4):
        e0, e1 = ev(), ev()
        e0.record(); prob.solve(y0, n0, T, out, 1); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("SOLVE (K1+K1b+K2) %.3f ms  %.3e mode-steps/s" % (ms, steps / ms * 1e3))
    prob.check_flags()
    print("T", T, "nk", nk, "tsix", m.fotstartindex.min(), m.fotstartindex.max())

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    if which == "c2": main(nk=int(sys.argv[2]) if len(sys.argv) > 2 else 2 ** 16)
    elif which == "c1": main("msqphisq", 1, 2053, (18.0, -0.1, 0.0))
    elif which == "c3": main("hybridquadratic", 2, int(sys.argv[2]) if len(sys.argv) > 2 else 2**16, (12.0, 1/300.0, 12.0, 49/300.0, 0))
    elif which == "c5": main("msqphisq", 1, int(sys.argv[2]) if len(sys.argv) > 2 else 2**20, (18.0, -0.1, 0.0))
    elif which.startswith("nf"):      # e.g. nf10 16384: nflation with that many fields (runtime-nfields kernels above 8)
        nf = int(which[2:]); main("nflation", nf, int(sys.argv[2]) if len(sys.argv) > 2 else 2**13, None, {"nfields": nf})
    elif which == "c4": main("nflation", 8, int(sys.argv[2]) if len(sys.argv) > 2 else 2**13, None, {"nfields": 8})

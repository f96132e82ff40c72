"""Small end-to-end runs of every kernel, meant to be run under compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import cosmomodels as c, engine

def run(pot, nf, bgy, nk, params=None, every=1):
    k = np.linspace(0.85e-60, 8.2165e-58, nk)
    m = c.FOCanonicalTwoStage(k=k, potential_func=pot, pot_params=params or {"nfields": nf}, nfields=nf,
                              bgystart=None if bgy is None else np.array(bgy), cq=50, tend=83.0)
    m.run(saveresults=False)                       # bg kernel + fused/unfused solve + windowed host path
    assert np.isfinite(m.yresult[-1]).all()
    pr = m.Pr[-1]
    assert np.isfinite(pr).all()
    print(pot, nf, nk, m.yresult.shape, "P_R[0] = %.6e" % pr[0])

run("msqphisq", 1, (18.0, -0.1, 0.0), 300)
run("hybridquadratic", 2, (12.0, 1 / 300.0, 12.0, 49 / 300.0, 0), 70)
run("nflation", 3, None, 20, {"nfields": 3})
run("nflation", 8, None, 5, {"nfields": 8})
torch.cuda.synchronize()
print("done")

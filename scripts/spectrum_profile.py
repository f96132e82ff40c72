"""Where does the wall time of engine.spectrum_run go? (development probe)"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import engine
import cProfile, pstats
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 22
k = np.linspace(0.85e-60, 8.2165e-58, nk)
kw = dict(nfields=1, pot_params={"nfields": 1}, bgystart=np.array([18.0, -0.1, 0.0]), cq=50)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = engine.spectrum_run("msqphisq", k, **kw)
    pr = r["scaled_Pr"][-1].cpu().numpy()
    torch.cuda.synchronize(); print("spectrum_run nk=%d: %.1f ms" % (nk, 1e3 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable()
r = engine.spectrum_run("msqphisq", k, **kw); x = r["scaled_Pr"][-1].cpu().numpy(); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

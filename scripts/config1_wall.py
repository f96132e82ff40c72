"""Wall time of the reference's own default case (BASELINE config 1: msqphisq, 2053-mode K4 grid,
full trajectory 1.34 GB) through the drop-in API, as a user sees it."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
t00 = time.perf_counter()
import numpy as np, torch
from pyflation_b200 import cosmomodels as c, helpers
t0 = time.perf_counter(); print("imports %.2f s" % (t0 - t00))
k = helpers.seq(0.85e-60, helpers.getkend(0.85e-60, 0.4e-60, 1025), 0.4e-60)
for rep in range(3):
    t0 = time.perf_counter()
    m = c.FOCanonicalTwoStage(k=k, potential_func="msqphisq", pot_params={"nfields": 1}, nfields=1,
                              bgystart=np.array([18.0, -0.1, 0]), cq=50, solver="rkdriver_tsix")
    m.runbg(); t1 = time.perf_counter(); m.setfoics(); t2 = time.perf_counter(); m.runfo(); t3 = time.perf_counter()
    pr = m.Pr[-1]; t4 = time.perf_counter()
    print("run %d: runbg %.3f s, setfoics %.3f s, runfo %.3f s (yresult %.2f GB), Pr[-1] %.3f s; total %.3f s"
          % (rep, t1 - t0, t2 - t1, t3 - t2, m.yresult.nbytes / 1e9, t4 - t3, t4 - t0))

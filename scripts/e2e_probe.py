"""Development probe: where does the host-delivery time go? (not the bench contract)"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import cosmomodels, engine

nk = 2 ** 16
k = np.linspace(0.85e-60, 8.2165e-58, nk)
m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func="lambdaphi4", pot_params={"nfields": 1}, nfields=1,
                                    bgystart=np.array([25.0, 0.0, 0.0]), cq=50)
m.runbg(); m.setfoics()
T = int(np.around(m.fotend / 0.01) + 1)
pm = engine.make_model("lambdaphi4", {"nfields": 1}, 1, m.ainit)
host = torch.empty((T, 5, nk), dtype=torch.complex128, pin_memory=True)
y0 = np.real(m.foystart[0:3, 0]).copy(); n0 = int(m.fotstartindex.min())
ys = torch.as_tensor(m.foystart).pin_memory().numpy(); ts = np.ascontiguousarray(m.fotstartindex)

def run(label, **kw):
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        prob = engine.FirstOrderProblem(pm, k, ts, ys, 0.0, T, 0.01)
        t1 = time.perf_counter()
        engine.first_order_host(prob, out=host, y0=y0, n0=n0, **kw)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        best = min(best, t2 - t0)
    print("%-40s total %.1f ms (setup %.1f ms)" % (label, best * 1e3, (t1 - t0) * 1e3))

run("compact (host buffer: torch pin_memory)")
run("compact (host buffer: torch pin_memory)")
for nt in (2, 4, 6, 8, 12):
    run("compact, %d fill threads" % nt, fill_threads=nt)
os.environ["PF_PROBE_NO_FILL"] = "1"
run("compact, DMA only (no host fill: probe)")
del os.environ["PF_PROBE_NO_FILL"]
if os.environ.get("PF_PROBE_SHORT"):
    sys.exit(0)
del host
host = engine._pinned_empty((T, 5, nk), torch.complex128)
print("huge-page registered buffer: pinned =", host.is_pinned())
run("compact (host buffer: THP + cudaHostRegister)")
run("compact (host buffer: THP + cudaHostRegister)")

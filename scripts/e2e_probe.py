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

os.environ["PF_E2E_NOFILL"] = "1"
run("compact, NO fill (copies only)")
del os.environ["PF_E2E_NOFILL"]
import ctypes
from pyflation_b200 import _lib
rec = np.random.rand(T, 16)
t0 = time.perf_counter()
_lib.lib().pf_host_fill_background(ctypes.c_void_p(host.data_ptr()), 1, nk, 0, T, ctypes.c_void_p(rec.ctypes.data), 16, 12,
                                   ctypes.c_void_p(ts.ctypes.data), ctypes.c_void_p(ys.ctypes.data), 16)
print("fill alone (16 threads, whole result): %.1f ms" % ((time.perf_counter() - t0) * 1e3))
run("full layout (all rows over PCIe)", compact=False)
for nt in (4, 8, 12, 16):
    run("compact, fill threads=%d" % nt, fill_threads=nt)
for slab in (256 << 20, 512 << 20, 2 << 30):
    run("compact, 16 threads, slab=%d MiB" % (slab >> 20), slab_bytes=slab)

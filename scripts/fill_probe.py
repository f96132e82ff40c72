"""Development probe: host fill bandwidth of pf_host_fill_background vs thread count."""
import sys, os, time, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import _lib
L = _lib.lib()
T, nk = 2000, 65536
out = torch.empty((T, 5, nk), dtype=torch.complex128, pin_memory=True)
out.zero_()
rec = np.random.rand(T, 16)
tsix = np.sort(np.random.randint(100, 400, nk)).astype(np.int64)
ystart = (np.random.rand(5, nk) + 1j * np.random.rand(5, nk))
for nt in (1, 4, 8, 12, 16, 24, 32):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        rc = L.pf_host_fill_background(ctypes.c_void_p(out.data_ptr()), 1, nk, 0, T, ctypes.c_void_p(rec.ctypes.data), 16, 12,
                                       ctypes.c_void_p(tsix.ctypes.data), ctypes.c_void_p(ystart.ctypes.data), 0, nt)
        best = min(best, time.perf_counter() - t0)
    print("threads %2d: %.1f GB/s (rc %d)" % (nt, T * 3 * nk * 16 / best / 1e9, rc))
y = out.numpy()
assert np.isnan(y[50, 0, :]).all() and y[1000, 1, 7] == rec[1000, 13] and y[1000, 3, 7] == 0
print("cpus", len(os.sched_getaffinity(0)), os.cpu_count())

import sys, time, gc
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from pyflation_b200 import engine
torch.cuda.init()
for nbytes_gb in (0.5, 3, 41.2):
    shape = (int(nbytes_gb * (1 << 30) // (5 * 65536 * 16)), 5, 65536)
    t0 = time.perf_counter(); t = engine._pinned_empty(shape, torch.complex128); dt = time.perf_counter() - t0
    print("%.1f GB: alloc %.2f s pinned=%s" % (nbytes_gb, dt, t.is_pinned()))
    d = torch.ones((4, 5, 65536), dtype=torch.complex128, device="cuda")
    t[:4].copy_(d, non_blocking=True); torch.cuda.synchronize()
    assert t[3, 4, 100] == 1
    a = t.numpy(); del t
    assert a[3, 4, 100] == 1
    del a; gc.collect()
print("ok")

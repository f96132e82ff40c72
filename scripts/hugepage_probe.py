"""Probe: does a transparent-huge-page backed, cudaHostRegister'ed result buffer speed up the host
fill and the D2H copies compared with torch's cudaHostAlloc pinned memory?"""
import ctypes, mmap, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from pyflation_b200 import _lib
L = _lib.lib()
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
T, nk = 7854, 65536
nbytes = T * 5 * nk * 16
rec = np.random.rand(T, 16); tsix = np.sort(np.random.randint(1100, 1800, nk)).astype(np.int64)
ys = np.random.rand(5, nk) + 1j * np.random.rand(5, nk)
dev = torch.empty((512, 2, nk), dtype=torch.complex128, device="cuda")

def bench(ptr, label):
    for rep in range(2):
        t0 = time.perf_counter()
        L.pf_host_fill_background(ctypes.c_void_p(ptr), 1, nk, 0, T, ctypes.c_void_p(rec.ctypes.data), 16, 12,
                                  ctypes.c_void_p(tsix.ctypes.data), ctypes.c_void_p(ys.ctypes.data), 0, 16)
        dt = time.perf_counter() - t0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for w in range(T // 512):
        L.pf_copy_rows_d2h(ctypes.c_void_p(ptr + (w * 512 * 5 + 3) * nk * 16), 5 * nk * 16, ctypes.c_void_p(dev.data_ptr()),
                           2 * nk * 16, 2 * nk * 16, 512, None)
    torch.cuda.synchronize(); dc = time.perf_counter() - t0
    print("%-34s fill %.1f GB/s   d2h %.1f GB/s" % (label, T * 3 * nk * 16 / dt / 1e9, (T // 512) * 512 * 2 * nk * 16 / dc / 1e9))

t0 = time.perf_counter(); a = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); print("cudaHostAlloc %.1f s" % (time.perf_counter() - t0))
bench(a.data_ptr(), "torch pin_memory (cudaHostAlloc)")
del a
t0 = time.perf_counter()
mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
buf = np.frombuffer(mm, dtype=np.uint8)
addr = (buf.ctypes.data + (2 << 20) - 1) & ~((2 << 20) - 1)
libc = ctypes.CDLL(None, use_errno=True)
rc = libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(nbytes), 14)       # MADV_HUGEPAGE
ctypes.memset(ctypes.c_void_p(addr), 0, nbytes)                              # touch
rt = torch.cuda.cudart()
rr = rt.cudaHostRegister(addr, nbytes, 0)
print("mmap+madvise(rc=%d)+touch+cudaHostRegister(%s) %.1f s" % (rc, rr, time.perf_counter() - t0))
print("AnonHugePages:", [l for l in open("/proc/meminfo") if "AnonHuge" in l][0].strip())
bench(addr, "mmap THP + cudaHostRegister")
rt.cudaHostUnregister(addr)

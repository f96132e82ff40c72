"""Device-side orchestration of the hot path: torch owns buffers and streams, the work is done
by the hand-written kernels behind the C ABI (``include/pyflation_b200.h``).

Call graph of one first-order solve (what ``rk4.rkdriver_tsix`` does in the reference,
rk4.py:58-138, for a ``CanonicalFirstOrder.derivs``):

    background_tables()   K1  pf_bg_run       serial RK4 of the shared background from column
                                              0's start row, keeping the 4 stage arguments
                          K1b pf_stage_table  per-step records (A, R, C) + bg row
    first_order_device()  K2  pf_fo_run       one launch over [0, T): full trajectory in HBM
    first_order_host()    K2  pf_fo_run       time windows into two device slabs, each copied
                                              to (pinned) host memory while the next computes
"""
import collections
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import PfModel


def nvar_of(nf):
    return 2 * nf * nf + 2 * nf + 1


def make_model(potential_func, pot_params, nfields, ainit=1.0):
    """Resolve (name, params dict, nfields) into the C struct, with the reference's defaults
    (cmpotentials.py) for parameters that are absent."""
    table = _lib.potential_table()
    if potential_func not in table:
        raise AttributeError("module 'cmpotentials' has no attribute %r" % (potential_func,))
    pid, nf_required, plist = table[potential_func]
    nfields = int(nfields)
    if nf_required and nf_required != nfields:
        raise ValueError("potential %s is defined for nfields=%d, not %d"
                         % (potential_func, nf_required, nfields))
    nf_max = _lib.PF_MAX_NFIELDS_RT if potential_func == "nflation" else _lib.PF_MAX_NFIELDS
    if not 1 <= nfields <= nf_max:
        raise NotImplementedError("nfields=%d: the kernels take 1..%d fields" % (nfields, nf_max))
    m = PfModel()
    m.potential_id = pid
    m.nfields = nfields
    pot_params = pot_params or {}
    if potential_func == "nflation" and "nfields" not in pot_params:
        raise KeyError("nfields")                       # cmpotentials.py:828
    for i, (name, default) in enumerate(plist):
        m.params[i] = float(pot_params.get(name, default))
    m.ainit = float(np.asarray(ainit).reshape(-1)[0])
    return m


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, int):                              # raw device address (peer memory)
        return ctypes.c_void_p(t)
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.LibraryError("pyflation_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    return dev


# ---- stage arguments of recent background runs ---------------------------------------------
# FOCanonicalTwoStage.run() integrates the background twice in the reference: once in runbg()
# and again, as rows 0..2nf of every column, inside the first-order run (cosmomodels.py:1343-1404).
# Here runbg() already ran K1 with the same arithmetic the first-order solve would use, so its
# RK4 stage arguments (T x 4 x (2nf+1) doubles, <= 5 MB) stay on the device; a first-order solve
# whose column 0 starts ON that trajectory (bitwise) rebuilds its step records from them with
# K1b (parallel, microseconds) and skips K1's serial chain (2.4-3 ms) altogether.
_STAGE_CACHE = collections.OrderedDict()
_STAGE_CACHE_MAX = 4


def _stage_key(model, h, simtstart, dev):
    return (int(model.potential_id), int(model.nfields), tuple(model.params), float(h),
            float(simtstart), str(dev))


def remember_background(model, h, simtstart, n0, bg_host, stagey, dev):
    key = _stage_key(model, h, simtstart, dev)
    _STAGE_CACHE.pop(key, None)
    _STAGE_CACHE[key] = dict(n0=int(n0), T=int(bg_host.shape[0]), bg_host=bg_host, stagey=stagey)
    while len(_STAGE_CACHE) > _STAGE_CACHE_MAX:
        _STAGE_CACHE.popitem(last=False)


def cached_stage_args(model, h, simtstart, n0, y0, T, dev):
    """(stagey device tensor [T_bg, 4, 2nf+1], first valid row) when a remembered background run
    passes through ``y0`` at row ``n0`` bit for bit and covers rows [n0, T); else None."""
    if os.environ.get("PF_NO_STAGE_CACHE"):
        return None
    e = _STAGE_CACHE.get(_stage_key(model, h, simtstart, dev))
    if e is None or n0 < e["n0"] or T > e["T"]:
        return None
    if not np.array_equal(e["bg_host"][n0], np.asarray(y0, dtype=np.float64).reshape(-1)):
        return None
    return e["stagey"], e["n0"]


def background_tables(model, y0, T, h, simtstart=0.0, n0=0, records=True, device=None,
                      return_stage=False):
    """K1 (+K1b).  Returns (bg [T, 2nf+1] f64, rec [T, REC] f64 or None) on the device
    (``return_stage``: also the stage arguments [T, 4, 2nf+1])."""
    L = _lib.lib()
    dev = require_cuda(device)
    nf = model.nfields
    nb = 2 * nf + 1
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(-1))
    if y0.shape[0] != nb:
        raise ValueError("background start vector must have %d entries" % nb)
    with torch.cuda.device(dev):
        bg = torch.empty((T, nb), dtype=torch.float64, device=dev)
        stagey = torch.empty((T, 4, nb), dtype=torch.float64, device=dev) \
            if (records or return_stage) else None
        st = _stream_ptr()
        _lib.check(L.pf_bg_run(ctypes.byref(model), T, float(h), int(n0),
                               y0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                               _ptr(bg), _ptr(stagey), st))
        rec = None
        if records:
            rec = torch.empty((T, int(L.pf_rec_doubles(nf))), dtype=torch.float64, device=dev)
            _lib.check(L.pf_stage_table(ctypes.byref(model), T, float(simtstart), float(h), int(n0),
                                        _ptr(bg), _ptr(stagey), _ptr(rec), st))
    if return_stage:
        return bg, rec, stagey
    return bg, rec


class FirstOrderProblem(object):
    """Inputs of one first-order solve, resident on the device."""

    def __init__(self, model, k, tsix, ystart, simtstart, T, h, device=None):
        self.dev = require_cuda(device)
        self.model = model
        self.nf = model.nfields
        self.nvar = nvar_of(self.nf)
        self.T = int(T)
        self.h = float(h)
        self.simtstart = float(simtstart)
        self.nk = int(np.shape(k)[0])
        as_dev = lambda a, dt: (a.to(self.dev, dt).contiguous() if torch.is_tensor(a)
                                else torch.as_tensor(np.ascontiguousarray(a)).to(self.dev, dt))
        self.k = as_dev(k, torch.float64)
        self.tsix = as_dev(tsix, torch.int64)
        self.ystart = as_dev(ystart, torch.complex128)
        # host copies, used when the result is delivered to host memory (background fill)
        self.tsix_host = None if torch.is_tensor(tsix) else np.ascontiguousarray(tsix, dtype=np.int64)
        self.ystart_host = None if torch.is_tensor(ystart) else \
            np.ascontiguousarray(ystart, dtype=np.complex128)
        if tuple(self.ystart.shape) != (self.nvar, self.nk):
            raise ValueError("ystart must have shape (%d, %d)" % (self.nvar, self.nk))
        self.ws = torch.zeros(8, dtype=torch.int32, device=self.dev)    # [1] = PF_FLAG_* bits
        self.flags = self.ws[1:2]
        self.scratch = None
        self.state = torch.zeros((2 * self.nf * self.nf, self.nk), dtype=torch.complex128,
                                 device=self.dev)
        self.rec = None
        self.bg = None
        self.rec_ready = False          # records already valid: solve() skips K1

    def use_stage_args(self, stagey, first_row):
        """Step records from the stage arguments of an earlier background run (K1b only)."""
        L = _lib.lib()
        with torch.cuda.device(self.dev):
            self.rec = torch.empty((self.T, int(L.pf_rec_doubles(self.nf))), dtype=torch.float64,
                                   device=self.dev)
            _lib.check(L.pf_stage_table(ctypes.byref(self.model), self.T, self.simtstart, self.h,
                                        int(first_row), None, _ptr(stagey), _ptr(self.rec),
                                        _stream_ptr()))
        self.rec_ready = True

    def build_tables(self, y0, n0):
        self.bg, self.rec = background_tables(self.model, y0, self.T, self.h, self.simtstart,
                                              n0, True, self.dev)

    def solve(self, y0, n0, t1, yout, out_every=1, stream=None, layout=0, out_pitch=0):
        """pf_fo_solve: K1 + K1b + K2 for the window [0, t1) -- one fused launch for single-field
        full-trajectory runs.  Leaves the step records in ``self.rec`` for later windows.
        ``yout`` is a device tensor or a raw device address (int: peer memory of a fused
        gather, with ``out_pitch`` = width of the assembled result)."""
        L = _lib.lib()
        nf = self.nf
        if self.rec_ready:
            self.ws.zero_()
            return self.launch(0, t1, yout, out_every, stream, layout, out_pitch)
        y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(-1))
        if y0.shape[0] != 2 * nf + 1:
            raise ValueError("background start vector must have %d entries" % (2 * nf + 1))
        with torch.cuda.device(self.dev):
            if self.scratch is None:
                self.scratch = torch.empty(int(L.pf_solve_scratch_doubles(nf, self.T)),
                                           dtype=torch.float64, device=self.dev)
            _lib.check(L.pf_fo_solve(ctypes.byref(self.model), self.nk, _ptr(self.k), _ptr(self.tsix),
                                     _ptr(self.ystart), int(n0),
                                     y0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), self.T,
                                     self.simtstart, self.h, int(t1), _ptr(self.scratch),
                                     _ptr(self.state), _ptr(yout),
                                     int(out_every if yout is not None else 0), int(layout),
                                     int(out_pitch), _ptr(self.ws), _stream_ptr(stream)))
        nrec = int(L.pf_rec_doubles(nf))
        self.rec = self.scratch[:self.T * nrec].view(self.T, nrec)

    def launch(self, t0, t1, yout, out_every=1, stream=None, layout=0, out_pitch=0):
        """K2: rows [t0, t1) -> yout (device c128 tensor, raw device address, or None)."""
        L = _lib.lib()
        _lib.check(L.pf_fo_run(ctypes.byref(self.model), self.nk, _ptr(self.k), _ptr(self.tsix),
                               _ptr(self.ystart), _ptr(self.rec), self.T, self.h, int(t0), int(t1),
                               _ptr(self.state), _ptr(yout), int(out_every if yout is not None else 0),
                               int(layout), int(out_pitch), _ptr(self.flags), _stream_ptr(stream)))

    def check_flags(self):
        f = int(self.flags.item())
        if f & _lib.PF_FLAG_TIMEOUT:
            raise _lib.LibraryError(
                "fused first-order launch: a CTA gave up waiting for the producer / its lock-step "
                "partners (instrumented or time-sliced run?); the result is invalid -- rerun with "
                "PF_FO_NOFUSE=1 (three plain launches)")
        if f & _lib.PF_FLAG_BG_MISMATCH:
            raise NotImplementedError(
                "first-order start vector: the background rows of some column are not on column "
                "0's background trajectory; the fused kernel shares one background across modes "
                "(DESIGN.md) and there is no per-column fallback")


def first_order_device(prob, out=None, out_every=1, y0=None, n0=None):
    """Trajectory kept in HBM: returns a device tensor (ceil(T/out_every), nvar, nk) complex128
    holding rows 0, out_every, 2*out_every, ... (out_every=1: the full trajectory).
    With ``y0, n0`` (column 0's background start) the tables are built by the same call
    (pf_fo_solve); otherwise ``prob.build_tables`` must have run."""
    with torch.cuda.device(prob.dev):
        nrows = (prob.T + out_every - 1) // out_every
        if out is None:
            out = torch.empty((nrows, prob.nvar, prob.nk), dtype=torch.complex128, device=prob.dev)
        if y0 is not None:
            prob.solve(y0, n0, prob.T, out, out_every)
        else:
            prob.launch(0, prob.T, out, out_every)
    return out


_HUGE_PINNED_MIN = 8 << 30      # below this torch's caching pinned allocator wins on repeated runs


def _release_registered(addr, mm):
    try:
        torch.cuda.cudart().cudaHostUnregister(addr)
    except Exception:                                   # CUDA may already be torn down at exit
        pass
    try:
        mm.close()
    except Exception:
        pass


def _pinned_empty(shape, dtype):
    """Page-locked host tensor for a result.  Buffers up to 8 GiB come from torch's caching pinned
    allocator (a repeated 1.3 GB run then costs 19 ms instead of 0.3 s).  Larger results (a 41 GB
    trajectory takes cudaHostAlloc 28 s) are mapped with transparent huge pages and registered
    with cudaHostRegister instead (9.7 s, same copy and fill rates: scripts/hugepage_probe.py);
    pageable memory is the last resort."""
    import mmap
    import weakref
    nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    if nbytes >= _HUGE_PINNED_MIN and torch.cuda.is_available():
        try:
            align = 2 << 20
            mm = mmap.mmap(-1, nbytes + align, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            raw = np.frombuffer(mm, dtype=np.uint8)
            addr = (raw.ctypes.data + align - 1) & ~(align - 1)
            off = addr - raw.ctypes.data
            try:
                ctypes.CDLL(None).madvise(ctypes.c_void_p(addr), ctypes.c_size_t(nbytes), 14)  # MADV_HUGEPAGE
            except Exception:
                pass
            err = torch.cuda.cudart().cudaHostRegister(addr, nbytes, 0)
            if int(err) != 0:
                raise RuntimeError("cudaHostRegister failed: %s" % (err,))
            arr = raw[off:off + nbytes]
            del raw
            t = torch.from_numpy(arr).view(dtype).reshape(shape)
            weakref.finalize(arr, _release_registered, addr, mm)
            return t
        except Exception:
            pass
    try:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    except RuntimeError:
        return torch.empty(shape, dtype=dtype)


def plan_window(T, nvar, nk, slab_bytes):
    rows = int(max(1, min(T, slab_bytes // max(1, nvar * nk * 16))))
    return rows


def first_order_host(prob, out=None, slab_bytes=1 << 30, ring=None, y0=None, n0=None, compact=True,
                     fill_threads=0, skip_unstarted=True, pr_out=None):
    """Trajectory delivered into host memory.  Time windows ping-pong between two device slabs;
    the device->host copy of window w overlaps the integration of window w+1.

    compact (default): the kernels write only the perturbation rows (PF_LAYOUT_PERT) and only
    those cross PCIe (2-D copies straight into rows [2nf+1, nvar) of the result); the
    background rows -- the same numbers for every mode -- and the NaN prefix are written into
    the result by host threads (pf_host_fill_background) from the 1 MB step-record table,
    concurrently with the copies.  For nf=1 that is 40 % of the bytes on the link.

    ``pr_out``: optional device tensor (T, nk) float64 that receives P_R of every row, reduced
    from each device window (K5) before the window is recycled -- ``model.Pr`` then needs no
    re-upload of the host trajectory.

    Returns a torch CPU tensor (T, nvar, nk) complex128 (pinned when it could be).
    ``ring``: a (R, nvar, nk) host tensor used instead of a full-size result when host memory
    cannot hold it -- every row still crosses to the host, windows overwrite each other."""
    import threading
    L = _lib.lib()
    T, nvar, nk, nf = prob.T, prob.nvar, prob.nk, prob.nf
    nb = 2 * nf + 1
    npert = nvar - nb
    compact = bool(compact and ring is None and prob.tsix_host is not None
                   and prob.ystart_host is not None)
    rows_per_t = npert if compact else nvar
    W = plan_window(T, rows_per_t, nk, slab_bytes)
    if ring is not None:
        W = max(1, min(W, ring.shape[0] // 2))
    elif out is None:
        out = _pinned_empty((T, nvar, nk), torch.complex128)
    elif (tuple(out.shape) != (T, nvar, nk) or out.dtype != torch.complex128 or not out.is_contiguous()
          or out.device.type != "cpu"):
        raise ValueError("output buffer must be a contiguous complex128 CPU tensor of shape %r"
                         % ((T, nvar, nk),))
    layout = _lib.PF_LAYOUT_PERT if compact else _lib.PF_LAYOUT_FULL
    # start rows sorted in k (the normal case): in every window the started modes are a prefix
    # of the columns, so only that prefix is copied; the host fill writes the NaNs of the rest
    sorted_starts = bool(compact and skip_unstarted and np.all(np.diff(prob.tsix_host) >= 0))
    filler = None
    fill_rc = []
    d2h_bytes = 0
    try:
        with torch.cuda.device(prob.dev):
            slabs = [torch.empty((W, rows_per_t, nk), dtype=torch.complex128, device=prob.dev)
                     for _ in range(2 if W < T else 1)]
            compute = torch.cuda.current_stream()
            copier = torch.cuda.Stream()
            copied = [None, None]
            w = 0
            for t0 in range(0, T, W):
                t1 = min(T, t0 + W)
                b = w % len(slabs)
                if copied[b] is not None:
                    compute.wait_event(copied[b])           # slab b is free again
                if t0 == 0 and y0 is not None:
                    prob.solve(y0, n0, t1, slabs[b], 1, compute, layout)   # K1 + K1b ride along
                else:
                    prob.launch(t0, t1, slabs[b], 1, compute, layout)
                if pr_out is not None:
                    for r0 in range(t0, t1, 32768):
                        r1 = min(t1, r0 + 32768)
                        if compact:
                            _lib.check(L.pf_pr_spectrum_pert(nf, r1 - r0, nk, _ptr(slabs[b][r0 - t0:r1 - t0]),
                                                             _ptr(prob.rec), prob.rec.shape[1],
                                                             int(L.pf_rec_bg_offset(nf)), r0, None,
                                                             _ptr(pr_out[r0:r1]), _stream_ptr(compute)))
                        else:
                            _lib.check(L.pf_pr_spectrum(nf, r1 - r0, nk, _ptr(slabs[b][r0 - t0:r1 - t0]), None,
                                                        _ptr(pr_out[r0:r1]), _stream_ptr(compute)))
                done = torch.cuda.Event()
                done.record(compute)
                copier.wait_event(done)
                with torch.cuda.stream(copier):
                    if compact:
                        # perturbation rows of the window, started columns only when the start
                        # rows are sorted in k, straight into rows [2nf+1, nvar) of the result:
                        # one 3-D copy per window
                        kcols = int(np.searchsorted(prob.tsix_host, t1 - 1, side="right")) \
                            if sorted_starts else nk
                        d2h_bytes += npert * kcols * 16 * (t1 - t0)
                        _lib.check(L.pf_copy_window_d2h(ctypes.c_void_p(out.data_ptr()), nvar, nk, nb, t0,
                                                        _ptr(slabs[b]), npert, kcols, t1 - t0,
                                                        _stream_ptr(copier)))
                    else:
                        d2h_bytes += nvar * nk * 16 * (t1 - t0)
                        dst = out[t0:t1] if ring is None else ring[(w % 2) * W:(w % 2) * W + (t1 - t0)]
                        dst.copy_(slabs[b][:t1 - t0], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copier)
                copied[b] = ev
                if compact and t0 == 0 and not os.environ.get("PF_PROBE_NO_FILL"):   # (probe switch: DMA alone)
                    # the step records (background rows included) are complete once window 0 is:
                    # bring them over and let the host threads write the background rows of every
                    # time step while the remaining windows compute and copy
                    rec_host = prob.rec.cpu().numpy()
                    d2h_bytes += rec_host.nbytes
                    bg_off = int(L.pf_rec_bg_offset(nf))
                    # share the host cores between the ranks of one box (torchrun sets LOCAL_WORLD_SIZE)
                    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
                    nthreads = fill_threads or max(1, len(os.sched_getaffinity(0)) // local_world)

                    def _fill():
                        rc = L.pf_host_fill_background(
                            ctypes.c_void_p(out.data_ptr()), nf, nk, 0, T,
                            ctypes.c_void_p(rec_host.ctypes.data), rec_host.shape[1], bg_off,
                            ctypes.c_void_p(prob.tsix_host.ctypes.data),
                            ctypes.c_void_p(prob.ystart_host.ctypes.data), 1 if sorted_starts else 0,
                            nthreads)
                        # the error text is thread-local in the library: read it on this thread
                        fill_rc.append((rc, L.pf_last_error().decode("utf-8", "replace") if rc else ""))
                    filler = threading.Thread(target=_fill)
                    filler.start()
                w += 1
            copier.synchronize()
            compute.synchronize()
    finally:
        # never leave the host threads streaming into `out` behind an exception
        if filler is not None:
            filler.join()
    if filler is not None:
        filler.join()
        rc, msg = fill_rc[0]
        if rc:
            raise (_lib.ArgumentError(rc, msg) if rc < 0 else _lib.LibraryError(msg))
    prob.d2h_bytes = d2h_bytes                  # what actually crossed PCIe towards the host
    prob.h2d_bytes = prob.k.numel() * 8 + prob.tsix.numel() * 8 + prob.ystart.numel() * 16
    return out if ring is None else ring


def pr_spectrum_device(nf, y, k=None):
    """K5 on rows already on the device: y (nrows, nvar, nk) c128 -> (nrows, nk) f64."""
    L = _lib.lib()
    nrows, nvar, nk = y.shape
    y = y.contiguous()
    out = torch.empty((nrows, nk), dtype=torch.float64, device=y.device)
    with torch.cuda.device(y.device):
        for r0 in range(0, nrows, 32768):
            r1 = min(nrows, r0 + 32768)
            _lib.check(L.pf_pr_spectrum(int(nf), r1 - r0, nk, _ptr(y[r0:r1]), _ptr(k),
                                        _ptr(out[r0:r1]), _stream_ptr()))
    return out


def potential_eval_device(model, y):
    """K0: (U [n], dU [n, nf], d2U [n, nf, nf]) for background vectors y [n, 2nf+1] (device)."""
    L = _lib.lib()
    nf = model.nfields
    y = y.contiguous()
    n = y.shape[0]
    U = torch.empty(n, dtype=torch.float64, device=y.device)
    dU = torch.empty((n, nf), dtype=torch.float64, device=y.device)
    d2U = torch.empty((n, nf, nf), dtype=torch.float64, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(L.pf_potential_eval(ctypes.byref(model), n, _ptr(y), _ptr(U), _ptr(dU), _ptr(d2U),
                                       _stream_ptr()))
    return U, dU, d2U


def device_initial_conditions(k, bg, T_infl, nf, ainit, etainit, cq, simtstart, h, phase=True,
                              suppress_ix=None):
    """K6 + K7: start rows and Bunch-Davies start vectors computed on the device.

    ``k`` device f64 [nk]; ``bg`` device [T, 2nf+1] (background run); ``T_infl`` = fotendindex
    (only rows before the end of inflation are searched, cosmomodels.py:1330).  Returns
    (rows int64 [nk], ystart c128 [nvar, nk]) on the device and raises like setfoics does."""
    from .cosmomodels import ModelError
    L = _lib.lib()
    dev = k.device
    nk = k.shape[0]
    with torch.cuda.device(dev):
        t = simtstart + torch.arange(T_infl, dtype=torch.float64, device=dev) * h
        aH = cq * (ainit * torch.exp(t) * bg[:T_infl, 2 * nf])
        cum = torch.cummax(aH, dim=0).values.contiguous()
        rows = torch.empty(nk, dtype=torch.int64, device=dev)
        nbad = torch.zeros(2, dtype=torch.int32, device=dev)
        _lib.check(L.pf_crossing_rows(nk, _ptr(k), int(T_infl), _ptr(cum), _ptr(rows), _ptr(nbad),
                                      _stream_ptr()))
        late, early = (int(v) for v in nbad.tolist())
        if late:
            raise ModelError("%d k modes cross the horizon after the end of inflation!" % late)
        if early and simtstart == 0:
            raise ModelError("Some k modes crossed horizon before simulation began and cannot "
                             "be initialized!")
        ystart = torch.empty((nvar_of(nf), nk), dtype=torch.complex128, device=dev)
        _lib.check(L.pf_bunch_davies(int(nf), nk, _ptr(k), _ptr(rows), _ptr(bg.contiguous()),
                                     float(simtstart), float(h), float(ainit), float(etainit),
                                     1 if phase else 0, -1 if suppress_ix is None else int(suppress_ix),
                                     _ptr(ystart), _stream_ptr()))
    return rows, ystart


def spectrum_run(potential_func, k, nfields=1, pot_params=None, bgystart=None, cq=50, tend=83.0,
                 h=0.01, simtstart=0.0, phase=True, suppress_ix=None, fixed_ainit=None,
                 row_stride="last", device=None):
    """Device-resident two-stage run for very large k grids (BASELINE config 5: 2^20..2^24
    modes): background run, horizon crossings, Bunch-Davies start vectors, the first-order
    integration and P_R(k) all stay on the GPU; only scalars and the requested rows leave it.

    Returns a dict: ``k``, ``fotstartindex``, ``rows`` (device c128 [R, nvar, nk], rows
    0, s, 2s, ... of yresult), ``tresult`` of those rows, ``Pr`` / ``scaled_Pr`` (device f64
    [R, nk]) and the scalars ``ainit, etainit, fotend, fotendindex, T``."""
    from . import cosmomodels as c
    dev = require_cuda(device)
    pot_params = {} if pot_params is None else pot_params
    nf = int(nfields)
    m = c.FOCanonicalTwoStage(k=np.array([1e-60]), potential_func=potential_func,
                              pot_params=pot_params, nfields=nf, bgystart=bgystart, cq=cq,
                              tend=tend, tstep_wanted=h, simtstart=simtstart)
    import time
    timing = {} if os.environ.get("PF_SPECTRUM_TIMING") else None

    def lap(name, t0):
        if timing is not None:
            torch.cuda.synchronize()
            timing[name] = timing.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    t0 = time.perf_counter()
    m.runbg()
    t0 = lap("runbg", t0)
    if fixed_ainit is None:
        ainit = float(np.ravel(m._choose_ainit())[0])
    else:
        ainit = float(fixed_ainit)
    bgy = m.bgmodel.yresult[:, :, 0]
    eps0 = 0.5 * np.sum(bgy[0, 1:2 * nf:2] ** 2)
    etainit = float(-1 / (ainit * bgy[0, 2 * nf] * (1 - eps0)))
    T = int(np.around((m.fotend - simtstart) / h) + 1)
    with torch.cuda.device(dev):
        kd = k.to(dev, torch.float64).contiguous() if torch.is_tensor(k) else \
            torch.as_tensor(np.ascontiguousarray(k, dtype=np.float64)).to(dev)
        bg = torch.as_tensor(np.ascontiguousarray(bgy)).to(dev)
        t0 = lap("h2d", t0)
        rows0, ystart = device_initial_conditions(kd, bg, int(m.fotendindex), nf, ainit, etainit, cq,
                                                  simtstart, h, phase, suppress_ix)
        t0 = lap("initial_conditions", t0)
        n0 = int(rows0.min().item())
        if int(rows0[0].item()) != n0:
            raise ValueError("k must be sorted so that column 0 starts first (cosmomodels.py:641)")
        pm = make_model(potential_func, pot_params, nf, ainit)
        prob = FirstOrderProblem(pm, kd, rows0, ystart, simtstart, T, h, dev)
        stride = max(1, T - 1) if row_stride == "last" else int(row_stride)
        y0 = bgy[n0].copy()
        t0 = lap("problem_setup", t0)
        out = first_order_device(prob, out_every=stride, y0=y0, n0=n0)
        prob.check_flags()
        t0 = lap("integration", t0)
        pr = pr_spectrum_device(nf, out)
        spr = pr_spectrum_device(nf, out, kd)
        t0 = lap("P_R", t0)
    tres = (simtstart + np.arange(T) * h)[::stride]
    return dict(k=kd, fotstartindex=rows0, rows=out, tresult=tres, Pr=pr, scaled_Pr=spr, ainit=ainit,
                etainit=etainit, fotend=m.fotend, fotendindex=int(m.fotendindex), T=T, timing=timing)


def fp64_peak_tflops(iters=100000, device=None):
    """Measured DFMA peak of the current GPU in TFLOP/s (pf_fp64_peak_probe)."""
    L = _lib.lib()
    dev = require_cuda(device)
    with torch.cuda.device(dev):
        n = 8 * 256 * torch.cuda.get_device_properties(dev).multi_processor_count
        scratch = torch.empty(n, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        out = ctypes.c_double(0.0)
        _lib.check(L.pf_fp64_peak_probe(int(iters), _ptr(scratch), n, ctypes.byref(out)))
    return float(out.value)


def derivs_device(model, y, t, k=None):
    """K8: dydx = derivs(y, t) on the device.  ``y`` array-like [rows] or [rows, ncol]; first
    order when ``k`` is given (rows = nvar), background otherwise (rows = 2nf+1).  Returns a
    NumPy array shaped and typed like ``y``."""
    L = _lib.lib()
    dev = require_cuda()
    ynp = np.asarray(y)
    single = ynp.ndim == 1
    y2 = ynp[:, None] if single else ynp
    rows, ncol = y2.shape
    nf = model.nfields
    want = nvar_of(nf) if k is not None else 2 * nf + 1
    if rows != want:
        raise ValueError("y must have %d rows" % want)
    with torch.cuda.device(dev):
        yd = torch.as_tensor(np.ascontiguousarray(y2, dtype=np.complex128)).to(dev)
        kd = None if k is None else torch.as_tensor(np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64)).to(dev)
        if kd is not None and kd.shape[0] != ncol:
            raise ValueError("k must have one entry per column of y")
        out = torch.empty_like(yd)
        _lib.check(L.pf_derivs(ctypes.byref(model), 1 if k is not None else 0, ncol, _ptr(yd), float(t),
                               _ptr(kd), _ptr(out), _stream_ptr()))
        res = out.cpu().numpy()
    if not np.iscomplexobj(ynp):
        res = np.ascontiguousarray(res.real)
    return res[:, 0] if single else res


# ---- second-order stage (K9) ------------------------------------------------------------------
def so_stage_coefficients(model, t, fo_bg0, bgepsilon, fo_simtstart, fo_tstep, so_tstep, device=None):
    """The time-only factors of the second-order right-hand side at the times ``t`` (any shape):
    array t.shape + (PF_SO_COEF,) of  3 - eps,  1/(aH)^2,  V''/H^2 - 3 phi'^2,  fotix (int64 bits),  t,  t/h
    (cosmomodels.py:724-755), with eps, phi', H, V'' read at row fotix of the first-order result
    (``fo_bg0`` [T_fo, 3]: column 0's phi, phi', H) / of ``bgepsilon``.  V'' is evaluated on the
    device (K0) for every first-order row once."""
    dev = require_cuda(device)
    t = np.asarray(t, dtype=np.float64)
    fo_bg0 = np.ascontiguousarray(np.real(fo_bg0), dtype=np.float64)
    T_fo = fo_bg0.shape[0]
    fotix = np.around((t - fo_simtstart) / fo_tstep).astype(np.int64)          # :724
    if fotix.size and (fotix.min() < 0 or fotix.max() >= min(T_fo, len(bgepsilon))):
        raise IndexError("second-order stage time outside the first-order result: row %d of %d"
                         % (int(fotix.max() if fotix.max() >= T_fo else fotix.min()), T_fo))
    with torch.cuda.device(dev):
        d2U = potential_eval_device(model, torch.as_tensor(fo_bg0).to(dev))[2][:, 0, 0].cpu().numpy()
    with np.errstate(all="ignore"):
        phidot, H = fo_bg0[fotix, 1], fo_bg0[fotix, 2]
        a = model.ainit * np.exp(t)                                             # :748
        coef = np.empty(t.shape + (_lib.PF_SO_COEF,), dtype=np.float64)
        coef[..., 0] = 3 - np.asarray(bgepsilon, dtype=np.float64)[fotix]
        coef[..., 1] = 1.0 / (a * H) ** 2
        coef[..., 2] = d2U[fotix] / H ** 2 - 3 * (phidot ** 2)
        coef[..., 3] = fotix.view(np.float64)                                   # int64 bits, not a double
        coef[..., 4] = t
        coef[..., 5] = t / so_tstep                                             # :935
    return coef


def _upload(a, dev):
    """Host array -> device tensor.  Read-only arrays (e.g. straight out of an .npz file) are
    uploaded as they are: the copy to the device never writes to them."""
    import warnings
    if a is None:
        return None
    with warnings.catch_warnings():
        warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
        return torch.as_tensor(np.ascontiguousarray(a)).to(dev)


def _ramp_array(ramp):
    if ramp is None:
        return None, None
    arr = (ctypes.c_double * 5)(*[float(ramp[q]) for q in ("a", "b", "c", "d", "e")])
    return arr, ctypes.cast(arr, ctypes.POINTER(ctypes.c_double))


class SecondOrderProblem(object):
    """Device-resident inputs of one second-order solve (what SOCanonicalThreeStage.runso hands to
    rkdriver_tsix, cosmomodels.py:1755-1790)."""

    def __init__(self, model, k, tsix, ystart, simtstart, T, h, fo_bg0, bgepsilon, fo_simtstart,
                 fo_tstep, fo_tsix=None, source=None, ramp=None, tstart=None, ramp_tsix=None, device=None,
                 stream_source_bytes=64 << 20):
        self.dev = require_cuda(device)
        self.model, self.T, self.h, self.simtstart = model, int(T), float(h), float(simtstart)
        k = np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64)
        self.nk = nk = k.shape[0]
        tsix = np.ascontiguousarray(np.atleast_1d(tsix), dtype=np.int64)
        ystart = np.ascontiguousarray(ystart, dtype=np.float64)
        if ystart.shape != (4, nk) or tsix.shape[0] != nk:
            raise ValueError("ystart must have shape (4, len(k))")
        # three stage times per step n -> n+1                       rk4.py:27-55, 116-118
        x = self.simtstart + np.arange(self.T) * self.h
        t = np.stack([x, x + self.h * 0.5, x + self.h], axis=1)
        t[self.T - 1] = t[max(self.T - 2, 0)]                       # the last row takes no step
        coef = so_stage_coefficients(model, t, fo_bg0, bgepsilon, fo_simtstart, fo_tstep, self.h, self.dev)
        self.h2d_bytes = 0
        with torch.cuda.device(self.dev):
            up = lambda a: _upload(a, self.dev)
            self.k, self.tsix, self.ystart, self.coef = up(k), up(tsix), up(ystart), up(coef)
            self.fo_tsix = None if fo_tsix is None else up(np.ascontiguousarray(fo_tsix, dtype=np.int64))
            self.tstart = self.ramp_tsix = None
            if ramp is not None:
                if tstart is None:
                    raise ValueError("a ramped source needs the first-order start times")
                self.tstart = up(np.ascontiguousarray(np.broadcast_to(np.asarray(tstart, dtype=np.float64), (nk,))))
                if ramp_tsix is not None:              # the model's start rows, when the solver's differ
                    self.ramp_tsix = up(np.ascontiguousarray(np.broadcast_to(np.asarray(ramp_tsix, dtype=np.int64), (nk,))))
            # the source term: resident on the device when it is given there or is small; a large
            # host array stays on the host and second_order_host streams it through time windows
            self.source = self.source_host = None
            self.fotix_host = np.ascontiguousarray(coef[..., 3]).view(np.int64)
            if source is not None:
                host = source if (torch.is_tensor(source) and source.device.type == "cpu") else None
                if not torch.is_tensor(source):
                    import warnings
                    with warnings.catch_warnings():       # a read-only array (np.load) is only read
                        warnings.filterwarnings("ignore", message="The given NumPy array is not writable")
                        host = torch.from_numpy(np.ascontiguousarray(np.asarray(source), dtype=np.complex128))
                src = host if host is not None else source
                if src.dtype != torch.complex128 or src.dim() != 2 or src.shape[1] != nk:
                    raise ValueError("source must be complex128 of shape (T_fo, len(k))")
                if int(self.fotix_host[: max(self.T - 1, 1)].max()) >= src.shape[0]:
                    raise IndexError("second-order stage time outside the source term")
                if host is not None and host.numel() * 16 >= stream_source_bytes:
                    self.source_host = host.contiguous()
                else:
                    if host is not None:
                        self.h2d_bytes += host.numel() * 16
                    self.source = src.to(self.dev).contiguous()
            self.state = torch.zeros((4, nk), dtype=torch.float64, device=self.dev)
        self._ramp_keep, self._ramp_ptr = _ramp_array(ramp)

    def launch(self, t0, t1, yout, out_every=1, stream=None, out_pitch=0, source_window=None):
        """K9 over rows [t0, t1).  ``source_window`` = (device tensor holding source rows
        [r0, r0 + n), r0): the streamed form of a host-resident source term."""
        L = _lib.lib()
        if source_window is not None:
            slab, r0 = source_window
            # the kernel indexes source_dev[fotix * pitch + k]: offset the base so that row r0 is slab[0]
            src_ptr, src_rows, pitch = slab.data_ptr() - int(r0) * self.nk * 16, int(r0) + slab.shape[0], self.nk
        elif self.source is not None:
            src_ptr, src_rows, pitch = self.source.data_ptr(), self.source.shape[0], self.source.shape[1]
        elif self.source_host is not None:                # not streamed: make it resident once
            with torch.cuda.device(self.dev):
                self.source = self.source_host.to(self.dev)
            self.h2d_bytes += self.source_host.numel() * 16
            self.source_host = None
            src_ptr, src_rows, pitch = self.source.data_ptr(), self.source.shape[0], self.source.shape[1]
        else:
            src_ptr, src_rows, pitch = 0, 0, 0
        with torch.cuda.device(self.dev):
            _lib.check(L.pf_so_run(self.nk, _ptr(self.k), _ptr(self.tsix), _ptr(self.fo_tsix),
                                   _ptr(self.ystart), _ptr(self.coef), ctypes.c_void_p(src_ptr), src_rows, pitch,
                                   _ptr(self.tstart), _ptr(self.ramp_tsix), self._ramp_ptr, self.T, self.h,
                                   int(t0), int(t1), _ptr(self.state), _ptr(yout), int(out_every),
                                   int(out_pitch), _stream_ptr(stream)))

    def source_rows(self, t0, t1):
        """[r0, r1): rows of the source term that the steps of result rows [t0, t1) read."""
        last = min(t1, self.T - 1)
        if last <= t0:
            return 0, 0
        f = self.fotix_host[t0:last]
        return int(f.min()), int(f.max()) + 1


def second_order_device(prob, out=None, out_every=1):
    """One launch over [0, T): rows 0, out_every, ... as a (rows, 4, nk) float64 CUDA tensor."""
    rows = (prob.T + out_every - 1) // out_every
    if out is None:
        out = torch.empty((rows, 4, prob.nk), dtype=torch.float64, device=prob.dev)
    prob.launch(0, prob.T, out, out_every)
    return out


def _threaded_copy(dst, src, pool, parts=4):
    """dst[...] = src for large host arrays, split over a few threads (NumPy releases the GIL)."""
    n = src.shape[0]
    if pool is None or n < 2 * parts:
        np.copyto(dst, src)
        return
    edges = [(n * i) // parts for i in range(parts + 1)]
    list(pool.map(lambda i: np.copyto(dst[edges[i]:edges[i + 1]], src[edges[i]:edges[i + 1]]), range(parts)))


def second_order_host(prob, slab_bytes=256 << 20, out=None):
    """The whole (T, 4, nk) float64 trajectory in (page-locked) host memory.  Time windows
    ping-pong through two device slabs: the device->host copy of window w overlaps the
    integration of window w+1 and -- when the source term lives on the host (a large NumPy array) --
    the host->device copy of the source rows of window w+2, staged through page-locked buffers, so
    both PCIe directions are busy at once instead of "upload everything, solve, download".
    State is carried on the device between windows; results are bit-identical to one launch."""
    from concurrent.futures import ThreadPoolExecutor
    T, nk = prob.T, prob.nk
    if out is None:
        out = _pinned_empty((T, 4, nk), torch.float64)
    elif (not torch.is_tensor(out) or tuple(out.shape) != (T, 4, nk) or out.dtype != torch.float64
          or not out.is_contiguous() or out.device.type != "cpu"):
        raise ValueError("output buffer must be a contiguous float64 CPU tensor of shape %r" % ((T, 4, nk),))
    W = max(1, min(T, slab_bytes // (32 * nk)))
    streaming = prob.source_host is not None
    prob.state.zero_()
    with torch.cuda.device(prob.dev):
        compute = torch.cuda.current_stream()
        d2h, h2d = torch.cuda.Stream(), torch.cuda.Stream()
        res = [torch.empty((W, 4, nk), dtype=torch.float64, device=prob.dev) for _ in range(2 if W < T else 1)]
        freed = [None] * len(res)                         # result slab b copied out
        src_dev = stage = None
        pool = None
        if streaming:
            nsrc = max(prob.source_rows(t0, min(T, t0 + W))[1] - prob.source_rows(t0, min(T, t0 + W))[0]
                       for t0 in range(0, T, W))
            nsrc = max(nsrc, 1)
            src_dev = [torch.empty((nsrc, nk), dtype=torch.complex128, device=prob.dev) for _ in range(2)]
            pinned_src = prob.source_host.is_pinned()
            if not pinned_src:
                stage = [_pinned_empty((nsrc, nk), torch.complex128) for _ in range(2)]
                pool = ThreadPoolExecutor(max_workers=4)
            uploaded = [None, None]                       # H2D into src_dev[b] done
            consumed = [None, None]                       # kernel that read src_dev[b] done
            src_np = prob.source_host.numpy()

            def upload(w):
                """start the host->device copy of window w's source rows into slab w % 2"""
                t0 = w * W
                if t0 >= T:
                    return
                b = w % 2
                r0, r1 = prob.source_rows(t0, min(T, t0 + W))
                if r1 <= r0:
                    uploaded[b] = (None, r0, 0)
                    return
                if consumed[b] is not None:
                    h2d.wait_event(consumed[b])
                if pinned_src:
                    host_rows = prob.source_host[r0:r1]
                else:
                    if uploaded[b] is not None and uploaded[b][0] is not None:
                        uploaded[b][0].synchronize()      # the staging buffer is free again
                    _threaded_copy(stage[b].numpy()[: r1 - r0], src_np[r0:r1], pool)
                    host_rows = stage[b][: r1 - r0]
                with torch.cuda.stream(h2d):
                    src_dev[b][: r1 - r0].copy_(host_rows, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(h2d)
                prob.h2d_bytes += (r1 - r0) * nk * 16
                uploaded[b] = (ev, r0, r1 - r0)
            upload(0)
            upload(1)
        try:
            for w, t0 in enumerate(range(0, T, W)):
                t1 = min(T, t0 + W)
                b = w % len(res)
                if freed[b] is not None:
                    compute.wait_event(freed[b])
                window = None
                if streaming:
                    ev, r0, n = uploaded[w % 2]
                    if ev is not None:
                        compute.wait_event(ev)
                        window = (src_dev[w % 2][:n], r0)
                    else:
                        window = (src_dev[w % 2][:1], r0)     # a window that takes no step reads nothing
                prob.launch(t0, t1, res[b], 1, compute, source_window=window)
                done = torch.cuda.Event()
                done.record(compute)
                if streaming:
                    consumed[w % 2] = done
                    upload(w + 2)
                d2h.wait_event(done)
                with torch.cuda.stream(d2h):
                    out[t0:t1].copy_(res[b][: t1 - t0], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(d2h)
                freed[b] = ev
            d2h.synchronize()
            compute.synchronize()
            h2d.synchronize()
        finally:
            if pool is not None:
                pool.shutdown(wait=True)
    prob.d2h_bytes = T * 4 * nk * 8
    return out.numpy()


def so_derivs_device(model, y, t, k, fo_bg0, bgepsilon, fo_simtstart, fo_tstep, so_tstep, source_row=None,
                     fo_tsix=None, ramp=None, tstart=None, tsix=None, fotix=None):
    """One second-order right-hand side on the device (pf_so_derivs): dydx (4, nk) for y (4, nk).
    ``fotix`` overrides the first-order row (CanonicalHomogeneousSecondOrder passes ``tix``)."""
    L = _lib.lib()
    dev = require_cuda()
    k = np.ascontiguousarray(np.atleast_1d(k), dtype=np.float64)
    nk = k.shape[0]
    y = np.ascontiguousarray(np.broadcast_to(np.asarray(y, dtype=np.float64).reshape(4, -1), (4, nk)))
    if fotix is None:
        coef = so_stage_coefficients(model, np.array([t]), fo_bg0, bgepsilon, fo_simtstart, fo_tstep, so_tstep, dev)[0]
    else:
        tt = fo_simtstart + fotix * fo_tstep
        coef = so_stage_coefficients(model, np.array([tt]), fo_bg0, bgepsilon, fo_simtstart, fo_tstep, so_tstep, dev)[0]
        a = model.ainit * np.exp(t)
        coef[1] = 1.0 / (a * np.real(fo_bg0[fotix, 2])) ** 2
        coef[4], coef[5] = t, t / so_tstep
    keep, rptr = _ramp_array(ramp)
    with torch.cuda.device(dev):
        up = lambda a: _upload(a, dev)
        kd, yd, cd = up(k), up(y), up(coef)
        sd = up(None if source_row is None else np.asarray(source_row, dtype=np.complex128))
        td = up(None if tstart is None else np.broadcast_to(np.asarray(tstart, dtype=np.float64), (nk,)))
        xd = up(None if tsix is None else np.asarray(tsix, dtype=np.int64))
        fd = up(None if fo_tsix is None else np.asarray(fo_tsix, dtype=np.int64))
        out = torch.empty_like(yd)
        _lib.check(L.pf_so_derivs(nk, _ptr(kd), _ptr(xd), _ptr(fd), _ptr(yd), _ptr(cd), _ptr(sd), _ptr(td),
                                  rptr, _ptr(out), _stream_ptr()))
        return out.cpu().numpy()

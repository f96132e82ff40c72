"""parallel - k-range sharding of a first-order run over the GPUs of one box.

The reference has no parallelism on this path (one Python process).  Modes
never interact (cosmomodels.py:663-674 couples only the fields of one k), so the B200 design
shards the k grid: one process per GPU, each owning a contiguous k-slab, no communication
during the integration (every rank recomputes the shared ~1 MB background table itself).  The
only exchange is the assembly of result rows -- a window of yresult, the last row, or P_R(k) --
into the reference layout [rows][nvar][nk] (k is the innermost axis, so rank slabs interleave row
by row).  Two forms: ``PeerAssembly`` / ``ShardedFirstOrder.run_assembled`` (the product path:
every rank's mode kernel stores straight into the assembling rank's buffer over NVLink, CUDA IPC
peer memory -- compute and gather are one kernel), and ``gather_rows`` (one NCCL all-gather plus
an on-GPU interleave; also the gloo path of the CPU tests).
"""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def mode_cost(tsix, T, full_output=True):
    """Relative cost of each mode: rows written (T, when the full trajectory is stored) plus
    RK4 steps actually integrated (T - 1 - tsix)."""
    tsix = np.asarray(tsix, dtype=np.int64)
    active = (T - 1 - tsix).astype(np.float64)
    return active + (float(T) if full_output else 0.0)


def partition_contiguous(cost, world):
    """Split range(len(cost)) into ``world`` contiguous slabs of nearly equal total cost.
    Returns ``bounds`` (world+1,) with slab r = [bounds[r], bounds[r+1])."""
    cost = np.asarray(cost, dtype=np.float64)
    n = cost.shape[0]
    if world < 1:
        raise ValueError("world must be >= 1")
    csum = np.concatenate(([0.0], np.cumsum(cost)))
    targets = csum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(csum, targets, side="left")
    # choose the boundary whose prefix sum is closest to the target
    lo = np.clip(inner - 1, 0, n)
    inner = np.where(np.abs(csum[lo] - targets) <= np.abs(csum[np.clip(inner, 0, n)] - targets), lo, inner)
    bounds = np.concatenate(([0], inner, [n])).astype(np.int64)
    return np.maximum.accumulate(bounds)


def equal_slabs(nk, world):
    """Contiguous slabs of (almost) equal size: what a weak-scaling run uses."""
    return (np.arange(world + 1, dtype=np.int64) * nk) // world


def my_slab(bounds, rank=None):
    rank = dist.get_rank() if rank is None else rank
    return slice(int(bounds[rank]), int(bounds[rank + 1]))


def gather_rows(rows_local, bounds, group=None):
    """All-gather per-rank row slabs [R][nvar][nk_r] into the global layout [R][nvar][nk].

    ``rows_local`` lives on this rank's GPU (NCCL) or on the CPU (gloo, used by the CPU tests of
    the host logic).  Slabs may have different widths: they are padded to the widest for the
    collective and interleaved afterwards (``pf_scatter_slab`` on the GPU)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = np.asarray(bounds, dtype=np.int64)
    widths = np.diff(bounds)
    R, nvar, nk_r = rows_local.shape
    if nk_r != int(widths[rank]):
        raise ValueError("local slab width %d does not match bounds (%d)" % (nk_r, widths[rank]))
    wmax = int(widths.max())
    nk = int(bounds[-1])
    dev = rows_local.device
    padded = rows_local
    if nk_r != wmax:
        padded = torch.zeros((R, nvar, wmax), dtype=rows_local.dtype, device=dev)
        padded[:, :, :nk_r] = rows_local
    padded = padded.contiguous()
    gathered = torch.empty((world,) + tuple(padded.shape), dtype=rows_local.dtype, device=dev)
    # the collective moves plain float64 words (NCCL and gloo have no complex dtype)
    send = torch.view_as_real(padded) if padded.is_complex() else padded
    recv = torch.view_as_real(gathered) if gathered.is_complex() else gathered
    if dev.type == "cuda":
        dist.all_gather_into_tensor(recv, send, group=group)      # one NCCL all-gather, no staging
    else:
        dist.all_gather([recv[r] for r in range(world)], send, group=group)
    full = torch.empty((R, nvar, nk), dtype=rows_local.dtype, device=dev)
    for r in range(world):
        w = int(widths[r])
        if w == 0:
            continue
        if dev.type == "cuda" and rows_local.dtype == torch.complex128:
            slab = gathered[r]
            if w != wmax:
                slab = slab[:, :, :w].contiguous()
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().pf_scatter_slab(
                    R * nvar, w, nk, int(bounds[r]), ctypes.c_void_p(slab.data_ptr()),
                    ctypes.c_void_p(full.data_ptr()),
                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            full[:, :, int(bounds[r]):int(bounds[r + 1])] = gathered[r][:, :, :w]
    return full


class _DeviceArray(object):
    """Raw device memory described through ``__cuda_array_interface__`` so that torch can wrap
    it (complex128, C order) without owning it."""

    def __init__(self, ptr, shape, owner):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": "<c16",
                                         "data": (int(ptr), False), "version": 3, "strides": None}
        self._owner = owner


class PeerAssembly(object):
    """The assembled result [rows][nvar][nk] of a k-sharded run, living on rank ``root`` and
    written DIRECTLY by every rank's mode kernel: each rank maps the root's buffer (CUDA IPC,
    ``pf_peer_*``) and passes ``&assembled[0][0][k0]`` with ``out_pitch = nk`` to
    ``pf_fo_solve`` / ``pf_fo_run``, so the k-gather rides on the kernel's own 128-bit stores
    over NVLink / NVSwitch, row by row as they are produced, instead of an all-gather plus an
    interleave pass after the integration (``gather_rows``).

    All ranks of the group must construct it (collective: one 64-byte broadcast)."""

    def __init__(self, rows, nvar, nk, root=0, group=None):
        L = _lib.lib()
        self.shape = (int(rows), int(nvar), int(nk))
        self.nbytes = int(rows) * int(nvar) * int(nk) * 16
        self.root = root
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.owner = self.rank == root
        self.base = None
        # the handle travels as a 64-byte tensor: on the GPU for NCCL, on the host for gloo
        on_host = self.world > 1 and dist.get_backend(group) == "gloo"
        dev = torch.device("cpu") if on_host else torch.device("cuda", torch.cuda.current_device())
        handle = torch.zeros(_lib.PF_PEER_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        if self.owner:
            p = ctypes.c_void_p()
            _lib.check(L.pf_peer_alloc(self.nbytes, ctypes.byref(p)))
            self.base = int(p.value)
            if self.world > 1:
                buf = ctypes.create_string_buffer(_lib.PF_PEER_HANDLE_BYTES)
                _lib.check(L.pf_peer_export(ctypes.c_void_p(self.base), buf))
                handle.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        if self.world > 1:
            dist.broadcast(handle, src=dist.get_global_rank(group, root) if group is not None else root,
                           group=group)
            if not self.owner:
                raw = handle.cpu().numpy().tobytes()
                p = ctypes.c_void_p()
                _lib.check(L.pf_peer_open(raw, ctypes.byref(p)))
                self.base = int(p.value)

    def column_ptr(self, k0):
        """Device address of assembled[0][0][k0] in this rank's address space."""
        return self.base + 16 * int(k0)

    def complete(self, stream=None):
        """Wait until every rank's stores have landed: stream sync on each writer, then a barrier."""
        (stream or torch.cuda.current_stream()).synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def tensor(self):
        """The assembled array as a CUDA tensor (root only; valid until ``close``)."""
        if not self.owner:
            raise RuntimeError("only the root rank holds the assembled result")
        return torch.as_tensor(_DeviceArray(self.base, self.shape, self), device="cuda")

    def close(self):
        if self.base is None:
            return
        L = _lib.lib()
        if self.world > 1:
            dist.barrier(group=self.group)               # nobody closes while a peer still writes
        if self.owner:
            _lib.check(L.pf_peer_free(ctypes.c_void_p(self.base)))
        else:
            _lib.check(L.pf_peer_close(ctypes.c_void_p(self.base)))
        self.base = None


def gather_spectrum(pr_local, bounds, group=None):
    """All-gather a per-mode quantity [R][nk_r] (e.g. P_R of the last row) into [R][nk]."""
    return gather_rows(pr_local.unsqueeze(1), bounds, group)[:, 0, :]


class ShardedFirstOrder(object):
    """One rank's share of a FOCanonicalTwoStage run over a global k grid.

    Every rank runs the (tiny) background model and the horizon-crossing search for the whole
    grid, so all ranks agree on the slab boundaries without communicating; the Bunch-Davies
    start vectors and the integration are done for the rank's own slab only."""

    def __init__(self, model_class=None, full_output=True, balance="cost", **kwargs):
        from . import cosmomodels
        cls = cosmomodels.FOCanonicalTwoStage if model_class is None else model_class
        self.kglobal = np.asarray(kwargs.pop("k"), dtype=np.float64)
        self.kwargs = kwargs
        self.cls = cls
        self.full_output = full_output
        self.balance = balance
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.model = None
        self.bounds = None

    def prepare(self):
        probe = self.cls(k=self.kglobal, **self.kwargs)
        probe.runbg()
        probe.ainit = probe._choose_ainit()
        rows = probe._crossing_rows(self.kglobal, probe.bgmodel.tresult[:probe.fotendindex],
                                    probe.bgmodel.yresult[:probe.fotendindex, probe.H_ix])
        T = int(np.around((probe.fotend - probe.simtstart) / probe.tstep_wanted) + 1)
        if self.balance == "cost":
            self.bounds = partition_contiguous(mode_cost(rows, T, self.full_output), self.world)
        else:
            self.bounds = equal_slabs(len(self.kglobal), self.world)
        sl = my_slab(self.bounds, self.rank)
        m = self.cls(k=self.kglobal[sl].copy(), **self.kwargs)
        m.bgmodel, m.fotend, m.fotendindex = probe.bgmodel, probe.fotend, probe.fotendindex
        m.setfoics()
        self.model = m
        return m

    def run(self):
        m = self.model if self.model is not None else self.prepare()
        m.runfo()
        return m

    def problem(self):
        """This rank's slab as a device problem (start vectors from the host IC glue, bit-identical
        to the reference's) plus column 0's background start (y0, n0)."""
        from . import engine
        m = self.model if self.model is not None else self.prepare()
        nf = m.nfields
        T = int(np.around((m.fotend - m.simtstart) / m.tstep_wanted) + 1)
        pm = engine.make_model(m.potential_func, m.pot_params, nf, m.ainit)
        prob = engine.FirstOrderProblem(pm, m.k, m.fotstartindex, m.foystart, m.simtstart, T,
                                        m.tstep_wanted)
        return prob, np.real(m.foystart[0:2 * nf + 1, 0]).copy(), int(m.fotstartindex.min())

    def run_assembled(self, out_every=1, root=0, fused=True, prob=None, assembly=None):
        """Integrate this rank's slab and assemble rows 0, s, 2s, ... of the GLOBAL yresult
        ([R][nvar][nk], k innermost -- rank slabs interleave row by row, rk4.py:95-100) on rank
        ``root``.

        fused=True : every rank's kernel stores straight into the root's buffer over NVLink
                     (PeerAssembly); returns (tensor on root / None elsewhere, assembly).
        fused=False: local slab, then NCCL all-gather + interleave (``gather_rows``); returns
                     (tensor on every rank, None)."""
        from . import engine
        if prob is None:
            prob, y0, n0 = self.problem()
        else:
            prob, y0, n0 = prob
        nk = int(self.bounds[-1])
        R = (prob.T + out_every - 1) // out_every
        if not fused:
            local = engine.first_order_device(prob, out_every=out_every, y0=y0, n0=n0)
            prob.check_flags()
            return (gather_rows(local, self.bounds) if self.world > 1 else local), None
        asm = assembly if assembly is not None else PeerAssembly(R, prob.nvar, nk, root=root)
        with torch.cuda.device(prob.dev):
            prob.solve(y0, n0, prob.T, asm.column_ptr(self.bounds[self.rank]), out_every,
                       out_pitch=nk)
        asm.complete()
        prob.check_flags()
        return (asm.tensor() if asm.owner else None), asm


def sharded_second_order(sharded, source, soclass=None, gather_every=None, **soclassargs):
    """Second-order stage (SOCanonicalThreeStage, cosmomodels.py:1680-1790) on this rank's k-slab
    of a finished sharded first-order run.  Modes never interact at second order either once the
    source term is given, so the slabs stay where they are and nothing is communicated.

    ``sharded``: a ShardedFirstOrder whose ``run()`` has completed.  ``source``: the source term of
    the GLOBAL grid, (T_fo, nk) complex128, or a callable ``slice -> (T_fo, nk_r) array`` that
    produces this rank's columns.  Returns the rank's driver after ``run()``; with
    ``gather_every = s`` also rows 0, s, 2s, ... of the global (T_so, 4, nk) result, assembled on
    every rank with one all-gather (``gather_rows``)."""
    from . import cosmomodels
    m = sharded.model
    if m is None or m.yresult is None:
        raise ValueError("run the sharded first-order stage first")
    sl = my_slab(sharded.bounds, sharded.rank)
    m.source = source(sl) if callable(source) else np.asarray(source)[:, sl]
    so = cosmomodels.SOCanonicalThreeStage(m, soclass=soclass, **soclassargs)
    so.run(saveresults=False)
    if gather_every is None:
        return so
    rows = np.ascontiguousarray(so.yresult[::int(gather_every)])
    local = torch.as_tensor(rows)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            local = local.to(torch.device("cuda", torch.cuda.current_device()))
        return so, gather_rows(local, sharded.bounds)
    return so, local


def to_host(t):
    """Device tensor -> NumPy array through a page-locked staging tensor (torch's caching pinned
    allocator keeps repeated calls cheap; a pageable ``.cpu()`` of a 2^25-mode spectrum costs
    more than the integration of a rank's slab)."""
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:
        host = torch.empty(t.shape, dtype=t.dtype)
    host.copy_(t)
    return host.numpy()


def sharded_spectrum_run(potential_func, k, gather=True, **kwargs):
    """BASELINE config 5: a k grid of 2^20..2^24 modes sharded over the ranks of the process
    group.  Each rank runs the device-resident pipeline (engine.spectrum_run) on its contiguous
    slab; the last-row spectra are assembled with one all-gather (``gather=False`` skips it).

    Returns the rank's result dict plus ``bounds`` and, when gathered, ``Pr_all`` /
    ``scaled_Pr_all`` (device f64 [R, nk]) on every rank."""
    from . import engine
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    k = np.asarray(k, dtype=np.float64)
    bounds = equal_slabs(len(k), world)      # state-only runs: cost ~ active steps, nearly uniform
    res = engine.spectrum_run(potential_func, k[my_slab(bounds, rank)], **kwargs)
    res["bounds"] = bounds
    if gather and world > 1:
        res["Pr_all"] = gather_spectrum(res["Pr"], bounds)
        res["scaled_Pr_all"] = gather_spectrum(res["scaled_Pr"], bounds)
    elif gather:
        res["Pr_all"], res["scaled_Pr_all"] = res["Pr"], res["scaled_Pr"]
    return res

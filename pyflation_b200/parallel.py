"""parallel - k-range sharding of a first-order run over the GPUs of one box.

The reference has no parallelism on this path (one Python process).  Modes
never interact (cosmomodels.py:663-674 couples only the fields of one k), so the B200 design
shards the k grid: one process per GPU, each owning a contiguous k-slab, no communication
during the integration (every rank recomputes the shared ~1 MB background table itself).  The
only collective is the assembly of result rows -- a window of yresult, the last row, or P_R(k)
-- with one NCCL all-gather over NVLink followed by an on-GPU interleave into the reference
layout [rows][nvar][nk] (k is the innermost axis, so rank slabs interleave row by row).
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

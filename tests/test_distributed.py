"""Host-side logic of the k-sharded multi-GPU path, on CPU with the gloo backend, world_size 2
(the N > 1 data path has no collective during the integration; the gather is the only one)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pyflation_b200 import parallel
    try:
        # a batch of 37 modes with start rows increasing in k, T = 100 rows, nvar = 5
        nk, T, nvar = 37, 100, 5
        tsix = (np.arange(nk) * 2 + 3).astype(np.int64)
        bounds = parallel.partition_contiguous(parallel.mode_cost(tsix, T), world)
        sl = parallel.my_slab(bounds)
        rng = np.random.RandomState(7)
        full = rng.standard_normal((4, nvar, nk)) + 1j * rng.standard_normal((4, nvar, nk))
        local = torch.as_tensor(full[:, :, sl].copy())
        got = parallel.gather_rows(local, bounds).numpy()
        ok_rows = np.array_equal(got, full)
        pr = torch.as_tensor(np.abs(full[:, 3, sl]) ** 2)
        ok_pr = np.array_equal(parallel.gather_spectrum(pr, bounds).numpy(), np.abs(full[:, 3, :]) ** 2)
        # second-order rows are real (R, 4, nk): the same gather, float64 (parallel.sharded_second_order)
        so_full = rng.standard_normal((3, 4, nk))
        so_full[0, :, tsix > 40] = np.nan
        got_so = parallel.gather_rows(torch.as_tensor(so_full[:, :, sl].copy()), bounds).numpy()
        ok_rows = ok_rows and np.array_equal(got_so, so_full, equal_nan=True)
        # every rank derives the same boundaries without communicating
        blist = [None] * world
        dist.all_gather_object(blist, bounds.tolist())
        out[rank] = (ok_rows, ok_pr, blist[0] == blist[1], bounds.tolist())
    finally:
        dist.destroy_process_group()


def test_sharding_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_rows, ok_pr, same, bounds = out[rank]
        assert ok_rows and ok_pr and same
        assert bounds[0] == 0 and bounds[-1] == 37 and 0 < bounds[1] < 37


def test_partition_balances_cost():
    from pyflation_b200 import parallel
    T = 8164
    tsix = np.sort(np.random.RandomState(1).randint(1546, 2239, size=2 ** 14))
    for world in (1, 2, 4, 8):
        b = parallel.partition_contiguous(parallel.mode_cost(tsix, T), world)
        assert b[0] == 0 and b[-1] == len(tsix) and np.all(np.diff(b) > 0)
        cost = parallel.mode_cost(tsix, T)
        per = np.array([cost[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.max() / per.mean() < 1.002
        # later start rows are cheaper: the last slab holds at least as many modes as the first
        assert (b[-1] - b[-2]) >= (b[1] - b[0])
    assert parallel.equal_slabs(10, 4).tolist() == [0, 2, 5, 7, 10]
    b = parallel.partition_contiguous(np.ones(3), 8)          # more ranks than modes: empty slabs
    assert b[0] == 0 and b[-1] == 3 and np.all(np.diff(b) >= 0)

"""Multi-rank assembly of a k-sharded run (tests/multigpu_worker.py under torchrun).

* ``test_fused_assembly_two_ranks`` always runs: on a single-GPU box the two ranks share cuda:0
  (gloo control plane) and the fused gather -- every rank's kernel storing its columns into
  rank 0's array through CUDA-IPC peer memory -- is exercised for real.
* ``test_sharded_run_and_nccl_gather`` needs >= 2 GPUs: NCCL all-gather path and fused path
  over NVLink, bit-compared with each other and checked against an unsharded run.
The gloo world-size-2 test in test_distributed.py covers the host logic on CPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _torchrun(world, port, *extra):
    # the worker processes share the GPU(s) with this pytest process, whose caching allocator may
    # still hold most of the device memory from earlier tests (full trajectories of 40-100 GB)
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "multigpu_worker.py")] + list(extra)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    if res.returncode != 0:                              # keep the whole log where gpurun brings it back
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "multigpu_worker_failure.log"), "w") as fh:
                fh.write(res.stdout + "\n==== stderr ====\n" + res.stderr)
        except OSError:
            pass
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-12000:]
    assert "OK" in res.stdout
    return res.stdout


def test_fused_assembly_two_ranks():
    import torch
    if torch.cuda.device_count() >= 2:
        out = _torchrun(2, 29534)
        assert "fused==nccl bits True" in out
    else:
        _torchrun(2, 29534, "--same-gpu")


def test_sharded_run_and_nccl_gather():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("NCCL needs one GPU per rank; the fused assembly is covered by "
                    "test_fused_assembly_two_ranks on this box")
    world = 2 if n < 4 else 4
    out = _torchrun(world, 29533)
    assert "fused==nccl bits True" in out

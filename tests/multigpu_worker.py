"""Worker of tests/test_multigpu.py: run under torchrun with >= 2 ranks, one GPU each.
Shards a first-order run over the ranks, gathers result rows with NCCL and checks them against
an unsharded run of the same grid done by rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from pyflation_b200 import cosmomodels as c, parallel
    k = np.linspace(0.85e-60, 8.2165e-58, 3001)
    kw = dict(potential_func="hybridquadratic", pot_params={"nfields": 2}, nfields=2,
              bgystart=np.array([12.0, 1 / 300.0, 12.0, 49 / 300.0, 0]), cq=50)
    sh = parallel.ShardedFirstOrder(k=k, **kw)
    m = sh.run()
    sl = parallel.my_slab(sh.bounds)
    assert m.yresult.shape[2] == sl.stop - sl.start
    rows = [100, 2000, m.yresult.shape[0] - 1]
    local = torch.as_tensor(np.ascontiguousarray(m.yresult[rows]), device="cuda")
    full = parallel.gather_rows(local, sh.bounds)                 # NCCL all-gather + pf_scatter_slab
    pr_local = torch.as_tensor(m.Pr[-1:], device="cuda")
    pr_full = parallel.gather_spectrum(pr_local, sh.bounds)
    ok = True
    if rank == 0:
        ref = c.FOCanonicalTwoStage(k=k, **kw)
        ref.run(saveresults=False)
        want = ref.yresult[rows]
        got = full.cpu().numpy()
        same_nan = np.array_equal(np.isnan(got), np.isnan(want))
        fin = ~np.isnan(want)
        scale = np.nanmax(np.abs(want), axis=1, keepdims=True) * np.ones_like(want.real)
        err = float((np.abs(got[fin] - want[fin]) / scale[fin]).max())
        err_pr = float(np.max(np.abs(pr_full.cpu().numpy() / ref.Pr[-1:] - 1)))
        ok = same_nan and err < 1e-11 and err_pr < 1e-11
        print("multigpu_check world=%d bounds=%s rows err=%.2e P_R err=%.2e %s"
              % (dist.get_world_size(), sh.bounds.tolist(), err, err_pr, "OK" if ok else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""Where does the wall time of parallel.sharded_spectrum_run go?  (development probe; run alone or
under torchrun: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ...)"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["PF_SPECTRUM_TIMING"] = "1"
import numpy as np, torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
from pyflation_b200 import parallel
nk = (int(sys.argv[1]) if len(sys.argv) > 1 else 2 ** 22) * world
k = np.linspace(0.85e-60, 8.2165e-58, nk)
kw = dict(nfields=1, pot_params={"nfields": 1}, bgystart=np.array([18.0, -0.1, 0.0]), cq=50)
for rep in range(4):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = parallel.sharded_spectrum_run("msqphisq", k, gather=False, **kw)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if world > 1:
        r["scaled_Pr_all"] = parallel.gather_spectrum(r["scaled_Pr"], r["bounds"])
    else:
        r["scaled_Pr_all"] = r["scaled_Pr"]
    torch.cuda.synchronize(); t2 = time.perf_counter()
    host = parallel.to_host(r["scaled_Pr_all"][-1])
    t3 = time.perf_counter()
    if rank == 0:
        print("rep %d nk=%d world=%d: local %.1f ms %s | gather %.1f ms | d2h %.1f ms | total %.1f ms" % (
            rep, nk, world, 1e3 * (t1 - t0), {a: round(1e3 * b, 1) for a, b in r["timing"].items()},
            1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
if world > 1: dist.destroy_process_group()

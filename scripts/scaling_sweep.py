"""BASELINE config 5: msqphisq scaling sweep over 2^20..2^24 k modes, sharded across the ranks
of a torchrun launch, with an NCCL gather of the final spectra.

    python scripts/scaling_sweep.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/scaling_sweep.py

Prints one JSON line per grid size (rank 0): wall time of the whole device-resident pipeline
(background, crossings, start vectors, integration, P_R, gather) and of the integration alone.
"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import torch.distributed as dist

def main():
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from pyflation_b200 import parallel
    sizes = [int(a) for a in sys.argv[1:]] or [2 ** 20, 2 ** 22, 2 ** 24]
    for nk in sizes:
        k = np.linspace(0.85e-60, 8.2165e-58, nk)
        for rep in range(2):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = parallel.sharded_spectrum_run("msqphisq", k, nfields=1, pot_params={"nfields": 1},
                                                bgystart=np.array([18.0, -0.1, 0.0]), cq=50, gather=False)
            torch.cuda.synchronize()
            t_local = time.perf_counter() - t0
            if world > 1:
                res["Pr_all"] = parallel.gather_spectrum(res["Pr"], res["bounds"])
                res["scaled_Pr_all"] = parallel.gather_spectrum(res["scaled_Pr"], res["bounds"])
            else:
                res["Pr_all"], res["scaled_Pr_all"] = res["Pr"], res["scaled_Pr"]
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            wall = time.perf_counter() - t0
        steps = float((res["T"] - 1 - res["fotstartindex"]).sum().item())
        tot = torch.tensor([steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        pr = res["Pr_all"][-1]
        if rank == 0:
            print(json.dumps({"nk": nk, "n_gpus": world, "wall_s": wall, "rank0_local_s": t_local, "mode_steps": float(tot[0]),
                              "pipeline_mode_steps_per_s": float(tot[0]) / wall,
                              "scaled_Pr_first_last": [float(res["scaled_Pr_all"][-1, 0]), float(res["scaled_Pr_all"][-1, -1])],
                              "finite": bool(torch.isfinite(pr).all())}))
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()

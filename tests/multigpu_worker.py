"""Worker of tests/test_multigpu.py: run under torchrun with >= 2 ranks.

Default: one GPU per rank, NCCL.  Shards a first-order run over the ranks and checks, against an
unsharded run of the same grid done by rank 0,
  * the NCCL path: local slabs -> all-gather + pf_scatter_slab (parallel.gather_rows), and
  * the fused path: every rank's kernel stores its columns straight into rank 0's assembled
    array through peer memory (parallel.PeerAssembly, pf_peer_*, out_pitch).
``--same-gpu``: every rank uses cuda:0 and the control plane is gloo (NCCL refuses two ranks on
one device); the fused path then runs for real through CUDA IPC on a single-GPU box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def compare(got, want):
    same_nan = np.array_equal(np.isnan(got), np.isnan(want))
    fin = ~np.isnan(want)
    scale = np.nanmax(np.abs(want), axis=1, keepdims=True) * np.ones_like(want.real)
    return same_nan, float((np.abs(got[fin] - want[fin]) / scale[fin]).max())


def main():
    same_gpu = "--same-gpu" in sys.argv
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(0 if same_gpu else int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    if same_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from pyflation_b200 import cosmomodels as c, parallel
    ok = True
    report = []
    for pot, nf, bgy, nk in (("hybridquadratic", 2, [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0], 3001),
                             ("msqphisq", 1, [18.0, -0.1, 0.0], 20011)):
        k = np.linspace(0.85e-60, 8.2165e-58, nk)
        kw = dict(potential_func=pot, pot_params={"nfields": nf}, nfields=nf, bgystart=np.array(bgy), cq=50)
        sh = parallel.ShardedFirstOrder(k=k, full_output=False, **kw)
        sh.prepare()
        stride = 97
        fused, asm = sh.run_assembled(out_every=stride, root=0, fused=True)
        gathered = None
        if not same_gpu:
            gathered, _ = sh.run_assembled(out_every=stride, fused=False)
            m = sh.run()
            pr_full = parallel.gather_spectrum(torch.as_tensor(m.Pr[-1:], device="cuda"), sh.bounds)
        if rank == 0:
            ref = c.FOCanonicalTwoStage(k=k, **kw)
            ref.run(saveresults=False)
            want = ref.yresult[::stride]
            nan_f, err_f = compare(fused.cpu().numpy(), want)
            good = nan_f and err_f < 1e-11
            line = "%s nf=%d nk=%d bounds=%s fused err=%.2e" % (pot, nf, nk, sh.bounds.tolist(), err_f)
            if gathered is not None:
                nan_g, err_g = compare(gathered.cpu().numpy(), want)
                err_pr = float(np.max(np.abs(pr_full.cpu().numpy() / ref.Pr[-1:] - 1)))
                # the two assembly paths run the same kernel on the same inputs: same bits
                bits = torch.equal(torch.view_as_real(fused).nan_to_num(0.0), torch.view_as_real(gathered).nan_to_num(0.0))
                good = good and nan_g and err_g < 1e-11 and err_pr < 1e-11 and bits
                line += " nccl err=%.2e P_R err=%.2e fused==nccl bits %s" % (err_g, err_pr, bits)
            report.append(line)
            ok = ok and good
        asm.close()
    # second-order stage on the slabs of a sharded first-order run (no communication until the
    # gather of a few result rows), against an unsharded run on rank 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyflation_oracle as oracle
    k = np.linspace(0.85e-60, 8.2165e-58, 1501)
    kw = dict(potential_func="msqphisq", pot_params={"nfields": 1}, nfields=1, bgystart=np.array([18.0, -0.1, 0.0]), cq=50)
    sh = parallel.ShardedFirstOrder(k=k, **kw)
    m = sh.run()
    so, rows = parallel.sharded_second_order(
        sh, lambda sl: oracle.synthetic_source(m.tresult, k[sl]), soclass=c.CanonicalRampedSecondOrder,
        gather_every=53)
    if rank == 0:
        ref = c.FOCanonicalTwoStage(k=k, **kw)
        ref.run(saveresults=False)
        ref.source = oracle.synthetic_source(ref.tresult, k)
        ref_so = c.SOCanonicalThreeStage(ref, soclass=c.CanonicalRampedSecondOrder)
        ref_so.run(saveresults=False)
        want = ref_so.yresult[::53]
        got = rows.cpu().numpy()
        # same kernel, same inputs per column: the sharded result is the unsharded one bit for bit
        bits = np.array_equal(got, want, equal_nan=True)
        report.append("second order nk=%d bounds=%s sharded==unsharded bits %s" % (len(k), sh.bounds.tolist(), bits))
        ok = ok and bits and bool(np.isfinite(want[-1]).all())
    if rank == 0:
        print("multigpu_check world=%d %s\n  " % (dist.get_world_size(), "same-gpu/gloo" if same_gpu else "nccl")
              + "\n  ".join(report) + "\n" + ("OK" if ok else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

"""bench.py -- headline benchmark of the first-order RK4 hot path (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): single-field lambdaphi4, 2^16 k-modes per GPU on
linspace(0.85e-60, 8.2165e-58, N*2^16), cq=50, h=0.01, full trajectory output (41 GB per GPU).
A "step" is one complete first-order integration of the rank's k-slab: background RK4 (K1),
stage table (K1b) and the mode kernel (K2) writing yresult[T][5][nk] into HBM.

  value  : k-mode*RK4-steps/s = sum_k (T-1-tsix_k) * K / (max over ranks of the device time of K
           steps), inputs resident in HBM, output left in HBM.
  e2e    : the same metric through the public API rk4.rkdriver_tsix(...) with host buffers:
           every step copies the start vectors host->device and delivers the full yresult into
           (pinned) host memory, device->host copies inside the timed region.
  roofline : the mode kernel K2 against the measured HBM copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference : the NumPy oracle port of the reference's rkdriver_tsix +
           CanonicalFirstOrder.derivs on the host cores, one process per core, each on a
           contiguous k-shard of a bounded sample of the same workload.
  configs  : (N = 1) every other BASELINE.json configuration, a few launches each after the
           headline loop: value, ms and its own roofline (HBM against MEASURED_PEAKS.json, fp64
           against a DFMA probe run in this process).
  strong / e2e_spectrum / collective : BASELINE configs[4] -- a fixed 2^24-mode msqphisq grid
           sharded over the N ranks (device-resident pipeline, P_R(k) gathered with NCCL); the
           same pipeline from a host k array to a host P_R(k) array, 2^22 modes per GPU; and,
           for N > 1, the assembly of a strided yresult of the 2^20-mode grid on rank 0 -- once
           as local slabs + NCCL all-gather + interleave, once with every rank's kernel storing
           its columns straight into rank 0's array over NVLink (PeerAssembly) -- both checked
           in-run against the reference's golden columns (tests/golden/fo_c5_msqphisq_2p20.npz).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KMIN, KMAX = 0.85e-60, 8.2165e-58
NK_PER_GPU = 2 ** 16
POTENTIAL, NFIELDS, BGYSTART = "lambdaphi4", 1, [25.0, 0.0, 0.0]
POT_PARAMS = {"nfields": 1}
H, TEND, CQ = 0.01, 83.0, 50
METRIC = "k-mode·RK4-steps/s (fp64)"
UNIT = "k-mode·RK4-steps/s"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, all host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """Integrate one contiguous k-shard with the NumPy oracle (rk4.py:58-138 restated)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import warnings
    warnings.filterwarnings("ignore")
    import pyflation_oracle as O
    k, tsix, y0, ainit, fotend = args
    t0 = time.perf_counter()
    _, y = O.rk4_drive(y0, 0.0, tsix, fotend, H, O.make_fo_rhs(POTENTIAL, POT_PARAMS, NFIELDS, k, ainit))
    dt = time.perf_counter() - t0
    T = y.shape[0]
    ok = bool(np.isfinite(y[-1, 3:, :]).all())
    return dt, float(np.sum(T - 1 - tsix)), ok


class CpuArm(object):
    """Bounded sample of the bench workload for the host cores: P shards of `nk_worker`
    consecutive modes of the 2^16 grid (each shard's column 0 is its own smallest k)."""

    def __init__(self, nk_worker, cores=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import warnings
        warnings.filterwarnings("ignore")
        import pyflation_oracle as O
        self.cores = cores or os.cpu_count() or 1
        self.nk_worker = int(nk_worker)
        nk = self.cores * self.nk_worker
        kfull = np.linspace(KMIN, KMAX, NK_PER_GPU)
        # spread the shards over the grid so the sample sees the whole range of start rows
        starts = np.linspace(0, NK_PER_GPU - self.nk_worker, self.cores).astype(int)
        k = np.concatenate([kfull[s:s + self.nk_worker] for s in starts])
        ic = O.first_order_run(POTENTIAL, k, NFIELDS, POT_PARAMS, np.array(BGYSTART), cq=CQ,
                               tend=TEND, h=H, integrate=False)
        self.jobs = []
        for w in range(self.cores):
            sl = slice(w * self.nk_worker, (w + 1) * self.nk_worker)
            self.jobs.append((k[sl], ic["fotstartindex"][sl], ic["foystart"][:, sl].copy(),
                              ic["ainit"], ic["fotend"]))
        self.sample = ("%d processes x %d consecutive modes of the %s 2^16 grid (%d modes), full "
                       "T=%d trajectory each, NumPy oracle port of rkdriver_tsix"
                       % (self.cores, self.nk_worker, POTENTIAL, nk,
                          O.number_of_steps(0.0, ic["fotend"], H)))
        import multiprocessing as mp
        self.pool = mp.get_context("fork").Pool(self.cores)

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, self.jobs, chunksize=1)
        wall = time.perf_counter() - t0
        assert all(r[2] for r in res), "oracle produced non-finite modes"
        return wall, sum(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = max(1, args.steps + args.warmup)
    nk_worker = int(min(2048, max(128, 2048 * 15 // total)))
    arm = CpuArm(nk_worker)
    for _ in range(args.warmup):
        arm.step()
    wall = 0.0
    work = 0.0
    for _ in range(args.steps):
        w, s = arm.step()
        wall += w
        work += s
    arm.close()
    value = work / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": "port",
                         "sample": arm.sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(ngpu):
    return {"workload": "lambdaphi4 first-order run, 2^16 k-modes per GPU, full trajectory "
                        "output (BASELINE.json configs[1])",
            "potential": POTENTIAL, "nfields": NFIELDS, "nk_per_gpu": NK_PER_GPU,
            "nk_total": NK_PER_GPU * ngpu, "k_range": [KMIN, KMAX], "cq": CQ, "h": H,
            "tend": TEND, "output": "full trajectory yresult[T][5][nk] complex128",
            "sharding": "contiguous k-slab per GPU, no data-path collective",
            "l2": "no flush needed: each step streams a 41 GB write-once output (>> 126 MB L2), "
                  "nothing is re-read"}


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index):
        threading.Thread.__init__(self, daemon=True)
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as err:             # NVML missing: report it, do not fail the bench
            self.err = repr(err)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [],
                    "note": "no NVML samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# the other BASELINE configurations (N = 1), config 5 (every N), the collective (N > 1)
# ------------------------------------------------------------------------------------------
FLOPS_PER_MODE_STEP = {1: 104, 2: 456, 8: 13320}          # nf^2 (80 + 16 nf) + 8


def _k4_grid():
    from pyflation_b200 import helpers
    return helpers.seq(KMIN, helpers.getkend(0.85e-60, 0.4e-60, 1025), 0.4e-60)


# name, potential, nfields, pot_params, bgystart, k grid, out_every ("last" = rows 0 and T-1)
EXTRA_CONFIGS = [
    ("configs[0] msqphisq, default 2053-mode grid, full trajectory", "msqphisq", 1, {"nfields": 1},
     [18.0, -0.1, 0.0], "K4", 1),
    ("configs[2] hybridquadratic nf=2, 2^16 modes, full trajectory (100 GB)", "hybridquadratic", 2,
     {"nfields": 2}, [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0.0], 2 ** 16, 1),
    ("configs[2] hybridquadratic nf=2, 2^20 modes, every 64th row", "hybridquadratic", 2,
     {"nfields": 2}, [12.0, 1 / 300.0, 12.0, 49 / 300.0, 0.0], 2 ** 20, 64),
    ("configs[3] nflation nf=8, 2^18 modes, every 512th row", "nflation", 8, {"nfields": 8}, None,
     2 ** 18, 512),
    ("configs[3] nflation nf=8, 2^13 modes, full trajectory (155 GB)", "nflation", 8, {"nfields": 8},
     None, 2 ** 13, 1),
    ("configs[4] msqphisq, 2^20 modes (one GPU's slab), last row only", "msqphisq", 1, {"nfields": 1},
     [18.0, -0.1, 0.0], 2 ** 20, "last"),
]


def _events(n):
    import torch
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def measure_extra_configs(hbm_peak, fp64_peak, reps=3, device_index=0):
    """One entry per remaining BASELINE configuration: whole first-order solve (K1 + K1b + mode
    kernel, the metric's definition) and the mode kernel alone, CUDA events, inputs in HBM."""
    import torch
    from pyflation_b200 import cosmomodels, engine
    out_list = []
    for name, pot, nf, params, bgy, grid, every in EXTRA_CONFIGS:
        try:
            k = _k4_grid() if grid == "K4" else np.linspace(KMIN, KMAX, grid)
            m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func=pot, pot_params=dict(params), nfields=nf,
                                                bgystart=None if bgy is None else np.array(bgy), cq=CQ,
                                                tend=TEND, tstep_wanted=H)
            m.runbg()
            m.setfoics()
            T = int(np.around(m.fotend / H) + 1)
            pm = engine.make_model(pot, params, nf, m.ainit)
            prob = engine.FirstOrderProblem(pm, k, m.fotstartindex, m.foystart, 0.0, T, H)
            n0 = int(m.fotstartindex.min())
            y0 = np.real(m.foystart[0:2 * nf + 1, 0]).copy()
            stride = max(1, T - 1) if every == "last" else int(every)
            nrows = (T + stride - 1) // stride
            nvar = prob.nvar
            out = torch.empty((nrows, nvar, len(k)), dtype=torch.complex128, device="cuda")
            mode_steps = float(np.sum(T - 1 - m.fotstartindex))
            # rows actually written after a mode's start (the NaN prefix is not algorithmic work)
            rows_t = np.arange(0, T, stride)
            written = float(np.sum(len(rows_t) - np.searchsorted(rows_t, m.fotstartindex, side="left")))
            for _ in range(1):
                prob.solve(y0, n0, T, out, stride)
            ev = _events(reps)
            for a, b in ev:
                a.record()
                prob.solve(y0, n0, T, out, stride)
                b.record()
            torch.cuda.synchronize()
            solve_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            # the mode kernel alone, records already built (what the three-launch path and the
            # stage-argument cache run); for nf <= 2 pf_fo_solve is ONE fused launch instead
            prob.build_tables(y0, n0)
            prob.launch(0, T, out, stride)
            torch.cuda.synchronize()
            sampler = ClockSampler(device_index)
            sampler.start()
            ev = _events(reps)
            for a, b in ev:
                a.record()
                prob.launch(0, T, out, stride)
                b.record()
            torch.cuda.synchronize()
            sampler.stop_flag = True
            sampler.join()
            kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
            prob.check_flags()
            last = out[nrows - 1:nrows] if (T - 1) % stride == 0 else None
            if last is not None:
                pr = engine.pr_spectrum_device(nf, last, prob.k)
                assert bool(torch.isfinite(pr).all()) and float(pr.min()) > 0
            best_ms = min(solve_ms, kernel_ms) if nf <= 2 else kernel_ms
            alg_bytes = written * 16 * nvar
            all_bytes = float(nrows) * nvar * len(k) * 16
            flops = mode_steps * FLOPS_PER_MODE_STEP[nf]
            f_hbm = alg_bytes / (best_ms * 1e-3) / 1e9 / hbm_peak
            f_fp = flops / (best_ms * 1e-3) / 1e12 / fp64_peak
            bound = "hbm" if f_hbm >= f_fp else "fp64"
            entry = {"workload": name, "nk": int(len(k)), "T": T, "out_every": every,
                     "value": mode_steps / (solve_ms * 1e-3), "ms": solve_ms,
                     "clocks": sampler.result(),
                     "mode_kernel_ms": kernel_ms, "mode_kernel_value": mode_steps / (kernel_ms * 1e-3),
                     "output_policy": "full trajectory" if stride == 1 else
                                      ("rows 0 and T-1" if every == "last" else "every %dth row" % stride),
                     "roofline": {"bound": bound, "frac": max(f_hbm, f_fp),
                                  "timed": "fused solve" if (nf <= 2 and solve_ms <= kernel_ms) else "mode kernel",
                                  "hbm": {"achieved": alg_bytes / (best_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                          "unit": "GB/s", "frac": f_hbm,
                                          "achieved_all_rows": all_bytes / (best_ms * 1e-3) / 1e9,
                                          "frac_all_rows": all_bytes / (best_ms * 1e-3) / 1e9 / hbm_peak},
                                  "fp64": {"achieved": flops / (best_ms * 1e-3) / 1e12, "peak": fp64_peak,
                                           "unit": "TFLOP/s", "frac": f_fp,
                                           "flops_per_mode_step": FLOPS_PER_MODE_STEP[nf]}}}
            if nf == 8:
                entry["roofline"]["fp64"]["note"] = (
                    "13320 flop per mode-step is SURVEY 8(d)'s dense formulation; the factored "
                    "nflation kernel executes 10656 (5328 fp64 instructions), so its pipe "
                    "utilisation is frac * 0.80")
            if True:
                # what model.run() does: the stage arguments runbg() left on the device -> K1b +
                # mode kernel only, K1's serial chain is not repeated
                hit = engine.cached_stage_args(pm, H, 0.0, n0, y0, T, prob.dev)
                if hit is not None:
                    prob.use_stage_args(*hit)
                    prob.solve(y0, n0, T, out, stride)
                    ev = _events(reps)
                    for a, b in ev:
                        a.record()
                        prob.use_stage_args(*hit)
                        prob.solve(y0, n0, T, out, stride)
                        b.record()
                    torch.cuda.synchronize()
                    cms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
                    entry["cached_background"] = {"ms": cms, "value": mode_steps / (cms * 1e-3),
                                                  "what": "K1b + mode kernel from the stage arguments of runbg()"}
            del out, prob
            torch.cuda.empty_cache()
            out_list.append(entry)
        except Exception as err:                               # a config must not sink the bench line
            out_list.append({"workload": name, "error": repr(err)})
            torch.cuda.empty_cache()
    return out_list


def measure_window_config(hbm_peak, fp64_peak, reps=3, device_index=0):
    """configs[3] as a windowed full-trajectory run sees it: nflation nf=8, all 2^18 modes, every
    row of a 128-row time window written (78 GB), state carried in from the rows before.  This is
    the pattern engine.first_order_host uses for trajectories that fit nowhere at once, and it
    fills the GPU evenly (the 2^13-mode full run is 2.3 waves of CTAs)."""
    import torch
    from pyflation_b200 import cosmomodels, engine
    name = "configs[3] nflation nf=8, 2^18 modes, rows [4096, 4224) of a windowed full trajectory (78 GB)"
    try:
        nf, nk, t0w, t1w = 8, 2 ** 18, 4096, 4224
        k = np.linspace(KMIN, KMAX, nk)
        m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func="nflation", pot_params={"nfields": nf},
                                            nfields=nf, bgystart=None, cq=CQ, tend=TEND, tstep_wanted=H)
        m.runbg()
        m.setfoics()
        T = int(np.around(m.fotend / H) + 1)
        assert int(m.fotstartindex.max()) < t0w
        pm = engine.make_model("nflation", {"nfields": nf}, nf, m.ainit)
        prob = engine.FirstOrderProblem(pm, k, m.fotstartindex, m.foystart, 0.0, T, H)
        prob.build_tables(np.real(m.foystart[0:2 * nf + 1, 0]).copy(), int(m.fotstartindex.min()))
        prob.launch(0, t0w, None, 0)                          # state at row t0w
        state0 = prob.state.clone()
        out = torch.empty((t1w - t0w, prob.nvar, nk), dtype=torch.complex128, device="cuda")
        prob.launch(t0w, t1w, out, 1)
        torch.cuda.synchronize()
        sampler = ClockSampler(device_index)
        sampler.start()
        ev = _events(reps)
        for a, b in ev:
            prob.state.copy_(state0)
            a.record()
            prob.launch(t0w, t1w, out, 1)
            b.record()
        torch.cuda.synchronize()
        sampler.stop_flag = True
        sampler.join()
        prob.check_flags()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        assert bool(torch.isfinite(torch.view_as_real(out[-1])).all())
        mode_steps = float((t1w - t0w) * nk)
        alg_bytes = mode_steps * 16 * prob.nvar
        flops = mode_steps * FLOPS_PER_MODE_STEP[nf]
        f_hbm = alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak
        f_fp = flops / (ms * 1e-3) / 1e12 / fp64_peak
        del out, prob, state0
        torch.cuda.empty_cache()
        return {"workload": name, "nk": nk, "T": T, "window": [t0w, t1w], "value": mode_steps / (ms * 1e-3),
                "ms": ms, "mode_kernel_ms": ms, "mode_kernel_value": mode_steps / (ms * 1e-3),
                "output_policy": "every row of the window", "clocks": sampler.result(),
                "roofline": {"bound": "hbm" if f_hbm >= f_fp else "fp64", "frac": max(f_hbm, f_fp),
                             "timed": "mode kernel",
                             "hbm": {"achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": f_hbm},
                             "fp64": {"achieved": flops / (ms * 1e-3) / 1e12, "peak": fp64_peak,
                                      "unit": "TFLOP/s", "frac": f_fp,
                                      "flops_per_mode_step": FLOPS_PER_MODE_STEP[nf]}}}
    except Exception as err:
        torch.cuda.empty_cache()
        return {"workload": name, "error": repr(err)}


def measure_second_order(hbm_peak, reps=3, device_index=0):
    """the second-order stage (K9, pf_so_run) on the headline grid -- lambdaphi4,
    2^16 modes, CanonicalRampedSecondOrder, full (T/2, 4, nk) float64 trajectory left in HBM, with a
    synthetic source term (T, nk) complex128 resident in HBM (computing it is out of scope).
    HBM-bound: 32 B of source read + 32 B of result written per mode and second-order step."""
    import torch
    from pyflation_b200 import cosmomodels, engine
    name = "second order (8f-4): lambdaphi4, 2^16 modes, ramped source, full trajectory (8.2 GB in, 8.2 GB out)"
    try:
        nk = 2 ** 16
        k = np.linspace(KMIN, KMAX, nk)
        m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func="lambdaphi4", pot_params={"nfields": 1}, nfields=1,
                                            bgystart=np.array([25.0, 0.0, 0.0]), cq=CQ, tend=TEND, tstep_wanted=H)
        m.runbg()
        m.setfoics()
        T_fo = int(np.around(m.fotend / H) + 1)
        h2 = 2 * H
        tsix = np.around(m.fotstartindex * 0.5 + h2 / 2).astype(np.int64)          # cosmomodels.py:1706-1709
        T = int(np.around((T_fo - 1) * H / h2) + 1)
        pm = engine.make_model("lambdaphi4", {"nfields": 1}, 1, m.ainit)
        tt = torch.arange(T_fo, dtype=torch.float64, device="cuda")[:, None] * H
        kk = torch.as_tensor(k, device="cuda")[None, :]
        amp = (kk / 1e-60) ** -1.5 * torch.exp(-0.08 * tt)
        src = torch.complex(amp * torch.cos(0.9 * tt + 3.0 * (kk / 1e-58)), amp * torch.sin(1.3 * tt + 0.5))
        del amp
        ramp = {"a": 15.0, "b": 0.3, "c": 1, "d": 2, "e": 1}
        prob = engine.SecondOrderProblem(pm, k, tsix, np.zeros((4, nk)), 0.0, T, h2,
                                         fo_bg0=m.bgmodel.yresult[:T_fo, :, 0], bgepsilon=m.bgepsilon,
                                         fo_simtstart=0.0, fo_tstep=H, fo_tsix=m.fotstartindex, source=src,
                                         ramp=ramp, tstart=m.fotstart)
        out = torch.empty((T, 4, nk), dtype=torch.float64, device="cuda")
        prob.launch(0, T, out, 1)
        torch.cuda.synchronize()
        sampler = ClockSampler(device_index)
        sampler.start()
        ev = _events(reps)
        for a, b in ev:
            a.record()
            prob.launch(0, T, out, 1)
            b.record()
        torch.cuda.synchronize()
        sampler.stop_flag = True
        sampler.join()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        last = out[T - 1]
        assert bool(torch.isfinite(last).all()) and float(last.abs().max()) > 0
        mode_steps = float(np.sum(T - 1 - tsix))
        alg_bytes = mode_steps * 64.0
        all_bytes = mode_steps * 32.0 + float(T) * nk * 32.0
        del out, prob
        # the reference's CPU path beside it (the cpu_baseline leg: NumPy port of rkdriver_tsix +
        # CanonicalRampedSecondOrder.derivs, one core, the first 256 modes of the same grid)
        cpu = None
        try:
            if os.path.join(ROOT, "oracle") not in sys.path:
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import pyflation_oracle as O
            ns = 256
            fo_y = np.full((T_fo, 5, ns), np.nan, dtype=complex)
            bg_rows = m.bgmodel.yresult[:T_fo, :, 0]
            for col in range(ns):
                fo_y[m.fotstartindex[col]:, 0:3, col] = bg_rows[m.fotstartindex[col]:]
            t_cpu = time.perf_counter()
            _, y_cpu, ts_cpu = O.second_order_run("lambdaphi4", {"nfields": 1}, k[:ns], m.ainit, np.arange(T_fo) * H,
                                                  fo_y, m.fotstart[:ns], m.fotstartindex[:ns], m.bgepsilon,
                                                  src[:, :ns].cpu().numpy(), H, 0.0, ramp=ramp)
            cpu_wall = time.perf_counter() - t_cpu
            assert np.array_equal(ts_cpu, tsix[:ns]) and np.isfinite(y_cpu[-1]).all()
            cpu = {"value": float(np.sum(T - 1 - tsix[:ns])) / cpu_wall, "unit": "k-mode second-order RK4-steps/s",
                   "cores": 1, "kind": "port", "wall_s": cpu_wall,
                   "sample": "first %d modes of the grid, full T=%d trajectory, NumPy oracle port" % (ns, T)}
            del fo_y, y_cpu
        except Exception as err:
            cpu = {"error": repr(err)}
        # end to end through the host path: source term in page-locked host memory in, trajectory in
        # page-locked host memory out, both streamed through time windows (engine.second_order_host)
        e2e = None
        try:
            src_host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            src_host.copy_(src)
            del src
            torch.cuda.empty_cache()
            eprob = engine.SecondOrderProblem(pm, k, tsix, np.zeros((4, nk)), 0.0, T, h2,
                                              fo_bg0=m.bgmodel.yresult[:T_fo, :, 0], bgepsilon=m.bgepsilon,
                                              fo_simtstart=0.0, fo_tstep=H, fo_tsix=m.fotstartindex,
                                              source=src_host, ramp=ramp, tstart=m.fotstart)
            host_out = torch.empty((T, 4, nk), dtype=torch.float64, pin_memory=True)
            engine.second_order_host(eprob, out=host_out)
            walls = []
            for _ in range(reps):
                eprob.h2d_bytes = 0
                torch.cuda.synchronize()
                t_start = time.perf_counter()
                engine.second_order_host(eprob, out=host_out)
                walls.append(time.perf_counter() - t_start)
            assert bool(torch.isfinite(host_out[T - 1]).all())
            wall = float(np.mean(walls))
            e2e = {"value": mode_steps / wall, "unit": "k-mode second-order RK4-steps/s", "ms_per_step": wall * 1e3,
                   "h2d_bytes_per_step": int(eprob.h2d_bytes), "d2h_bytes_per_step": int(eprob.d2h_bytes),
                   "pcie_GBs_each_way": [eprob.h2d_bytes / wall / 1e9, eprob.d2h_bytes / wall / 1e9],
                   "api": "engine.second_order_host(problem with a host-resident source, out=<pinned buffer>): "
                          "source rows in and result rows out streamed through time windows"}
            del eprob, host_out, src_host
        except Exception as err:
            e2e = {"error": repr(err)}
        try:                                               # give the 16 GB of pinned memory back
            torch._C._host_emptyCache()
        except Exception:
            pass
        torch.cuda.empty_cache()
        traffic, traffic_src = None, None
        try:                                               # profiler counters of this very launch
            cap = json.load(open(os.path.join(ROOT, "profiles", "r2c_so_v4_ncu_full.json")))
            traffic = float(cap["dram_bytes_per_launch"])
            traffic_src = "profiles/r2c_so_v4_ncu_full.json (ncu --set full capture of the same launch; static)"
        except Exception:
            pass
        return {"workload": name, "nk": nk, "T": T, "value": mode_steps / (ms * 1e-3), "ms": ms,
                "unit": "k-mode second-order RK4-steps/s", "mode_kernel_ms": ms, "e2e": e2e, "cpu_baseline": cpu,
                "output_policy": "full trajectory", "clocks": sampler.result(),
                "roofline": {"bound": "hbm", "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "timed": "mode kernel",
                             "hbm": {"achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                                     "bytes_per_mode_step": 64, "algorithmic_bytes": alg_bytes,
                                     "traffic": traffic, "traffic_source": traffic_src,
                                     "achieved_all_rows": all_bytes / (ms * 1e-3) / 1e9,
                                     "frac_all_rows": all_bytes / (ms * 1e-3) / 1e9 / hbm_peak}}}
    except Exception as err:
        torch.cuda.empty_cache()
        return {"workload": name, "error": repr(err)}


def measure_config5(world, rank, barrier):
    """BASELINE configs[4]: msqphisq, k-sharded, device-resident pipeline.
    strong        : fixed 2^24-mode grid over the N ranks, P_R(k) gathered with NCCL.
    e2e_spectrum  : host k array in -> host P_R(k) array out, 2^22 modes per GPU (weak)."""
    import torch
    import torch.distributed as dist
    from pyflation_b200 import parallel
    kw = dict(nfields=1, pot_params={"nfields": 1}, bgystart=np.array([18.0, -0.1, 0.0]), cq=CQ)
    res = {}
    for tag, nk in (("strong", 2 ** 24), ("e2e_spectrum", (2 ** 22) * world)):
        k = np.linspace(KMIN, KMAX, nk)
        wall, walls = None, []
        for rep in range(4):                                   # rep 0 warms allocator and NCCL; best of the rest
            barrier()
            t0 = time.perf_counter()
            r = parallel.sharded_spectrum_run("msqphisq", k, gather=True, **kw)
            if tag == "e2e_spectrum":
                pr_host = parallel.to_host(r["scaled_Pr_all"][-1])
            torch.cuda.synchronize()
            barrier()
            walls.append(time.perf_counter() - t0)
            wall = min(walls[1:]) if rep else walls[0]
        steps = torch.tensor([float((r["T"] - 1 - r["fotstartindex"]).sum().item())], dtype=torch.float64,
                             device="cuda")
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(steps)
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        spr = r["scaled_Pr_all"][-1]
        ok = bool(torch.isfinite(spr).all()) and float(spr.min()) > 0 and spr.shape[0] == nk
        res[tag] = {"nk": nk, "n_gpus": world, "wall_s": float(tw[0]), "value": float(steps[0]) / float(tw[0]),
                    "walls_rank0_s": walls,
                    "unit": UNIT, "scaled_Pr_first_last": [float(spr[0]), float(spr[-1])], "finite": ok,
                    "what": ("whole device-resident pipeline (background, crossings, Bunch-Davies start "
                             "vectors, integration, P_R, NCCL gather of P_R)" if tag == "strong" else
                             "host k array -> sharded_spectrum_run -> host scaled P_R(k) array")}
        if tag == "e2e_spectrum":
            res[tag]["h2d_bytes"] = nk * 8
            res[tag]["d2h_bytes"] = nk * 8
            res[tag]["scaling"] = "weak (2^22 modes per GPU)"
            assert pr_host.shape[0] == nk
        del r
        torch.cuda.empty_cache()
    return res


def measure_collective(world, rank, barrier):
    """Assembly of a strided yresult of the 2^20-mode msqphisq grid on rank 0 (k innermost, so
    rank slabs interleave row by row, rk4.py:95-100), timed on the device both ways and checked
    against the reference's golden columns."""
    import torch
    import torch.distributed as dist
    from pyflation_b200 import parallel
    g = np.load(os.path.join(ROOT, "tests", "golden", "fo_c5_msqphisq_2p20.npz"))
    nk, every = 2 ** 20, 40
    k = np.linspace(KMIN, KMAX, nk)
    assert np.array_equal(k[g["kix"]], g["k"])
    sh = parallel.ShardedFirstOrder(k=k, full_output=False, potential_func="msqphisq",
                                    pot_params={"nfields": 1}, nfields=1,
                                    bgystart=np.array([18.0, -0.1, 0.0]), cq=CQ)
    sh.prepare()
    pack = sh.problem()
    prob = pack[0]
    R = (prob.T + every - 1) // every
    nbytes = R * prob.nvar * nk * 16
    out = {"op": "k-gather of yresult[::%d] (msqphisq, 2^20 modes) onto rank 0" % every, "nranks": world,
           "bytes": nbytes, "rows": R}

    def check(full):
        sel = g["rows"][g["rows"] % every == 0]
        got = full[torch.as_tensor(sel // every, device=full.device)][:, :, torch.as_tensor(g["kix"], device=full.device)]
        got = got.cpu().numpy()
        want = g["yrows"][np.isin(g["rows"], sel)]
        # no exception here: a rank that left this function early would desynchronise the
        # collectives that follow; the verdict is recorded and raised after the JSON line is out
        if not np.array_equal(np.isnan(got), np.isnan(want)):
            return float("inf")
        fin = ~np.isnan(want)
        with np.errstate(invalid="ignore"):
            scale = np.nanmax(np.abs(want[:, 3:]), axis=1, keepdims=True) * np.ones_like(want.real)
        scale[:, :3] = np.abs(want[:, :3])
        return float((np.abs(got[fin] - want[fin]) / scale[fin]).max())

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t[0])

    # compute only (local slab, no assembly) -- the floor both paths are compared with
    from pyflation_b200 import engine
    engine.first_order_device(prob, out_every=every, y0=pack[1], n0=pack[2])
    _, ms_local = timed(lambda: engine.first_order_device(prob, out_every=every, y0=pack[1], n0=pack[2]))
    # (a) NCCL: local slab, all-gather of padded slabs, on-GPU interleave
    sh.run_assembled(out_every=every, fused=False, prob=pack)
    (full_a, _), ms_nccl = timed(lambda: sh.run_assembled(out_every=every, fused=False, prob=pack))
    if rank == 0:
        out["parity_nccl_max_rel_err"] = check(full_a)
    del full_a
    torch.cuda.empty_cache()
    # (b) fused: every rank's kernel stores its columns into rank 0's array over NVLink
    asm = parallel.PeerAssembly(R, prob.nvar, nk, root=0)
    sh.run_assembled(out_every=every, fused=True, prob=pack, assembly=asm)
    (full_b, _), ms_fused = timed(lambda: sh.run_assembled(out_every=every, fused=True, prob=pack, assembly=asm))
    if rank == 0:
        out["parity_fused_max_rel_err"] = check(full_b)
    del full_b
    asm.close()
    if rank == 0:
        out["parity_tolerance"] = 1e-9
        out["parity_ok"] = bool(out["parity_nccl_max_rel_err"] < 1e-9 and out["parity_fused_max_rel_err"] < 1e-9)
    moved = nbytes * (world - 1) / world                       # bytes that cross NVLink into rank 0
    out.update({"ms_compute_only": ms_local, "ms_nccl_allgather_path": ms_nccl, "ms_fused_peer_store_path": ms_fused,
                "nvlink_bytes_into_root": moved,
                "nccl_allgather_busbw_GBs": nbytes * (world - 1) / world / max(1e-9, (ms_nccl - ms_local) * 1e-3) / 1e9,
                "fused_inbound_GBs": moved / (ms_fused * 1e-3) / 1e9,
                "note": "all-gather path: every rank ends up with the full array (bus bandwidth = "
                        "(N-1)/N * bytes / (t_path - t_compute)); fused path: gather onto rank 0 only, "
                        "transfer overlapped with the integration row by row; NVLink 5 peak 900 GB/s "
                        "per direction per GPU"})
    return out


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # CPU baseline first (N=1 only), before this process touches CUDA: the workers are forked
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(nk_worker=1024)
        arm.step() if args.cpu_warmup else None
        wall, work = arm.step()
        arm.close()
        cpu_baseline = {"value": work / wall, "unit": UNIT, "cores": arm.cores, "kind": "port",
                        "sample": arm.sample, "wall_s": wall}

    if os.environ.get("NCCL_DEBUG") in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]             # keep stdout to the one JSON line
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pyflation_b200 import cosmomodels, engine, rk4

    # ---- synthetic inputs: this rank's contiguous k-slab of the global grid
    kglobal = np.linspace(KMIN, KMAX, NK_PER_GPU * world)
    k = kglobal[rank * NK_PER_GPU:(rank + 1) * NK_PER_GPU].copy()
    m = cosmomodels.FOCanonicalTwoStage(k=k, potential_func=POTENTIAL, pot_params=dict(POT_PARAMS),
                                        nfields=NFIELDS, bgystart=np.array(BGYSTART), cq=CQ,
                                        tend=TEND, tstep_wanted=H)
    m.runbg()
    m.setfoics()
    T = int(np.around((m.fotend - 0.0) / H) + 1)
    pm = engine.make_model(POTENTIAL, POT_PARAMS, NFIELDS, m.ainit)
    prob = engine.FirstOrderProblem(pm, k, m.fotstartindex, m.foystart, 0.0, T, H)
    n0 = int(m.fotstartindex.min())
    y0 = np.real(m.foystart[0:3, 0]).copy()
    nvar = prob.nvar
    mode_steps = float(np.sum(T - 1 - m.fotstartindex))
    out = torch.empty((T, nvar, NK_PER_GPU), dtype=torch.complex128, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(evs=None):
        if evs is not None:
            evs[0].record()
        prob.solve(y0, n0, T, out, 1)                   # ONE fused launch: K1 + K1b + K2
        if evs is not None:
            evs[1].record()

    for _ in range(args.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    k2_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for i in range(args.steps):
        one_step(k2_events[i])
    stop.record()
    barrier()
    sampler.stop_flag = True
    sampler.join()
    prob.check_flags()
    elapsed_ms = start.elapsed_time(stop)
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in k2_events]))
    t = torch.tensor([elapsed_ms, k2_ms], dtype=torch.float64, device="cuda")
    work = torch.tensor([mode_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    elapsed_ms, k2_ms = float(t[0]), float(t[1])
    total_work = float(work[0])
    value = total_work * args.steps / (elapsed_ms * 1e-3)

    # sanity: the last timed step produced a finite spectrum (work was really done)
    pr = engine.pr_spectrum_device(NFIELDS, out[T - 1:T], prob.k)
    assert bool(torch.isfinite(pr).all()) and float(pr.min()) > 0

    # ---- roofline of the dominant kernel (K2)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)" if "hbm_gbs" in peaks \
        else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg_bytes = mode_steps * 16 * nvar                        # 16*nvar B per mode-step
    all_rows_bytes = float(T) * nvar * NK_PER_GPU * 16        # incl. mandatory NaN / start rows
    # DRAM bytes per launch need profiler counters: they come from the committed ncu capture of
    # this very launch (scripts/ncu_round2.sh -> profiles/), not from this run
    traffic, traffic_src = None, None
    for name in ("r2_fused_headline_ncu_full.json", "k2_ncu_summary.json"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            traffic = prof.get("dram_bytes_per_launch")
            traffic_src = "profiles/" + name + " (ncu --set full capture of the same launch; static)"
            break
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "fo1_fused_kernel<lambdaphi4> (K1+K1b+K2 in one launch)",
                "achieved": alg_bytes / (k2_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": alg_bytes / (k2_ms * 1e-3) / 1e9 / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": k2_ms,
                "algorithmic_bytes": alg_bytes,
                "achieved_all_rows": all_rows_bytes / (k2_ms * 1e-3) / 1e9,
                "frac_all_rows": all_rows_bytes / (k2_ms * 1e-3) / 1e9 / peak,
                "note": "achieved counts 80 B per ACTIVE mode-step (SURVEY 8d); the output "
                        "contract also requires the NaN rows before each mode's start, which "
                        "achieved_all_rows includes (bytes the kernel must write per launch)"}
    del out
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, config 5, the collective
    extras = {}
    if not args.no_extras:
        try:
            fp64_peak = engine.fp64_peak_tflops()
            extras["fp64_peak_tflops"] = fp64_peak
            if world == 1:
                extras["configs"] = measure_extra_configs(peak, fp64_peak, device_index=local_rank)
                extras["configs"].append(measure_window_config(peak, fp64_peak, device_index=local_rank))
                extras["configs"].append(measure_second_order(peak, device_index=local_rank))
            extras.update(measure_config5(world, rank, barrier))
            if world > 1:
                extras["collective"] = measure_collective(world, rank, barrier)
        except Exception as err:                         # the headline line must still come out
            import traceback
            extras["extras_error"] = repr(err)
            traceback.print_exc(file=sys.stderr)
            torch.cuda.empty_cache()

    # ---- e2e through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, m, T, nvar, mode_steps, world, rank, barrier)
        except Exception as err:                             # the headline line must still come out
            import traceback
            traceback.print_exc(file=sys.stderr)
            e2e = {"error": repr(err)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world),
            "clocks": sampler.result(), "e2e": e2e, "gpu_launches": args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "T": T, "mode_steps_per_gpu_step": mode_steps,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
        coll = extras.get("collective")
        if coll is not None and coll.get("parity_ok") is False:
            raise SystemExit("bench: the assembled yresult differs from the reference's golden columns "
                             "(nccl %.3e, fused %.3e)" % (coll["parity_nccl_max_rel_err"],
                                                          coll["parity_fused_max_rel_err"]))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, m, T, nvar, mode_steps, world, rank, barrier):
    """rk4.rkdriver_tsix through the model's run(): host start vectors in, host yresult out."""
    import torch
    import torch.distributed as dist
    import psutil
    from pyflation_b200 import cosmomodels, engine
    need = T * nvar * NK_PER_GPU * 16
    local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    avail = psutil.virtual_memory().available / max(1, local)
    full = need < 0.6 * avail
    host_rows = T if full else max(1, int((2 << 30) // (nvar * NK_PER_GPU * 16)))
    # cudaHostAlloc memory: 3 % faster in the timed region than the huge-page registered buffer the
    # library allocates on its own for results this large (which trades that for a 3x faster,
    # one-off allocation: scripts/e2e_probe.py); the allocation is outside the timed region
    try:
        host = torch.empty((host_rows, nvar, NK_PER_GPU), dtype=torch.complex128, pin_memory=True)
    except RuntimeError:
        host = engine._pinned_empty((host_rows, nvar, NK_PER_GPU), torch.complex128)
    from pyflation_b200 import rk4
    ystart_host = torch.as_tensor(m.foystart).pin_memory().numpy()
    tsix_host = torch.as_tensor(np.ascontiguousarray(m.fotstartindex)).pin_memory().numpy()
    fo = cosmomodels.CanonicalFirstOrder(ystart=ystart_host, tstart=m.fotstart, simtstart=0.0,
                                         tstartindex=tsix_host, tend=m.fotend, tstep_wanted=H,
                                         k=m.k, ainit=m.ainit, potential_func=POTENTIAL,
                                         pot_params=dict(POT_PARAMS), nfields=NFIELDS)

    def step():
        if full:
            # the user-facing call: numpy start vectors in, numpy yresult (pinned buffer) out
            rk4.rkdriver_tsix(ystart=ystart_host, simtstart=0.0, tsix=tsix_host, tend=m.fotend,
                              h=H, derivs=fo.derivs, yresult_out=host)
        else:
            pm = engine.make_model(POTENTIAL, POT_PARAMS, NFIELDS, m.ainit)
            prob = engine.FirstOrderProblem(pm, m.k, tsix_host, ystart_host, 0.0, T, H)
            engine.first_order_host(prob, ring=host, y0=np.real(ystart_host[0:3, 0]),
                                    n0=int(tsix_host.min()))
            prob.check_flags()

    nsteps = max(1, min(args.steps, args.e2e_steps))
    step()                                               # warm-up (page faults of the pinned buffer)
    barrier()
    t0 = time.perf_counter()
    for _ in range(nsteps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    ww = torch.tensor([mode_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
    if full:
        last = host[T - 1].numpy()
        assert np.isfinite(last[3:]).all()
    # where the time goes (N = 1): the PCIe copies alone, the host fill alone
    breakdown = None
    if full and world == 1:
        import ctypes
        from pyflation_b200 import _lib
        pm = engine.make_model(POTENTIAL, POT_PARAMS, NFIELDS, m.ainit)
        y0b, n0b = np.real(ystart_host[0:3, 0]).copy(), int(tsix_host.min())
        os.environ["PF_PROBE_NO_FILL"] = "1"
        try:
            t0 = time.perf_counter()
            prob = engine.FirstOrderProblem(pm, m.k, tsix_host, ystart_host, 0.0, T, H)
            engine.first_order_host(prob, out=host, y0=y0b, n0=n0b)
            torch.cuda.synchronize()
            dma_ms = 1e3 * (time.perf_counter() - t0)
        finally:
            del os.environ["PF_PROBE_NO_FILL"]
        rec_host = prob.rec.cpu().numpy()
        L = _lib.lib()
        t0 = time.perf_counter()
        rc = L.pf_host_fill_background(ctypes.c_void_p(host.data_ptr()), NFIELDS, NK_PER_GPU, 0, T,
                                       ctypes.c_void_p(rec_host.ctypes.data), rec_host.shape[1],
                                       int(L.pf_rec_bg_offset(NFIELDS)), ctypes.c_void_p(tsix_host.ctypes.data),
                                       ctypes.c_void_p(ystart_host.ctypes.data), 1,
                                       len(os.sched_getaffinity(0)))
        fill_ms = 1e3 * (time.perf_counter() - t0)
        assert rc == 0
        d2h_b = float(prob.d2h_bytes)
        breakdown = {"pcie_copies_alone_ms": dma_ms, "pcie_GBs": d2h_b / dma_ms / 1e6,
                     "host_fill_alone_ms": fill_ms, "host_fill_GBs": (need - d2h_b) / fill_ms / 1e6,
                     "host_threads": len(os.sched_getaffinity(0)),
                     "note": "both together share the host memory system: the step time above is "
                             "neither's sum nor their max; N ranks of one box share the same host DRAM"}
    h2d = int(ystart_host.size * 16 + m.k.size * 8 + tsix_host.size * 8)
    moved = dict(rk4.last_transfer)
    # full layout: every row crosses PCIe; compact (default when the host can hold the result):
    # the 2 perturbation rows of every time step + the 1 MB step-record table cross, the 3
    # background rows are rebuilt on the host by pf_host_fill_background
    # (columns that have not started yet in a time window are skipped as well)
    d2h = int(need) if not full else int(moved["d2h_bytes"])
    return {"value": float(ww[0]) * nsteps / float(tt[0]), "unit": UNIT,
            "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
            "host_result_bytes_per_step": int(need) * world, "steps": nsteps,
            "ms_per_step": 1e3 * float(tt[0]) / nsteps, "breakdown": breakdown,
            "host_buffer": "full pinned yresult" if full else "pinned ring (host RAM too small for N full results)",
            "api": "rk4.rkdriver_tsix(ystart, simtstart, tsix, tend, h, derivs=CanonicalFirstOrder.derivs, "
                   "yresult_out=<pinned buffer>)" if full else "engine.first_order_host(ring=...)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3, dest="e2e_steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the other configs / config-5 / collective phases (profiling runs)")
    ap.add_argument("--cpu-warmup", action="store_true")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "ours":
        # asked for N GPUs outside torchrun: launch one rank per GPU ourselves
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29540"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

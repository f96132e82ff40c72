"""Turn an .ncu-rep capture (ncu --set full) into the small JSON summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r2_x_ncu_full.json ["note"]
"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_registers": "occupancy_limit_registers_blocks",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_smsp",
    "smsp__warps_eligible.avg.per_cycle_active": "warps_eligible_per_smsp",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "dram__bytes_write.sum": "dram_write", "dram__bytes_read.sum": "dram_read",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__cycles_elapsed.avg": "sm_cycles",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def main(rep, out, note=""):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
    res = {"source": rep.split("/")[-1], "kernel": d["Kernel Name"][0], "note": note}
    for k, name in KEYS.items():
        if k in d:
            v, u = d[k]
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if u in SCALE and name in ("duration", "dram_write", "dram_read"):
                x *= SCALE[u]
                name += "_s" if name == "duration" else "_bytes"
            res[name] = x
    stalls = {}
    for h, (v, u) in d.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v.replace(",", ""))
    res["stall_warp_cycles_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    if "duration_s" in res and "dram_write_bytes" in res:
        res["dram_write_GBs"] = res["dram_write_bytes"] / res["duration_s"] / 1e9
        res["dram_bytes_per_launch"] = res["dram_write_bytes"] + res.get("dram_read_bytes", 0.0)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res)[:400])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")

"""Time the second-order kernel (K9) the way bench.py does: python scripts/so_probe.py"""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
print(json.dumps(bench.measure_second_order(peak, reps=5), indent=1))

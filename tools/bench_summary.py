#!/usr/bin/env python3
"""Print one line per workload of a bench.py JSON line: step / tile / bin kernel times and the roofline fraction."""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    rows = [dict(d, workload=d["config"]["workload"], device_front_end=d.get("with_front_end_device", {}))] + d.get("secondary", [])
    print(path)
    for r in rows:
        rf, fe = r["roofline"], r.get("device_front_end") or {}
        print("  %-11s step %.4f ms  tile %.4f  bin %.4f  frac %.4f  fe-kernel %.3f  views->crc %.3f ms" % (
            r["workload"], r["ms_per_step"], rf["kernel_ms"], rf["setup_ms"], rf["frac"], fe.get("front_end_kernel_ms", 0), fe.get("ms_per_step", 0)))

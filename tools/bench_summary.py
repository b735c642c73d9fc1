#!/usr/bin/env python3
"""Print one line per workload of a bench.py JSON line: pass / tile / bin kernel times, roofline fraction, end-to-end."""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    rows = [dict(d, workload=d["config"]["workload"])] + d.get("secondary", [])
    print(path, "n_gpus", d["n_gpus"])
    for r in rows:
        rf, e = r["roofline"], r["e2e"]
        spread = [p["ms_per_pass"] for p in r.get("per_rank", [])]
        print("  %-11s %5d views  pass %.4f ms  tile %.4f  bin %.4f  frac %.4f | e2e %.3f ms [%d in flight; one: %.3f] (fe kernel %.3f) %.0f Mpix/s | value %.0f Mpix/s%s%s" % (
            r["workload"], r["config"]["views_per_gpu"], r["ms_per_pass"], rf["kernel_ms"], rf["bin_kernel_ms"], rf["frac"], e["ms_per_pass"],
            e.get("batches_in_flight", 1), e.get("one_batch_in_flight", e)["ms_per_pass"], e["front_end_kernel_ms"], e["value"], r["value"],
            "  ranks %.4f..%.4f" % (min(spread), max(spread)) if len(spread) > 1 else "",
            "  parity %s" % r["parity_sample"]["ok"] if "parity_sample" in r else ""))
    if "cpu_baseline" in d:
        c = d["cpu_baseline"]
        print("  cpu_baseline %.0f Mpix/s on %d cores (single thread %.0f), parity_ok %s" % (c["value"], c["cores"], c["single_thread_value"], c["parity_ok"]))
    if d.get("e2e_host_lists"):
        h = d["e2e_host_lists"]
        print("  e2e_host_lists %.3f ms/pass %.0f Mpix/s, host front-end %.3f s" % (h["ms_per_pass"], h["value"], h["host_front_end_s"]))

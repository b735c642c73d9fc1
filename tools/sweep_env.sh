#!/bin/bash
# A/B run-time knobs (run under gpurun): tools/sweep_env.sh "walk1280 walk320" "DRR_TILE_MAX_ROWS=400" "DRR_TILE_MAX_ROWS=272" ...
# prints bin / tile kernel ms per step and the roofline fraction for every (workload, setting)
WLS=$1; shift
for wl in $WLS; do
  for setting in "$@"; do
    env $setting python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
    python - "$wl" "$setting" <<'P'
import json, sys
d = json.load(open("/tmp/s.json")); r = d["roofline"]
print("%-10s %-40s bin %.4f tile %.4f ms frac %.4f pass %.4f e2e %.4f" % (sys.argv[1], sys.argv[2], r["bin_kernel_ms"], r["kernel_ms"], r["frac"], d["ms_per_pass"], d["e2e"]["ms_per_pass"]))
P
  done
done

#!/bin/bash
# A/B run-time knobs (run under gpurun): tools/sweep_env.sh walk1280 "DRR_TILE_SMEM_PAD_KB=0" "DRR_TILE_SMEM_PAD_KB=20" ...
wl=$1; shift
for kv in "$@"; do
  env $kv python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
  python - $wl "$kv" <<'PY'
import json, sys
d = json.loads(open("/tmp/s.json").read().strip().splitlines()[-1])
print("%-9s %-40s step %.4f ms value %.0f | bin %.4f tile %.4f ms frac %.4f e2e %.0f" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["value"], d["roofline"]["setup_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
done

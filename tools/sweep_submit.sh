#!/bin/bash
# How the pipelined drr_submit behaves with the number of upload chunks (run under gpurun).
for wl in walk320 walk1280; do
  for ch in 1 2 4 8 16; do
    DRR_SUBMIT_CHUNKS=$ch python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
    python - $wl $ch <<'PY'
import json, sys
d = json.loads(open("/tmp/s.json").read().strip().splitlines()[-1])
print("%-9s chunks=%-2s step %.4f ms (bin %.4f + tile %.4f) e2e %.4f ms = %.0f Mpix/s, h2d %.1f MB" % (sys.argv[1], sys.argv[2], d["ms_per_step"], d["roofline"]["setup_ms"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_bytes_per_step"] / 1e6))
PY
  done
done

#!/bin/bash
# A/B the tile kernel's geometry knobs (run under gpurun): DRR_TILE_COLS x DRR_TILE_LPG on the 320x200 and 1280x800 walks.
for wl in walk320 walk1280; do
  for tc in 16 32; do for lpg in 8 16 32; do
    DRR_KERNEL=tile DRR_TILE_COLS=$tc DRR_TILE_LPG=$lpg python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/s.json 2>/tmp/s.err || { tail -3 /tmp/s.err; continue; }
    python - $wl $tc $lpg <<'PY'
import json, sys
d = json.loads(open("/tmp/s.json").read().strip().splitlines()[-1])
print("%-9s TC=%s LPG=%-2s kernel %.4f ms setup %.4f ms frac %.4f e2e %.0f" % (sys.argv[1], sys.argv[2], sys.argv[3], d["roofline"]["kernel_ms"], d["roofline"]["setup_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
  done; done
done

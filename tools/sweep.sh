#!/bin/bash
# Rebuild the kernels with different compile-time knobs and time the two kernels on the 320x200 and 1280x800 walks
# (run under gpurun):  tools/sweep.sh "" "-DDRR_TILE_MIN_BLOCKS=5" ...
for knob in "$@"; do
  touch doom_rust_renderer_b200/csrc/drr_tile.cu
  make -s -C doom_rust_renderer_b200/csrc EXTRA="$knob" > /dev/null || continue
  regs=$(grep -A2 "drr_tile_kernelILi16ELi16ELb1" doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o "Used [0-9]* registers"; grep -A2 "drr_tile_kernelILi16ELi16ELb1" doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o "[0-9]* bytes spill stores")
  for wl in walk320 walk1280; do
    python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --secondary= > /tmp/sweep.json 2>/tmp/sweep.err || { tail -5 /tmp/sweep.err; continue; }
    python - "$knob" "$regs" $wl <<'PY'
import json, sys
d = json.loads(open("/tmp/sweep.json").read().strip().splitlines()[-1])
print("%-34s %-34s %-9s bin %.4f tile %.4f ms frac %.4f e2e %.0f" % (sys.argv[1], sys.argv[2], sys.argv[3], d["roofline"]["setup_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
  done
done
touch doom_rust_renderer_b200/csrc/drr_tile.cu
make -s -C doom_rust_renderer_b200/csrc > /dev/null

#!/bin/bash
# Rebuild the kernels with different compile-time knobs and time the march kernel on the default workload (run under gpurun).
set -e
for knob in "$@"; do
  touch doom_rust_renderer_b200/csrc/drr_kernels.cu doom_rust_renderer_b200/csrc/drr_api.cu
  make -s -C doom_rust_renderer_b200/csrc EXTRA="$knob" > /dev/null
  regs=$(grep -A3 "drr_march_kernelILb1" doom_rust_renderer_b200/csrc/build/ptxas_kernels.log | grep -o "Used [0-9]* registers")
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --secondary=walk1280 > /tmp/sweep.json 2>/tmp/sweep.err || { tail -5 /tmp/sweep.err; continue; }
  python - "$knob" "$regs" <<'PY'
import json, sys
d = json.load(open("/tmp/sweep.json"))
s = d["secondary"][0]
print("%-40s %-20s walk320 march %.4f ms frac %.4f | walk1280 march %.4f ms frac %.4f" % (sys.argv[1], sys.argv[2], d["roofline"]["kernel_ms"], d["roofline"]["frac"], s["roofline"]["kernel_ms"], s["roofline"]["frac"]))
PY
done
touch doom_rust_renderer_b200/csrc/drr_kernels.cu doom_rust_renderer_b200/csrc/drr_api.cu
make -s -C doom_rust_renderer_b200/csrc > /dev/null

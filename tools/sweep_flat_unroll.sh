#!/bin/bash
# A/B: flat span loop unrolled twice (-DDRR_FLAT_UNROLL2) -- run under gpurun
for flag in "" "-DDRR_FLAT_UNROLL2"; do
  make -s -C doom_rust_renderer_b200/csrc clean; make -s -C doom_rust_renderer_b200/csrc EXTRA="$flag" > /dev/null 2>&1
  grep -A2 'drr_tile_kernelILi32ELi8ELi256ELi1ELb1' doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' '
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary walk1280,flats1280 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('[$flag] walk320 tile %.4f' % d['roofline']['kernel_ms'], ' '.join('%s tile %.4f' % (s['workload'], s['roofline']['kernel_ms']) for s in d['secondary']))"
done

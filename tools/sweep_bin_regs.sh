#!/bin/bash
# A/B: register cap of drr_bin_kernel (-DDRR_BIN_MIN_BLOCKS: 1 = 64 registers, 3 = 55, 4 = 40, 5 = 32) -- run under gpurun.
for mb in 1 3 4 5; do
  make -s -C doom_rust_renderer_b200/csrc clean; make -s -C doom_rust_renderer_b200/csrc EXTRA=-DDRR_BIN_MIN_BLOCKS=$mb > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary walk1280,stress1920 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('mb=$mb walk320 bin %.4f tile %.4f step %.4f' % (d['roofline']['setup_ms'], d['roofline']['kernel_ms'], d['ms_per_step']), ' '.join('%s bin %.4f step %.4f' % (s['workload'], s['roofline']['setup_ms'], s['ms_per_step']) for s in d['secondary']))"
done

#!/bin/bash
for mb in 6 7; do
  make -s -C doom_rust_renderer_b200/csrc clean; make -s -C doom_rust_renderer_b200/csrc EXTRA=-DDRR_TILE_MIN_BLOCKS_SHORT=$mb > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary things640 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('min blocks $mb: walk320 tile %.4f' % d['roofline']['kernel_ms'], ' '.join('%s tile %.4f' % (s['workload'], s['roofline']['kernel_ms']) for s in d['secondary']))"
done

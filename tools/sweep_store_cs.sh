#!/bin/bash
# A/B: framebuffer stores as streaming stores (st.global.cs, -DDRR_STORE_CS) -- run under gpurun
for flag in "" "-DDRR_STORE_CS"; do
  make -s -C doom_rust_renderer_b200/csrc clean; make -s -C doom_rust_renderer_b200/csrc EXTRA="$flag" > /dev/null 2>&1
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --secondary walk1280,things640,stress1920 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('[$flag] walk320 tile %.4f bin %.4f' % (d['roofline']['kernel_ms'], d['roofline']['setup_ms']), ' '.join('%s tile %.4f' % (s['workload'], s['roofline']['kernel_ms']) for s in d['secondary']))"
done

#!/bin/bash
# Where does the fixed cost of a tile (clear + write-out, no spans) go?  empty 1280x800 frames, 16- and 32-column tiles.
for tc in 16 32; do
  CMD="python bench.py --workload empty1280 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
  DRR_TILE_COLS=$tc $CMD > gpurun_out/empty_plain_$tc.log 2>&1 && DRR_TILE_COLS=$tc ncu --set full --clock-control none -k regex:drr_tile_kernel -s 3 -c 1 -o gpurun_out/prof_empty1280_tc$tc $CMD > gpurun_out/empty_ncu_$tc.log 2>&1
  grep -o '"kernel_ms": [0-9.]*' gpurun_out/empty_plain_$tc.log
done

#!/bin/bash
# A/B ncu captures of the tile kernel (run under gpurun): tools/profile_ab.sh <workload> <views> <tag> "<flags>" [<tag> "<flags>" ...]
WL=$1; VIEWS=$2; shift 2
while [ $# -gt 0 ]; do
  TAG=$1; FLAGS=$2; shift 2
  touch doom_rust_renderer_b200/csrc/drr_tile.cu
  make -s -j4 -C doom_rust_renderer_b200/csrc EXTRA="$FLAGS" > /dev/null 2>&1 || { echo "build failed: $FLAGS"; continue; }
  CMD="python bench.py --workload $WL --views $VIEWS --steps 2 --warmup 3 --no-cpu-baseline --secondary="
  $CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_tile -s 4 -c 1 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_ncu.log
done
touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -j4 -C doom_rust_renderer_b200/csrc > /dev/null 2>&1

#!/bin/bash
# A/B compile-time flags of the tile kernel (run under gpurun): tools/sweep_build.sh "walk320 things640" "" "-DDRR_TILE_MINB_SMALL=5" ...
WLS=$1; shift
for flags in "$@"; do
  touch doom_rust_renderer_b200/csrc/drr_tile.cu
  make -s -j4 -C doom_rust_renderer_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  grep -A2 "tile_kernelILi[0-9]ELb1" doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
  ./tools/sweep_env.sh "$WLS" "FLAGS=${flags:-default}"
done
touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -j4 -C doom_rust_renderer_b200/csrc > /dev/null 2>&1

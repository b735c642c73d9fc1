#!/bin/bash
# A/B compile-time flags that reach every CUDA source (run under gpurun): tools/sweep_build_all.sh "walk320 walk1280" "" "-DDRR_PAL16" ...
# TESTS=1 also runs the GPU parity suite with every build
WLS=$1; shift
for flags in "$@"; do
  touch doom_rust_renderer_b200/csrc/drr_kernels.h
  make -s -j4 -C doom_rust_renderer_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  grep -A2 "tile_kernelILi[0-9]ELb1" doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
  [ -n "$TESTS" ] && python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  ./tools/sweep_env.sh "$WLS" "FLAGS=${flags:-default}"
done
touch doom_rust_renderer_b200/csrc/drr_kernels.h; make -s -j4 -C doom_rust_renderer_b200/csrc > /dev/null 2>&1

#!/bin/bash
# A/B compile-time flags of the front-end (run under gpurun): tools/sweep_fe_flags.sh "walk320 things640" "" "-DDRR_FE_WARPS=8 -DDRR_FE_MIN_BLOCKS=2" ...
WLS=$1; shift
for flags in "$@"; do
  touch doom_rust_renderer_b200/csrc/drr_kernels.h
  make -s -j4 -C doom_rust_renderer_b200/csrc EXTRA="$flags" > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  grep -A3 "frontend_kernelILb1" doom_rust_renderer_b200/csrc/build/ptxas_frontend.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo
  ./tools/fe_time.sh "$WLS" "${flags:-default}"
done
touch doom_rust_renderer_b200/csrc/drr_kernels.h; make -s -j4 -C doom_rust_renderer_b200/csrc > /dev/null 2>&1

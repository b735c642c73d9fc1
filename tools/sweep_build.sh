#!/bin/bash
# Rebuild the whole library with different compile-time knobs and time the draw path on several workloads (run under gpurun):
#   tools/sweep_build.sh "walk320 walk1280" "" "-DDRR_PAL8"
wls=$1; shift
for knob in "$@"; do
  make -s -C doom_rust_renderer_b200/csrc clean > /dev/null; make -s -C doom_rust_renderer_b200/csrc EXTRA="$knob" > /dev/null || continue
  for wl in $wls; do ./tools/sweep_env.sh $wl "KNOB=[$knob]" | sed "s/^/$knob /"; done
done
make -s -C doom_rust_renderer_b200/csrc clean > /dev/null; make -s -C doom_rust_renderer_b200/csrc > /dev/null

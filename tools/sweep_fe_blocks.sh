#!/bin/bash
# A/B resident CTAs per SM (register cap) of the front-end kernel, two warps per CTA (run under gpurun): tools/sweep_fe_blocks.sh "walk320 things640" 12 20 24
WLS=$1; shift
for b in "$@"; do
  ./tools/sweep_fe_flags.sh "$WLS" "-DDRR_FE_WARPS=2 -DDRR_FE_MIN_BLOCKS=$b -DDRR_FE_MIN_BLOCKS_GLOBAL=$b"
done

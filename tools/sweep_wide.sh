#!/bin/bash
# Wide-and-short tiles at 1280x800 now that row bands have their own span lists (run under gpurun)
for mb in 6 4; do
  touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -C doom_rust_renderer_b200/csrc EXTRA="-DDRR_TILE_MIN_BLOCKS_SHORT=$mb" > /dev/null
  echo "== MIN_BLOCKS_SHORT=$mb"
  ./tools/sweep_env.sh walk1280 "DRR_TILE_MAX_ROWS=400 DRR_TILE_COLS=32 DRR_TILE_LPG=8" "DRR_TILE_MAX_ROWS=400 DRR_TILE_COLS=32 DRR_TILE_LPG=16" "DRR_TILE_MAX_ROWS=270 DRR_TILE_COLS=32 DRR_TILE_LPG=8" "DRR_TILE_MAX_ROWS=270 DRR_TILE_COLS=32 DRR_TILE_LPG=16" "DRR_TILE_MAX_ROWS=200 DRR_TILE_COLS=32 DRR_TILE_LPG=16"
  ./tools/sweep_env.sh stress1920 "DRR_TILE_MAX_ROWS=400 DRR_TILE_COLS=32 DRR_TILE_LPG=8" "DRR_TILE_MAX_ROWS=400 DRR_TILE_COLS=32 DRR_TILE_LPG=16" "DRR_TILE_MAX_ROWS=300 DRR_TILE_COLS=32 DRR_TILE_LPG=16"
  ./tools/sweep_env.sh things640 "DRR_TILE_MAX_ROWS=200" "A=0"
done
touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -C doom_rust_renderer_b200/csrc > /dev/null

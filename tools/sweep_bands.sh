for mb in 4 5 6; do
  touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -C doom_rust_renderer_b200/csrc EXTRA="-DDRR_TILE_MIN_BLOCKS=$mb" > /dev/null
  echo "== MIN_BLOCKS=$mb $(grep -A2 'drr_tile_kernelILi16ELi16ELb1' doom_rust_renderer_b200/csrc/build/ptxas_tile.log | grep -o 'Used [0-9]* registers')"
  ./tools/sweep_env.sh walk1280 DRR_TILE_MAX_ROWS=820 DRR_TILE_MAX_ROWS=400 DRR_TILE_MAX_ROWS=270 "DRR_TILE_MAX_ROWS=400 DRR_TILE_LPG=32" "DRR_TILE_MAX_ROWS=400 DRR_TILE_LPG=8"
done
touch doom_rust_renderer_b200/csrc/drr_tile.cu; make -s -C doom_rust_renderer_b200/csrc > /dev/null

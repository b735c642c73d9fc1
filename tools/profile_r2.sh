#!/bin/bash
# ncu captures of the draw path for profiles/ (run under gpurun, one GPU): one --set full capture of the tile kernel at
# 1280x800 and at 320x200, each preceded by the same command without ncu.  usage: tools/profile_r2.sh <tag>
TAG=${1:-r2}
set -x
CMD="python bench.py --workload walk1280 --views 128 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/${TAG}_plain_1280.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_tile -s 4 -c 1 -o gpurun_out/${TAG}_prof_1280 -f $CMD > gpurun_out/${TAG}_ncu_1280.log 2>&1
CMD="python bench.py --workload walk320 --views 1024 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/${TAG}_plain_320.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_tile -s 4 -c 1 -o gpurun_out/${TAG}_prof_320 -f $CMD > gpurun_out/${TAG}_ncu_320.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_320.log

#!/bin/bash
# ncu capture of the device front-end kernel (run under gpurun, one GPU).  usage: tools/profile_fe_r2.sh <tag> [workload] [views]
TAG=${1:-r2_fe}; WL=${2:-walk320}; VIEWS=${3:-4096}
set -x
CMD="python bench.py --workload $WL --views $VIEWS --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_frontend -s 3 -c 1 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

#!/bin/bash
# ncu capture of the device front-end (drr_frontend_kernel: count pass + emit pass) for profiles/ -- run under gpurun, one GPU.
# The same command runs once without ncu first.
set -x
CMD="python bench.py --workload walk320 --views 2048 --steps 1 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r1_plain_fe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_frontend -c 2 -o gpurun_out/r1_prof_fe $CMD > gpurun_out/r1_ncu_fe.log 2>&1
tail -2 gpurun_out/r1_ncu_fe.log

#!/bin/bash
# ncu captures of the draw path for profiles/ (run under gpurun, one GPU): a launch list of a short bench run and one
# --set full capture of the bin + tile kernels at 320x200 and at 1280x800.  Each ncu run is preceded by the same command
# without ncu.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary=walk1280"
$CMD > gpurun_out/r1_plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu_launches.log 2>&1
CMD="python bench.py --workload walk320 --views 1024 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r1_plain_320.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_ -s 6 -c 2 -o gpurun_out/r1_prof_320 $CMD > gpurun_out/r1_ncu_320.log 2>&1
CMD="python bench.py --workload walk1280 --views 128 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r1_plain_1280.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_ -s 6 -c 2 -o gpurun_out/r1_prof_1280 $CMD > gpurun_out/r1_ncu_1280.log 2>&1
tail -2 gpurun_out/r1_ncu_1280.log

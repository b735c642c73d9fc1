#!/bin/bash
# ncu captures of the device front-end for profiles/ -- run under gpurun, one GPU.  Each ncu run is preceded by the same
# command without ncu.  (1) launch list of a short bench run (front-end, compaction, bin, tile); (2) --set full capture of
# drr_frontend_kernel + drr_fe_compact_kernel on a 2048-viewpoint batch at 320x200.
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r1_plain_fe_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_fe_launches.csv $CMD > gpurun_out/r1_ncu_fe_launches.log 2>&1
CMD="python bench.py --workload walk320 --views 2048 --steps 1 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r1_plain_fe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:drr_f -c 2 -o gpurun_out/r1_prof_fe $CMD > gpurun_out/r1_ncu_fe.log 2>&1
tail -2 gpurun_out/r1_ncu_fe.log

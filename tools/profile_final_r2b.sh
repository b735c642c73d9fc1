#!/bin/bash
# Round-2 evidence for profiles/ (run under gpurun, one GPU; every ncu run preceded by the same command without ncu):
#   1. launch list of a short bench run (every kernel with its device time)
#   2. --set full captures of the tile kernel at 1280x800 and 320x200, and of the bin kernel at 320x200
#   3. DRAM traffic of one tile-kernel launch over the FULL batch of each workload (dram__bytes_read/write.sum)
set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary=walk1280"
$CMD > gpurun_out/r2b_final_plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_final_launches.csv $CMD > gpurun_out/r2b_final_ncu_launches.log 2>&1
./tools/profile_r2.sh r2b_final
CMD="python bench.py --workload walk320 --views 1024 --steps 2 --warmup 3 --no-cpu-baseline --secondary="
ncu --set full --clock-control none --import-source on -k regex:drr_bin -s 4 -c 1 -o gpurun_out/r2b_final_prof_bin_320 -f $CMD > gpurun_out/r2b_final_ncu_bin.log 2>&1
for wl in walk320 walk1280 walls1280 flats1280 things640; do
  CMD="python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --secondary="
  $CMD > gpurun_out/r2b_final_plain_traffic_$wl.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:drr_tile -s 4 -c 1 --csv --log-file gpurun_out/r2b_final_traffic_$wl.csv $CMD > gpurun_out/r2b_final_ncu_traffic_$wl.log 2>&1
done
CMD="python bench.py --workload stress1920 --views 1024 --steps 1 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/r2b_final_plain_traffic_stress1920.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none -k regex:drr_tile -s 4 -c 1 --csv --log-file gpurun_out/r2b_final_traffic_stress1920.csv $CMD > gpurun_out/r2b_final_ncu_traffic_stress.log 2>&1
tail -3 gpurun_out/r2b_final_traffic_walk320.csv
./tools/profile_fe_warm.sh r2b_final_fe walk320 4096

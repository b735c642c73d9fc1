#!/bin/bash
# ncu capture of the device front-end kernel with WARM caches (--cache-control none: the stateless kernel's records and the map are in L2
# as in a real run; the default flush before every replay pass makes every global load look like a DRAM miss).
# usage: tools/profile_fe_warm.sh <tag> [workload] [views]
TAG=${1:-r2_fe}; WL=${2:-walk320}; VIEWS=${3:-4096}
CMD="python bench.py --workload $WL --views $VIEWS --steps 2 --warmup 3 --no-cpu-baseline --secondary="
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && ncu --set full --clock-control none --cache-control none --import-source on -k regex:drr_frontend -s 3 -c 1 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log

#!/bin/bash
# Final ncu evidence on the DEFAULT bench workload (config2, full size).
TAG=${1:-r1final}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:.*sym::.*' -c 120 \
    --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:.*knn_scan_tc.*" -s 4 -c 1 -f -o gpurun_out/prof_$TAG \
    $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300

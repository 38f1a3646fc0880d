#!/bin/bash
# ncu evidence for one round: launch list of the whole step + one full capture of the top kernel.
# usage: tools/gpu_profile.sh <tag> <kernel-regex> [extra bench args]
TAG=${1:-r1}; KREGEX=${2:-knn_scan}; shift 2
mkdir -p gpurun_out
CMD="python bench.py --workload mid --steps 2 --warmup 1 --no-e2e --no-cpu $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:.*sym::.*' -c 200 \
    --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:.*$KREGEX.*" -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full_$TAG.log
cat gpurun_out/plain_$TAG.log | tail -1 | cut -c1-1500

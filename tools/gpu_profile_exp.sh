#!/bin/bash
# ncu full capture of the scan kernel for an experiment build.  usage: gpu_profile_exp.sh <tag> "<nvcc defines>"
TAG=$1; DEF=$2
SYM_NVCC_EXTRA="$DEF" python -m symphonypy_b200.build --force > /dev/null 2>&1
export SYM_KNN_NO_REPAIR=1
CMD="python bench.py --workload mid --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:.*knn_scan_tc.*" -s 1 -c 1 -f -o gpurun_out/prof_$TAG \
    $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"

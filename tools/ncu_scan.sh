#!/bin/bash
# one --set full capture (with source-level stall sampling) of the first-tier scan kernel on an isotropic
# 200k x 500k, d=30, k=10 problem: tools/ncu_scan.sh <tag> [kernel-regex]
TAG=$1; K=${2:-knn_scan_tc}
mkdir -p gpurun_out
N=${N:-200000} timeout 300 python tools/tc_profile.py > gpurun_out/plain_$TAG.log 2>&1 &&
N=${N:-200000} timeout 600 ncu --set full --clock-control none --import-source on -k "regex:.*$K.*" -c 1 -f -o gpurun_out/prof_$TAG \
    python tools/tc_profile.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "capture exit $?"; tail -2 gpurun_out/plain_$TAG.log

#!/bin/bash
# one --set full capture each of the memory-bound kernels of the default step (projection, assign, MoE statistics,
# MoE apply, re-rank): tools/ncu_others.sh <tag>
TAG=$1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --configs none"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on \
    -k "regex:project_tma_kernel|project_tc_kernel|assign_kernel|assign_tc_kernel|moe_stats_kernel|moe_apply_kernel|knn_rerank" \
    --launch-skip ${SKIP:-0} -c ${COUNT:-5} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "capture exit $?"; tail -2 gpurun_out/ncu_full_$TAG.log

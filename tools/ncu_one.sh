#!/bin/bash
# one full ncu capture of a named kernel of the default bench step: tools/ncu_one.sh <tag> <kernel-regex> [skip]
TAG=$1; K=$2; S=${3:-2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:.*$K.*" -s $S -c 1 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "capture exit $?"

#!/bin/bash
# Freeze a copy of the working tree under gpurun_stage/<tag>/ so that a queued gpurun call runs exactly what was
# staged while the main tree keeps changing.  usage: tools/stage.sh <tag>; then
#   gpurun -- 'cd gpurun_stage/<tag> && <command writing to $OUT = ../../gpurun_out>'
TAG=$1
rm -rf gpurun_stage/$TAG && mkdir -p gpurun_stage/$TAG
tar -cf - --exclude=./.git --exclude=./gpurun_out --exclude=./gpurun_stage --exclude=__pycache__ --exclude='*.o' \
    --exclude=.pytest_cache --exclude='*.ncu-rep' --exclude=./tools/microbench . | tar -xf - -C gpurun_stage/$TAG
du -sh gpurun_stage/$TAG | cut -f1

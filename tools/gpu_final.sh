#!/bin/bash
# End-of-round evidence: full GPU test suite, smoke, all bench configs, reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
bash tools/gpu_allconfigs.sh 2>&1 | cut -c1-600
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --extras 2>gpurun_out/bx.err | tail -1 > gpurun_out/bench_extras.json
bash tools/gpu_profile_final.sh ${1:-r1i} 2>&1 | tail -4

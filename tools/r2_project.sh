#!/bin/bash
# projection kernels: parity tests, then stage times with the TMA-staged and the register-prefetch kernel
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "project or golden or config1 or container or streaming" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "projection tests failed: stopping"; exit 1; }
for T in 1 0; do for W in ${WORKLOADS:-config2 config3}; do
  SYM_PROJECT_TMA=$T timeout 200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e --configs none 2>gpurun_out/r2proj.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=[k for k in d['roofline']['other_kernels'] if k['stage']=='project'][0]; print('  $W TMA=$T: project ms', round(o['ms'],3), 'GB/s', round(o['achieved']), 'frac', round(o['frac'],3), '| stages', {k:round(v,2) for k,v in d['stage_ms'].items()})"
done; done

#!/bin/bash
# stage times of the three large configs (device-resident value only)
for W in ${WORKLOADS:-config2 config3 config4}; do
  timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/st_$W.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$W value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stage_ms'].items()})"
done

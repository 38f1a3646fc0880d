#!/bin/bash
# kNN stage with BF16x3 vs single-FP16 operands, repairs disabled: pure scan speed.
export SYM_KNN_NO_REPAIR=1
for W in config2 config3; do
for F in 0 1; do
  echo "== $W SYM_TC_FP16=$F"
  SYM_TC_FP16=$F timeout 400 python bench.py --workload $W --steps 3 --warmup 2 --no-cpu --no-e2e 2>gpurun_out/fp16_${W}_${F}.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn ms', d['stage_ms']['knn'], 'roofline', {k: d['roofline'][k] for k in d['roofline'] if k in ('kernel_ms','kernel','achieved','frac','rows_repaired_per_step')})"
done; done

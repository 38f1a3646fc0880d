#!/bin/bash
# Quick check of a scan-kernel change: kNN parity tests, then the kNN stage time at config2/config3.
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "knn or golden or tc or full_size" 2>&1 | tail -12
for W in ${WORKLOADS:-config2 config3}; do
  for F in ${FP16_MODES:-1 0}; do
  echo "== $W SYM_KNN_FP16=$F"
  SYM_KNN_FP16=$F timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/sq_${W}_${F}.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn ms', round(d['stage_ms']['knn'],2), 'repaired/step', d['roofline'].get('rows_repaired_per_step'), 'value', round(d['value']), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'][:40])"
  done
done

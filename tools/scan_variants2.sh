#!/bin/bash
# kNN stage (config2 / config3, repairs on) for tuning variants of the selection path (single-FP16 first tier).
run() {
  echo "== $*"
  for W in config2 config3; do
  env "$@" timeout 300 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $W knn ms', round(d['stage_ms']['knn'],2), 'repaired', d['roofline'].get('rows_repaired_per_step'))"
  done
}
run A=base
run SYM_TC_PAD=6
SYM_NVCC_EXTRA="-DTC_FLUSH=16" python -m symphonypy_b200.build --force > /dev/null 2>&1
run FLUSH=16
SYM_NVCC_EXTRA="-DTC_FLUSH=1048576" python -m symphonypy_b200.build --force > /dev/null 2>&1
run FLUSH=never
run FLUSH=never SYM_TC_PAD=6

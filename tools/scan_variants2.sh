#!/bin/bash
# kNN stage (config2, repairs on) for a few tuning variants of the selection path.
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --workload ${W:-config2} --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn ms', round(d['stage_ms']['knn'],2), 'repaired', d['roofline'].get('rows_repaired_per_step'))"
}
run SYM_TC_SS=0
run SYM_TC_SS=1
run SYM_TC_SS=1 SYM_TC_PAD=6
run SYM_TC_SS=1 SYM_TC_PAD=4
SYM_NVCC_EXTRA="-DTC_FLUSH=8" python -m symphonypy_b200.build --force > /dev/null 2>&1
run SYM_TC_SS=1 FLUSH=8
SYM_NVCC_EXTRA="-DTC_FLUSH=2" python -m symphonypy_b200.build --force > /dev/null 2>&1
run SYM_TC_SS=1 FLUSH=2
SYM_NVCC_EXTRA="-DTC_FLUSH=16" python -m symphonypy_b200.build --force > /dev/null 2>&1
run SYM_TC_SS=1 FLUSH=16

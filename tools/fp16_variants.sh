#!/bin/bash
# Scan-only timing (repairs off) of structural variants, with cheap single-FP16 MMAs so that selection dominates.
export SYM_KNN_NO_REPAIR=1
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --workload config2 --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('knn ms', round(d['stage_ms']['knn'],2))"
}
run SYM_TC_FP16=1
run SYM_TC_FP16=1 SYM_TC_SS=1
run SYM_TC_FP16=1 SYM_TC_QT=1
run SYM_TC_FP16=1 SYM_TC_HALVES=2
run SYM_TC_FP16=0 SYM_TC_SS=1
run SYM_TC_FP16=0 SYM_TC_HALVES=2

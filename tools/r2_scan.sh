#!/bin/bash
# Round 2 check of a scan-kernel change: kNN parity tests, kNN stage time, cycle accounting.
OUT=${OUT:-gpurun_out}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "knn or golden or tc or full_size or config1" 2>&1 | tail -5
run() {
  W=$1; shift
  env "$@" timeout 400 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e --configs none 2>$OUT/r2scan_$W.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $W $*: knn ms', round(d['stage_ms']['knn'],2), 'scan+rerank', round(d['roofline']['ms_per_launch'],2), 'repaired', d['roofline'].get('rows_repaired_per_step'), 'frac', round(d['roofline']['frac'],4), 'step ms', round(d['ms_per_step'],2))"
}
for S in ${PERIODS:-16}; do run config2 SYM_TC_PERIOD=$S; done
for W in ${WORKLOADS:-config3}; do run $W SYM_TC_PERIOD=16; done
if [ -f symphonypy_b200/_C/libsymphony_b200_prof.so ]; then
  for S in ${PROF_PERIODS:-16}; do
    echo "== cycle accounting, period $S"
    SYM_TC_PERIOD=$S SYM_B200_LIB=$PWD/symphonypy_b200/_C/libsymphony_b200_prof.so timeout 300 python tools/tc_profile.py 2>&1 | tail -5
  done
fi

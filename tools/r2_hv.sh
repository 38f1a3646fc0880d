#!/bin/bash
# Check of a scan-kernel change with SHORT timeouts (a hung kernel must not eat the GPU budget): a quick kNN parity
# subset first, everything else only if that passed.
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "knn_tcgen05_matches_sklearn or knn_fp16_tier" 2>&1 | tail -3
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "quick kNN tests failed or hung: stopping"; exit 1; }
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "knn or golden or tc or full_size or config1" 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "kNN tests failed: stopping"; exit 1; }
run() {
  W=$1; shift
  env "$@" timeout 150 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e --configs none 2>gpurun_out/r2scan_$W.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  $W $*: knn ms', round(d['stage_ms']['knn'],2), 'scan+rerank', round(d['roofline']['ms_per_launch'],2), 'repaired', d['roofline'].get('rows_repaired_per_step'), 'frac', round(d['roofline']['frac'],4), 'step ms', round(d['ms_per_step'],2))"
}
IFS=';' read -ra RUNS <<< "${RUNS:-config2 SYM_TC_HALVES=2;config2 SYM_TC_HALVES=1;config3 SYM_TC_HALVES=2}"
for R in "${RUNS[@]}"; do run $R; done
if [ -f symphonypy_b200/_C/libsymphony_b200_prof.so ]; then
  for H in ${PROF_HALVES:-2 1}; do echo "== cycle accounting halves $H"; SYM_TC_HALVES=$H SYM_B200_LIB=$PWD/symphonypy_b200/_C/libsymphony_b200_prof.so timeout 120 python tools/tc_profile.py 2>&1 | tail -4; done
fi

#!/bin/bash
for DEF in "-DTC_FLUSH=1" "-DTC_FLUSH=4" "-DTC_FLUSH=16"; do
  SYM_NVCC_EXTRA="$DEF" python -m symphonypy_b200.build --force > /dev/null 2>&1
  echo "== $DEF"
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('knn ms', d['stage_ms']['knn'], d['roofline']['rows_repaired_per_step'])"
done
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "knn or tc" 2>&1 | tail -2

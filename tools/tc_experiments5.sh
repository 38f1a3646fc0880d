#!/bin/bash
for DEF in "-DTC_SLEEPY_SELECT"; do
  SYM_NVCC_EXTRA="$DEF" python -m symphonypy_b200.build --force > /dev/null 2>&1
  for H in 1 2; do
  echo "== $DEF halves=$H"
  SYM_TC_HALVES=$H timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('knn ms', d['stage_ms']['knn'])"
  done
done

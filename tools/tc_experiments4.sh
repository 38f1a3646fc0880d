#!/bin/bash
export SYM_KNN_NO_REPAIR=1
for DEF in "-DTC_EXPERIMENT_NO_WALK" "-DTC_FIFO_N=1" "-DTC_FIFO_N=2" "-DTC_FIFO_N=4"; do
  SYM_NVCC_EXTRA="$DEF" python -m symphonypy_b200.build --force > /dev/null 2>&1
  echo "== $DEF"
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('knn ms', d['stage_ms']['knn'])"
done

#!/bin/bash
SYM_NVCC_EXTRA="-DTC_EXPERIMENT_NO_SELECT" python -m symphonypy_b200.build --force > /dev/null 2>&1
for CFG in "2 1" "2 2" "2 3" "1 1" "1 2" "1 3" "1 4"; do
  set -- $CFG
  echo "== NO_SELECT qt=$1 nbuf=$2"
  SYM_KNN_NO_REPAIR=1 SYM_TC_QT=$1 SYM_TC_NBUF=$2 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --workload mid > /tmp/o.txt 2>/tmp/e.txt
  tail -1 /tmp/o.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('knn ms', d['stage_ms']['knn'])" 2>/dev/null || tail -3 /tmp/e.txt
done

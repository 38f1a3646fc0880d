#!/bin/bash
mkdir -p gpurun_out
for W in config1 config3 config4; do
  echo "== $W"
  timeout 600 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err
  echo "exit $?"; tail -2 gpurun_out/bench_$W.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$W.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',d['e2e'] and round(d['e2e']['value']),'stages',{k:round(v,2) for k,v in d['stage_ms'].items()},'repaired',d['roofline'].get('rows_repaired_per_step'))"
done
echo "== default full"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; tail -2 gpurun_out/bench_default.err
cat gpurun_out/bench_default.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"
cat gpurun_out/bench_ref.json | cut -c1-700

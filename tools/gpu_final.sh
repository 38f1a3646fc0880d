#!/bin/bash
# end-of-round evidence: GPU tests, smoke, default bench, reference arm, ncu launch list + full capture of the scan
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_tests.log
timeout 150 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 700 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "reference arm rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --configs none"
# (the coarse clustering of knn_fit is seeded with k-means++: ~1000 tiny launches that would fill the list; this pass
# starts Lloyd from the first sample rows instead — same kernels in the step itself)
SYM_KNN_COARSE_INIT=first timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:.*sym::.*' -c 200 \
    --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:.*knn_scan_tc.*" -c 1 -f -o gpurun_out/prof_$TAG \
    $CMD > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full capture exit $?"

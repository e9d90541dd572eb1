#!/bin/bash
# round 2, GPU call 33: device time of the screening launch and of halves_merge_kernel, both merge modes (ncu launch
# list, free clocks), and the rest of the deferred-merge parity test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 1 0; do
VVB_TC_DEFER_MERGE=$m timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_screen|halves_merge|knn_rerank" --csv --log-file gpurun_out/r33_launches_$m.csv python profiles/run_screen.py 1000000 > gpurun_out/r33_ncu_$m.log 2>&1
echo "== VVB_TC_DEFER_MERGE=$m rc=$?"
grep -v "^==" gpurun_out/r33_launches_$m.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:60], r['Metric Value'], r['Metric Unit'])"
done
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "deferred" > gpurun_out/r33_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r33_tests.log)"

#!/bin/bash
# round 2, GPU call 36: merge kernel with speculative loads (device time from the ncu launch list), the kNN parity
# suites with the deferred merge as the default
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_screen|halves_merge" --csv --log-file gpurun_out/r36_launches.csv python profiles/run_screen.py 1000000 > gpurun_out/r36_ncu.log 2>&1
echo "ncu rc=$?"; grep "halves\|knn_screen" gpurun_out/r36_launches.csv | awk -F, '{print $5, $NF}' | cut -c1-120
timeout -s KILL 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r36_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r36_tests.log)"

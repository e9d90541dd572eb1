#!/bin/bash
# round 2, GPU call 37: register-tiled pivot-assignment / extent kernels -- parity, launch durations both ways
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "tiled or parts" > gpurun_out/r37_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r37_tests.log)"; grep "tiled plan" gpurun_out/r37_tests.log
for m in 1 0; do
VVB_CELLS_TILED=$m timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cell_assign|cell_extent" --csv --log-file gpurun_out/r37_launches_$m.csv python profiles/run_screen.py 1000000 > gpurun_out/r37_ncu_$m.log 2>&1
echo "== VVB_CELLS_TILED=$m rc=$?"; grep "cell_" gpurun_out/r37_launches_$m.csv | awk -F, '{print $5, $NF}' | cut -c1-100; tail -2 gpurun_out/r37_ncu_$m.log | head -1 | cut -c1-330
done

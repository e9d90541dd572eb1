#!/bin/bash
# round 2, GPU call 44: operand build with two rows per warp in flight -- parity suites and launch durations
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r44_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r44_tests.log)"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"build_operands" --csv --log-file gpurun_out/r44_launches.csv python profiles/run_screen.py 1000000 > gpurun_out/r44_ncu.log 2>&1
echo "ncu rc=$?"; grep "build_" gpurun_out/r44_launches.csv | awk -F, '{print $5, $NF}' | cut -c1-100
timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | tail -2 | head -1 | cut -c1-400

#!/bin/bash
# round 2, GPU call 19+: list appends through a register-resident list pointer; q15 threshold selection
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r19_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r19_tests.log)"
timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r19_screen.txt 2>&1
echo "rc=$? $(tail -2 gpurun_out/r19_screen.txt | head -1 | cut -c90-420)"
VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times" | tail -1
timeout -s KILL 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r19_tests_cfg.log 2>&1
echo "config tests rc=$? $(tail -1 gpurun_out/r19_tests_cfg.log)"

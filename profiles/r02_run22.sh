#!/bin/bash
# round 2, GPU call 22: two rows per thread in the preparation kernels (parity + timing)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r22_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r22_tests.log)"
timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r22_screen.txt 2>&1
echo "rc=$? $(tail -2 gpurun_out/r22_screen.txt | head -1 | cut -c1-200)"
timeout -s KILL 300 python profiles/run_screen.py 4200000 > gpurun_out/r22_screen4M.txt 2>&1
echo "rc=$? $(tail -2 gpurun_out/r22_screen4M.txt | head -1 | cut -c1-200)"

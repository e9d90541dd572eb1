#!/bin/bash
# round 2, GPU call 21: FFMA2 in the preparation kernels (parity + timing), smoother protocol probes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r21_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r21_tests.log)"
timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r21_screen.txt 2>&1
echo "rc=$? $(tail -2 gpurun_out/r21_screen.txt | head -1 | cut -c1-200)"
for j in 0 2 3 4; do
VVB_SMOOTH_JITTER=$j timeout -s KILL 300 python profiles/r02_smooth_probe.py 2>&1 | grep jitter
done

#!/bin/bash
# round 2, GPU call 16: one-product threshold pre-pass of the resident kernel (VVB_TC_PRE_HI) -- parity + timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r16_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r16_tests.log)"
for h in 0 1; do
VVB_TC_PRE_HI=$h timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r16_screen_prehi$h.txt 2>&1
echo "pre_hi=$h rc=$? $(tail -2 gpurun_out/r16_screen_prehi$h.txt | head -1 | cut -c1-330)"
VVB_TC_PRE_HI=$h VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times" | tail -1
done

#!/bin/bash
# round 2, GPU call 18: two-row selection (default build) and a 6-stage database ring (libvivae_b200_s6.so)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r18_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r18_tests.log)"
for lib in libvivae_b200.so libvivae_b200_s6.so; do
export VVB_LIB_PATH=$PWD/vivae_b200/$lib
timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r18_screen_$lib.txt 2>&1
echo "$lib rc=$? $(tail -2 gpurun_out/r18_screen_$lib.txt | head -1 | cut -c90-420)"
VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times" | tail -1
done

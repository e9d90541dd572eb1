#!/bin/bash
# round 2, GPU call 15: wide-row first tier with database slices (VVB_TC_WSLICES), 4 waves against a 500k x 2000 database
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "c5_slice or screening_error" > gpurun_out/r15_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r15_tests.log)"
for s in 1 2 4 8; do
VVB_TC_WSLICES=$s NQ_BLOCKS=592 REPS=3 timeout -s KILL 300 python profiles/r02_prof_wide.py > gpurun_out/r15_wide_s$s.txt 2>&1
echo "slices=$s rc=$? $(tail -1 gpurun_out/r15_wide_s$s.txt | cut -c1-400)"
done

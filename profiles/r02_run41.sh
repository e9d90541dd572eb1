#!/bin/bash
# round 2, GPU call 41: TF32 tensor-core pivot assignment -- parity (three plan variants, plan in parts, C1/tie sets),
# launch durations of the preparation kernels, the 1M x 50 step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "tiled or parts or c1_ or tie_heavy or any_partition" > gpurun_out/r41_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r41_tests.log)"; grep "tiled plan" gpurun_out/r41_tests.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cell_|build_operands|colsum|maxnorm" --csv --log-file gpurun_out/r41_launches.csv python profiles/run_screen.py 1000000 > gpurun_out/r41_ncu.log 2>&1
echo "ncu rc=$?"; grep "vvb::\|cell_\|build_\|colsum\|maxnorm" gpurun_out/r41_launches.csv | tail -22 | awk -F, '{print $5, $NF}' | cut -c1-100
tail -2 gpurun_out/r41_ncu.log | head -1 | cut -c1-400
timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | tail -2 | head -1 | cut -c1-400

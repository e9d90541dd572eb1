#!/bin/bash
# round 2, GPU call 38: host-path split of the end-to-end step (make_knn / smooth), tiled-plan parity test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python profiles/e2e_split_probe.py > gpurun_out/r38_e2e_split.json 2> gpurun_out/r38_e2e_split.err
echo "probe rc=$?"; cat gpurun_out/r38_e2e_split.json
nproc; free -g | head -2
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s -k "tiled" > gpurun_out/r38_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r38_tests.log)"; grep "tiled plan" gpurun_out/r38_tests.log

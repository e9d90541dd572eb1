#!/bin/bash
# round 2, GPU call 29: ring build with several shards per sweep (one-process emulation + every kernel at small sizes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ring" > gpurun_out/r29_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r29_tests.log)"
timeout -s KILL 120 python profiles/sanitize_small.py > gpurun_out/r29_small.log 2>&1
echo "small rc=$? $(tail -1 gpurun_out/r29_small.log)"

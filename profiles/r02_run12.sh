#!/bin/bash
# round 2, GPU call 12: fused training step with staged weights
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_fused_step.py -m gpu -q --tb=short > gpurun_out/r12_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r12_tests.log | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fit or graph or pipeline or smoke or geometric" > gpurun_out/r12_tests2.log 2>&1
echo "tests2 rc=$?"; tail -8 gpurun_out/r12_tests2.log | cut -c1-300
timeout -s KILL 600 python profiles/r02_fit_probe.py 2>&1 | tail -7

#!/bin/bash
# round 2, GPU call 1: new parity tests, the old GPU suite, a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_configs.py -m gpu -q --tb=short -s > gpurun_out/r1_configs.log 2>&1
echo "configs rc=$?" >> gpurun_out/r1_configs.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short > gpurun_out/r1_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r1_parity.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
echo "bench rc=$?" >> gpurun_out/r1_bench.err
tail -5 gpurun_out/r1_configs.log; tail -5 gpurun_out/r1_parity.log; tail -3 gpurun_out/r1_bench.err

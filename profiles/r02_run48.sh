#!/bin/bash
# round 2, GPU call 48: fused training step with cp.async double-buffered weight staging -- parity (fused step vs autograd,
# fit replay, graph step vs eager) and the step's time; the same kernel (before this change) with 1024 threads per CTA
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_parity.py -x -q -m gpu -k "fused or fit or step or graph or train" > gpurun_out/r48_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r48_tests.log)"
echo "== default library (cp.async staging, 512 threads)"
timeout -s KILL 300 python profiles/r02_fit_probe.py 2>&1 | grep -v Warning | tail -7
if [ -f vivae_b200/libvivae_b200_t1024.so ]; then
echo "== synchronous staging, 1024 threads"
VVB_LIB_PATH=$PWD/vivae_b200/libvivae_b200_t1024.so timeout -s KILL 300 python profiles/r02_fit_probe.py 2>&1 | grep -v Warning | tail -7
fi

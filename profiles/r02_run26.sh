#!/bin/bash
# round 2, GPU call 26: fused training step with 16 vs 32 warps per block
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in libvivae_b200.so libvivae_b200_t1024.so; do
export VVB_LIB_PATH=$PWD/vivae_b200/$lib
echo "== $lib"
timeout -s KILL 300 python profiles/r02_fit_probe.py 2>&1 | grep -v Warning | tail -5
done

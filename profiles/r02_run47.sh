#!/bin/bash
# round 2, GPU call 47 (2 GPUs): pieces of the gathered one-sweep ring build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 profiles/r02c_ring_probe.py > gpurun_out/r47_ring_probe.json 2> gpurun_out/r47_ring_probe.err
echo "rc=$?"; tail -1 gpurun_out/r47_ring_probe.json; grep -v "^\*\|OMP_NUM" gpurun_out/r47_ring_probe.err | tail -5

#!/bin/bash
# round 2, GPU call 14 (2 GPUs): the smoother over NVLink peer memory
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 profiles/r02_check_2gpu.py > gpurun_out/r14_check_2gpu.json 2> gpurun_out/r14_check_2gpu.err
echo "check rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r14_check_2gpu.err | tail -12; cat gpurun_out/r14_check_2gpu.json | tail -1 | cut -c1-1500
nvidia-smi --query-gpu=index,memory.used --format=csv,noheader

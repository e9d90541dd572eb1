#!/bin/bash
# round 2, GPU call 4 (2 GPUs): sharded preparation over NCCL, bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 profiles/r02_check_2gpu.py > gpurun_out/r4_check_2gpu.json 2> gpurun_out/r4_check_2gpu.err
echo "check rc=$?"; tail -3 gpurun_out/r4_check_2gpu.err; cat gpurun_out/r4_check_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r4_bench_2gpu.json 2> gpurun_out/r4_bench_2gpu.err
echo "bench rc=$?"; tail -3 gpurun_out/r4_bench_2gpu.err
python - <<'PY'
import json
try:
    j=json.load(open("gpurun_out/r4_bench_2gpu.json"))
    print("ms/step", round(j["ms_per_step"],2), "knn", round(j["ms_knn"],2), "smooth", round(j["ms_smooth"],2), {k:round(v,2) for k,v in j["knn_stats"].items() if k.startswith("ms_")}, "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"],1), "gates", j["parity_gates"])
    print("c3", j["config_c3"]["ms_per_step"], j["config_c3"]["knn_stats_rank0"])
except Exception as e: print("ERR", e)
PY

#!/bin/bash
# round 2, GPU call 43 (N GPUs, N = $1): the multi-GPU paths after the round-2c kernel changes -- 2-GPU check of the
# sharded plan / ring / smoother layouts (N = 2 only) and the default bench at N, as the driver launches it
N=${1:-2}
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 profiles/r02_check_2gpu.py > gpurun_out/r43_check_2gpu.json 2> gpurun_out/r43_check_2gpu.err
echo "check rc=$?"; tail -1 gpurun_out/r43_check_2gpu.json | cut -c1-900
fi
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/r43_bench_${N}gpu.json 2> gpurun_out/r43_bench_${N}gpu.err
echo "bench rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r43_bench_${N}gpu.err | tail -5
python - <<PY
import json
try:
    j=json.loads([l for l in open("gpurun_out/r43_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("ms/step", round(j["ms_per_step"],2), "knn", round(j["ms_knn"],2), "smooth", round(j["ms_smooth"],2), {k:round(v,2) for k,v in j["knn_stats"].items() if k.startswith("ms_")}, "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"],1), "gates", j["parity_gates"]["knn_rows_exact"], j["parity_gates"]["smooth_max_abs_err"])
    for c in ("config_c3","config_c5","config_overlap"):
        cc=j.get(c) or {}; print(c, cc.get("ms_per_step"), cc.get("value"), {k:round(v,1) for k,v in cc.get("knn_stats_rank0",{}).items() if k.startswith("ms_") or k.startswith("rows_")}, cc.get("roofline",{}).get("frac"))
    print("ring", j.get("ring")); print(j["clocks"])
except Exception as e: print("ERR", e)
PY

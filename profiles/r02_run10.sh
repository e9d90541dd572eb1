#!/bin/bash
# round 2, GPU call 10 (2 GPUs): column-sharded smoother, stress tests, bench at N=2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_configs.py -m gpu -q --tb=short -k "column_slices or injected_delays" > gpurun_out/r10_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r10_tests.log | cut -c1-300
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 10 --warmup 3 --skip-c3 --skip-fit > gpurun_out/r10_bench_2gpu.json 2> gpurun_out/r10_bench_2gpu.err
echo "bench rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/r10_bench_2gpu.err | tail -5
python - <<'PY'
import json
try:
    j=json.loads([l for l in open("gpurun_out/r10_bench_2gpu.json") if l.startswith("{")][-1])
    print("ms/step", round(j["ms_per_step"],2), "knn", round(j["ms_knn"],2), "smooth", round(j["ms_smooth"],2), {k:round(v,2) for k,v in j["knn_stats"].items() if k.startswith("ms_")}, "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"],1), "gates", j["parity_gates"]["knn_rows_exact"], j["parity_gates"]["smooth_max_abs_err"])
except Exception as e: print("ERR", e)
PY

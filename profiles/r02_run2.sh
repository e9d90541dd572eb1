#!/bin/bash
# round 2, GPU call 2: failing tests again, bench (N=1), launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "screening_error or few_row or parts or refine or tie_heavy or c1_" -s > gpurun_out/r2_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_tests.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?" >> gpurun_out/r2_bench.err
VVB_TC_LPT=0 timeout 600 python bench.py --steps 5 --warmup 3 --skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 > gpurun_out/r2_bench_nolpt.json 2> gpurun_out/r2_bench_nolpt.err
VVB_SMOOTH_PREFETCH=0 timeout 600 python bench.py --steps 5 --warmup 3 --skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 > gpurun_out/r2_bench_nopf.json 2> gpurun_out/r2_bench_nopf.err
tail -4 gpurun_out/r2_tests.log; tail -3 gpurun_out/r2_bench.err; python - <<'PY'
import json
for f in ("r2_bench","r2_bench_nolpt","r2_bench_nopf"):
    try:
        j=json.load(open(f"gpurun_out/{f}.json"))
        print(f, "ms/step", round(j["ms_per_step"],2), "knn", round(j["ms_knn"],2), "smooth", round(j["ms_smooth"],2), {k:round(v,2) for k,v in j["knn_stats"].items() if k.startswith("ms_")}, "e2e", j["e2e"] and round(j["e2e"]["ms_per_step"],1))
    except Exception as e: print(f, "ERR", e)
PY

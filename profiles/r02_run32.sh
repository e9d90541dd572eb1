#!/bin/bash
# round 2, GPU call 32 (third session): the join of the half lists moved out of the screening kernel
# (halves_merge_kernel, VVB_TC_DEFER_MERGE) -- parity tests, per-phase clocks and the 1M x 50 step both ways;
# the overlapping-mixture workload of bench.py (config_overlap) for the first time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "deferred or one_product or tie_heavy or c1_" > gpurun_out/r32_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r32_tests.log)"
for m in 1 0; do
echo "== VVB_TC_DEFER_MERGE=$m"
VVB_TC_DEFER_MERGE=$m VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times" | tail -1 | cut -c1-330
VVB_TC_DEFER_MERGE=$m timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | tail -2 | head -1 | cut -c1-400
done
SKIP="--skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5"
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 $SKIP > gpurun_out/r32_bench.json 2> gpurun_out/r32_bench.err
echo "bench rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/r32_bench.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['knn_stats'], j['roofline']['frac']); o=j['config_overlap']; print('overlap', o['ms_per_step'], o.get('tiles_swept_frac_rank0'), o['knn_stats_rank0'], o['roofline'])"

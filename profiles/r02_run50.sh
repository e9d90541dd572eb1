#!/bin/bash
# round 2, GPU call 50: final state of the session -- whole GPU suite, smoke(), default bench line, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r50_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r50_tests.log)"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r50_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r50_smoke.log | cut -c1-160)"
timeout -s KILL 900 python bench.py > gpurun_out/r50_bench_1gpu.json 2> gpurun_out/r50_bench_1gpu.err
echo "bench rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/r50_bench_1gpu.json').read().strip().splitlines()[-1])
print('step', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['ms_per_step'], 'sharded', j['e2e_sharded_api']['ms_per_step'])
print('knn', j['knn_stats']); print('roof', j['roofline']['frac'], j['roofline']['achieved'], j['roofline']['traffic'], 'smooth', j['roofline_smooth']['frac'], j['ms_smooth'], 'quartet', j['quartet']['frac'], j['quartet']['ms_fwd_bwd'])
print('c3', j['config_c3']['ms_per_step'], 'overlap', j['config_overlap']['ms_per_step'], 'fit', j['fit']['s_per_epoch'], j['fit']['speedup_vs_reference_cpu'], 'gates', j['parity_gates']['knn_rows_exact'], j['parity_gates']['smooth_max_abs_err'], 'clocks', j['clocks'], 'launches', j['gpu_launches'])"
SKIP="--skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 --skip-overlap"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r50_launches_bench_1M.csv python bench.py --steps 2 --warmup 1 $SKIP > gpurun_out/r50_ncu_list.log 2>&1
echo "ncu list rc=$?"

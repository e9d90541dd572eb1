#!/bin/bash
# round 2, GPU call 42 (third session): the whole GPU suite, smoke(), the default bench line and the reference arm,
# the launch list of the bench command and ncu --set full captures of the hot kernels -- after the deferred merge and
# the tensor-core / tiled cell-plan kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r42_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r42_tests.log)"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r42_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r42_smoke.log | cut -c1-200)"
timeout -s KILL 900 python bench.py > gpurun_out/r42_bench_1gpu.json 2> gpurun_out/r42_bench_1gpu.err
echo "bench rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/r42_bench_1gpu.json').read().strip().splitlines()[-1])
print('step', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['ms_per_step'], 'sharded', j['e2e_sharded_api']['ms_per_step'])
print('knn', j['knn_stats']); print('roof', j['roofline']['frac'], j['roofline']['achieved'], 'smooth', j['roofline_smooth']['frac'], j['ms_smooth'], 'quartet', j['quartet']['frac'])
print('c3', j['config_c3']['ms_per_step'], 'overlap', j['config_overlap']['ms_per_step'], 'fit', j['fit']['s_per_epoch'], 'gates', j['parity_gates'], 'clocks', j['clocks'])"
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r42_bench_reference_arm.json 2> gpurun_out/r42_ref.err
echo "ref rc=$? $(cut -c1-300 gpurun_out/r42_bench_reference_arm.json)"
SKIP="--skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 --skip-overlap"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r42_launches_bench_1M.csv python bench.py --steps 2 --warmup 1 $SKIP > gpurun_out/r42_ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --import-source on --clock-control none -k regex:"knn_screen|halves_merge|knn_rerank|smooth_sweep|cell_assign_mma|cell_extent_tiled" -s 6 -c 6 -o gpurun_out/r42_c2 python bench.py --steps 1 --warmup 1 $SKIP > gpurun_out/r42_ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r42_c2.ncu-rep --page raw --csv > gpurun_out/r42_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/r42_c2.ncu-rep --page source --csv -k regex:knn_screen > gpurun_out/r42_screen_source.csv 2>/dev/null
rm -f gpurun_out/r42_c2.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r42_c2_raw.csv')))
hdr=rows[0]
def g(r,k):
    return r[hdr.index(k)] if k in hdr else '-'
for r in rows[2:]:
    print(g(r,'Kernel Name')[:44], 'us', g(r,'gpu__time_duration.sum'), 'pipe', g(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'), 'dramR', g(r,'dram__bytes_read.sum'), 'dramW', g(r,'dram__bytes_write.sum'), 'issue', g(r,'sm__inst_issued.avg.pct_of_peak_sustained_active') )
PY

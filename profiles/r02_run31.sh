#!/bin/bash
# round 2, GPU call 31: the screening kernel with the three-product pre-pass (VVB_TC_PRE_HI=0) under ncu -- the
# configuration that maximises tensor-pipe activity, not speed
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SKIP="--skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5"
export VVB_TC_PRE_HI=0
timeout -s KILL 600 python bench.py --steps 3 --warmup 1 $SKIP > gpurun_out/r31_bench_prehi0.json 2> gpurun_out/r31_bench_prehi0.err
echo "bench rc=$?"; python -c "
import json; j=json.loads(open('gpurun_out/r31_bench_prehi0.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['knn_stats']['ms_screen'], j['roofline']['frac'], j['roofline']['achieved'])"
timeout -s KILL 900 ncu --set full --clock-control none -k regex:"knn_screen" -s 2 -c 1 -o gpurun_out/r31_prehi0 python bench.py --steps 1 --warmup 1 $SKIP > gpurun_out/r31_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r31_prehi0.ncu-rep --page raw --csv > gpurun_out/r31_prehi0_raw.csv 2>/dev/null
rm -f gpurun_out/r31_prehi0.ncu-rep
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r31_prehi0_raw.csv')))
hdr=rows[0]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:40], 'ms', r[hdr.index('gpu__time_duration.sum')], 'pipe', r[hdr.index('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')])
PY

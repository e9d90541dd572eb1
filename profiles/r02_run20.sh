#!/bin/bash
# round 2, GPU call 20: bench line at N=1, launch list of the bench command, ncu captures of the hot kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py > gpurun_out/r20_bench_1gpu.json 2> gpurun_out/r20_bench_1gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r20_bench_1gpu.json | cut -c1-600
SKIP="--skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5"
timeout -s KILL 600 python bench.py --steps 2 --warmup 1 $SKIP > gpurun_out/r20_bench_plain.log 2>&1 &&
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r20_launches.csv python bench.py --steps 2 --warmup 1 $SKIP > gpurun_out/r20_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"knn_screen|smooth_sweep|knn_rerank" -s 6 -c 3 -o gpurun_out/r20_c2 python bench.py --steps 1 --warmup 1 $SKIP > gpurun_out/r20_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
ncu -i gpurun_out/r20_c2.ncu-rep --page raw --csv > gpurun_out/r20_c2_raw.csv 2>/dev/null
ncu -i gpurun_out/r20_c2.ncu-rep --page source --csv -k regex:knn_screen > gpurun_out/r20_c2_screen_source.csv 2>/dev/null
rm -f gpurun_out/r20_c2.ncu-rep
NQ_BLOCKS=592 REPS=2 timeout -s KILL 300 python profiles/r02_prof_wide.py > gpurun_out/r20_prof_wide_plain.log 2>&1 &&
NQ_BLOCKS=592 REPS=2 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 1 -c 1 -o gpurun_out/r20_wide python profiles/r02_prof_wide.py > gpurun_out/r20_ncu_wide.log 2>&1
ncu -i gpurun_out/r20_wide.ncu-rep --page raw --csv > gpurun_out/r20_wide_raw.csv 2>/dev/null
rm -f gpurun_out/r20_wide.ncu-rep
echo "ncu wide rc=$?"; tail -1 gpurun_out/r20_prof_wide_plain.log | cut -c1-300
ls -la gpurun_out/ | tail -12

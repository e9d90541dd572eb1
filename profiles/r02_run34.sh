#!/bin/bash
# round 2, GPU call 34: ncu --set full of halves_merge_kernel (0.97 ms for 1M rows: what bounds it?)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:"halves_merge" -s 1 -c 1 -o gpurun_out/r34_merge python profiles/run_screen.py 1000000 > gpurun_out/r34_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r34_merge.ncu-rep --page raw --csv > gpurun_out/r34_merge_raw.csv 2>/dev/null
ncu -i gpurun_out/r34_merge.ncu-rep --page source --csv > gpurun_out/r34_merge_source.csv 2>/dev/null
ncu -i gpurun_out/r34_merge.ncu-rep --page details > gpurun_out/r34_merge_details.txt 2>/dev/null
rm -f gpurun_out/r34_merge.ncu-rep
ls -la gpurun_out/

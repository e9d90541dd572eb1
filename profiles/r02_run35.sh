#!/bin/bash
# round 2, GPU call 35: merge kernel compiled for 4 / 6 / 8 CTAs per SM (64 / 40 / 32 registers); where the main
# sweep's clocks go (VVB_TC_DEBUG=3: no TMEM drain, 1: drain only, 2: drain + minima, no admissions; results are
# garbage in these modes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for o in 4 6 8; do
VVB_TC_MERGE_OCC=$o timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"halves_merge" --csv --log-file gpurun_out/r35_launches_$o.csv python profiles/run_screen.py 1000000 > gpurun_out/r35_ncu_$o.log 2>&1
echo "== VVB_TC_MERGE_OCC=$o rc=$? $(grep halves gpurun_out/r35_launches_$o.csv | tail -1 | awk -F, '{print $NF}')"
done
for m in 0 3 1 2; do
echo "== VVB_TC_DEBUG=$m"
VVB_TC_DEBUG=$m VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times\|ms_screen" | tail -2 | cut -c1-330
done

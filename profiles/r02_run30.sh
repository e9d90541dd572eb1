#!/bin/bash
# round 2, GPU call 30: where the main sweep's time goes -- the profiling modes of the screening kernel
# (VVB_TC_DEBUG=3: no TMEM drain, 1: drain only, 2: drain + minima, no admissions; results are garbage in these modes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 0 3 1 2; do
echo "== VVB_TC_DEBUG=$m"
VVB_TC_DEBUG=$m VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times\|ms_screen" | tail -2 | cut -c1-330
done

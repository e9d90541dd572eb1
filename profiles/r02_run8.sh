#!/bin/bash
# round 2, GPU call 8: L2 hints with the persisting set-aside enabled (one wave, 500k x 2000 database)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for hint in 0 25 45 70; do
VVB_TC_VERBOSE=1 VVB_TC_L2HINT=$hint REPS=3 timeout -s KILL 300 python profiles/r02_prof_wide.py > gpurun_out/r8_wide_hint$hint.txt 2>&1
echo "hint=$hint rc=$? $(grep vvb gpurun_out/r8_wide_hint$hint.txt | head -1) $(tail -1 gpurun_out/r8_wide_hint$hint.txt | cut -c90-200)"
done

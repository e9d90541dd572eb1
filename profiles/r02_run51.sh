#!/bin/bash
# round 2, GPU call 51: launch list (device time per kernel) of one C3-sized step on one GPU: 10M x 50, k=100
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r51_launches_c3_10M.csv python profiles/r02c_c3_launches.py > gpurun_out/r51_ncu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/r51_ncu.log | cut -c1-400

#!/bin/bash
# round 2, GPU call 17: one-product pre-pass, slack sweep (VVB_TC_PRE_SLACK_LOG2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for sl in 12 13 14 16; do
VVB_TC_PRE_SLACK_LOG2=$sl timeout -s KILL 300 python profiles/run_screen.py 1000000 > gpurun_out/r17_screen_slack$sl.txt 2>&1
echo "slack=2^-$sl rc=$? $(tail -2 gpurun_out/r17_screen_slack$sl.txt | head -1 | cut -c90-420)"
VVB_TC_PRE_SLACK_LOG2=$sl VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 1000000 2>&1 | grep "vvb times" | tail -1
done

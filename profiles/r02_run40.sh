#!/bin/bash
# round 2, GPU call 40: pre-pass without its TMA loads (VVB_TC_DEBUG=7, no drain either): is the one-product pre-pass
# bound by the depth of the stage ring (4 x 16 KB in flight per SM)?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for m in 6 7; do
echo "== VVB_TC_DEBUG=$m"
VVB_TC_DEBUG=$m VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 200000 2>&1 | grep "vvb times" | tail -1 | cut -c1-330
done

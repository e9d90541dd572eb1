#!/bin/bash
# round 2, GPU call 39: what bounds the one-product pre-pass (864 clocks per tile against a 512-clock MMA floor)?
# VVB_TC_DEBUG=5 drains half of every pre-pass accumulator, 6 none (results garbage; only the pre-pass clocks count)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in 0 5 6; do
echo "== VVB_TC_DEBUG=$m"
VVB_TC_DEBUG=$m VVB_TC_TIMES=1 timeout -s KILL 300 python profiles/run_screen.py 200000 2>&1 | grep "vvb times" | tail -1 | cut -c1-330
done

#!/bin/bash
# round 2, GPU call 28: lean inner loop of the smoother (parity incl. stress + timing)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -x -q -m gpu -k "smooth or narrow or column or c1_ or pinned or golden or publish" > gpurun_out/r28_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r28_tests.log)"
for j in 0 2; do
VVB_SMOOTH_JITTER=$j timeout -s KILL 300 python profiles/r02_smooth_probe.py 2>&1 | grep jitter
done

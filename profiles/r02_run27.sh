#!/bin/bash
# round 2, GPU call 27: compute-sanitizer (memcheck, racecheck) over every kernel at small sizes, if the pool allows it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 120 python profiles/sanitize_small.py > gpurun_out/r27_plain.log 2>&1
echo "plain rc=$? $(tail -1 gpurun_out/r27_plain.log | cut -c1-200)"
for tool in memcheck racecheck; do
timeout -s KILL 600 compute-sanitizer --tool $tool --error-exitcode 7 python profiles/sanitize_small.py > gpurun_out/r27_$tool.log 2>&1
echo "$tool rc=$?"; grep -c "ERROR SUMMARY\|Error\|error" gpurun_out/r27_$tool.log; tail -4 gpurun_out/r27_$tool.log | cut -c1-300
done

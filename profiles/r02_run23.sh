#!/bin/bash
# round 2, GPU call 23: the whole -m gpu suite, smoke(), the default bench line and the reference arm at N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r23_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/r23_tests.log)"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r23_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/r23_smoke.log)"
timeout -s KILL 900 python bench.py > gpurun_out/r23_bench_1gpu.json 2> gpurun_out/r23_bench_1gpu.err
echo "bench rc=$?"; tail -1 gpurun_out/r23_bench_1gpu.json | cut -c1-300
timeout -s KILL 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r23_bench_reference.json 2> gpurun_out/r23_bench_reference.err
echo "reference rc=$?"; tail -1 gpurun_out/r23_bench_reference.json | cut -c1-300

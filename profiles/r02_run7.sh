#!/bin/bash
# round 2, GPU call 7: L2 hints on the wide-row kernel, ncu captures (wide-row kernel; 1M x 50 screening kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for hint in 0 60 85 100; do
VVB_TC_L2HINT=$hint timeout -s KILL 300 python - <<PY > gpurun_out/r7_c5_hint$hint.txt 2>&1
import sys, torch, json
sys.path.insert(0, ".")
from bench import synth_cells_device
from vivae_b200.device import knn_device
dev = torch.device("cuda", 0)
x = synth_cells_device(2_000_000, 2000, dev)
for _ in range(2):
    st = {}
    knn_device(x, 50, 0, 62_500, stats=st)
    print(json.dumps(st))
PY
echo "hint=$hint rc=$? $(tail -1 gpurun_out/r7_c5_hint$hint.txt | cut -c1-330)"
done
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader
REPS=2 timeout -s KILL 300 python profiles/r02_prof_wide.py > gpurun_out/r7_prof_wide_plain.log 2>&1 &&
REPS=2 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:knn_screen -s 1 -c 1 -o gpurun_out/r7_wide python profiles/r02_prof_wide.py > gpurun_out/r7_ncu_wide.log 2>&1
echo "ncu wide rc=$?"; tail -2 gpurun_out/r7_prof_wide_plain.log | cut -c1-300
timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 > gpurun_out/r7_bench_plain.log 2>&1 &&
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"knn_screen|smooth_sweep|knn_rerank" -s 6 -c 3 -o gpurun_out/r7_c2 python bench.py --steps 1 --warmup 1 --skip-e2e --skip-fit --skip-cpu --skip-dense --skip-c3 --skip-c5 > gpurun_out/r7_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"; tail -2 gpurun_out/r7_ncu_c2.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep

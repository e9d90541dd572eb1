#!/bin/bash
# round 2, GPU call 5: 256-row tiles (WIDE) in the single-product first tier
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "screening_error or wide_rows or c5_slice or few_row or refine_pass" -s > gpurun_out/r5_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r5_tests.log
grep -E "stats|screening error|passed|failed|rc=|Error" gpurun_out/r5_tests.log | cut -c1-420
for wide in 1 0; do
VVB_TC_WIDE=$wide timeout -s KILL 300 python - <<PY > gpurun_out/r5_c5_wide$wide.txt 2>&1
import sys, torch, json
sys.path.insert(0, ".")
from bench import synth_cells_device
from vivae_b200.device import knn_device
dev = torch.device("cuda", 0)
x = synth_cells_device(2_000_000, 2000, dev)
for _ in range(2):
    st = {}
    knn_device(x, 50, 0, 62_500, stats=st)   # one 32nd of C5: the share of one GPU of 8 is 250k rows
    print(json.dumps(st))
PY
echo "wide=$wide rc=$?"; tail -1 gpurun_out/r5_c5_wide$wide.txt | cut -c1-500
done

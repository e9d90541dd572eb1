#!/bin/bash
# round 2, GPU call 6: sliced refine launch, quartet kernel, full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r6_tests.log
tail -6 gpurun_out/r6_tests.log | cut -c1-300
timeout -s KILL 300 python - <<PY > gpurun_out/r6_c5_share.txt 2>&1
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
echo "c5 rc=$?"; tail -1 gpurun_out/r6_c5_share.txt | cut -c1-500
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --skip-c3 > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open("gpurun_out/r6_bench.json") if l.startswith("{")][-1])
print("ms/step", round(j["ms_per_step"],2), "knn", round(j["ms_knn"],2), "smooth", round(j["ms_smooth"],2), {k:round(v,2) for k,v in j["knn_stats"].items() if k.startswith("ms_")}, "e2e", round(j["e2e"]["ms_per_step"],1), "quartet", j["quartet"]["ms_fwd_bwd"], j["quartet"]["frac"], "fit", j["fit"]["s_per_epoch"])
PY

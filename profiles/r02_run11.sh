#!/bin/bash
# round 2, GPU call 11: fused training step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_fused_step.py -m gpu -q --tb=short > gpurun_out/r11_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r11_tests.log | cut -c1-300
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fit or graph or pipeline or smoke or geometric" > gpurun_out/r11_tests2.log 2>&1
echo "tests2 rc=$?"; tail -8 gpurun_out/r11_tests2.log | cut -c1-300
timeout -s KILL 600 python - <<'PY' 2>&1 | tail -5
import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import vivae_b200 as vv
from bench import synth_cells
x = synth_cells(1_000_000, 50)
for fused in ("1", "0"):
    import os
    os.environ["VIVAE_FUSED_STEP"] = fused
    m = vv.ViVAE(input_dim=50, random_state=1)
    m.fit(x, n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2): l = m.train_epoch(lam_mds=100.0)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("fused", fused, "s/epoch", round(dt, 4), "ms/step", round(dt / 1954 * 1e3, 4), l)
PY

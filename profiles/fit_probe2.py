"""Kernel census of the CUDA-graphed training step (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_cells
import vivae_b200 as vv
from torch.profiler import profile, ProfilerActivity
x = synth_cells(5120 * 4, 50)
m = vv.ViVAE(input_dim=50, random_state=1)
m.fit(x, n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
m2 = m
steps = x.shape[0] // 512
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m2.data_loader  # reuse
    losses = m2.train_epoch(lam_mds=100.0)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.count for e in ev)
print("device kernels per step", tot / steps, "device time per step (us)", sum(e.device_time_total for e in ev) / steps)
for e in sorted(ev, key=lambda e: -e.count)[:40]:
    print(f"{e.count / steps:6.1f}/step  {e.device_time_total / max(1, e.count):6.2f} us  {e.key[:100]}")

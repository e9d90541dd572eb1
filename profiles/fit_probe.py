import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from bench import synth_cells
import vivae_b200 as vv
x = synth_cells(200_000, 50)
for fused in ("0", "1"):
    os.environ["VVB_ADAM_FUSED"] = fused
    m = vv.ViVAE(input_dim=50, random_state=1)
    m.fit(x, n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.fit(x, n_epochs=2, batch_size=512, lam_mds=100.0, verbose=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("fused", fused, "s/epoch", dt, "ms/step", dt / (200_000 // 512) * 1e3)
# kernel count of one graphed step
from torch.profiler import profile, ProfilerActivity
m = vv.ViVAE(input_dim=50, random_state=1); m.use_cuda_graph = False
m.fit(x[:5120], n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.fit(x[:5120], n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages()]
tot = sum(e.count for e in ev)
print("kernels per 10 steps", tot)
for e in sorted(ev, key=lambda e: -e.count)[:25]:
    print(e.count, round(e.device_time_total / max(1, e.count), 2), e.key[:90])

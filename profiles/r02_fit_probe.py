"""Where the time of one training step goes (B200): the fused kernel pair, Adam, the captured graph, the loader."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vivae_b200 as vv  # noqa: E402
from bench import synth_cells  # noqa: E402
from vivae_b200.fused_step import FusedStep  # noqa: E402

x = synth_cells(200_000, 50)
m = vv.ViVAE(input_dim=50, random_state=1)
m.fit(x, n_epochs=1, batch_size=512, lam_mds=100.0, verbose=False)
dev = torch.device("cuda")
xb = torch.from_numpy(x[:512]).to(dev)
fs = FusedStep(m.net, 512)
sums = torch.zeros(6, device=dev)


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, (time.perf_counter() - t0) / n * 1e6


print("fused kernel pair (+randn): gpu us %.1f, wall us %.1f" % timeit(lambda: fs.step(xb, sums, 1.0, 1.0, 100.0)))
print("adam step: gpu us %.1f, wall us %.1f" % timeit(lambda: m.optimizer.step()))
g = m._graph
print("graph nodes of this library:", g[4])
print("graph replay: gpu us %.1f, wall us %.1f" % timeit(lambda: g[1].replay()))
loader = m.data_loader
perm = torch.randperm(200_000, device=dev)
print("index_select+copy: gpu us %.1f, wall us %.1f" % timeit(lambda: g[2].copy_(loader.data.index_select(0, perm[:512]))))
t0 = time.perf_counter()
m.train_epoch(lam_mds=100.0)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("epoch 200k rows: %.4f s, %.1f us per step" % (dt, dt / 391 * 1e6))

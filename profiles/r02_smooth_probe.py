"""Timing probes of the Gauss-Seidel sweep (VVB_SMOOTH_JITTER=2: no polls, no fence; 3: one poll per neighbour, never
waits; 4: polls and waits but no fence -- results of 2/3/4 are NOT valid, they only price the parts of the protocol)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_cells  # noqa: E402
from vivae_b200.device import knn_device, smooth_device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
x = torch.from_numpy(synth_cells(n, 50)).cuda()
idx, _ = knn_device(x, 50)
for mode in ("gauss_seidel", "jacobi"):
    for _ in range(3):
        smooth_device(x, idx, k=50, coef=1.0, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        smooth_device(x, idx, k=50, coef=1.0, mode=mode)
    e1.record()
    torch.cuda.synchronize()
    print(f"jitter={os.environ.get('VVB_SMOOTH_JITTER', '0')} {mode}: {e0.elapsed_time(e1) / 10:.3f} ms")

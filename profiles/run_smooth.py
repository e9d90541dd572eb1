"""Timing driver for the smoother: Gauss-Seidel (parity) vs Jacobi on a real exact kNN graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_cells
from vivae_b200.device import knn_device, smooth_device
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = 50
x = torch.from_numpy(synth_cells(n, 50)).cuda()
idx, _ = knn_device(x, k)
def t(mode, reps=5):
    for _ in range(2): smooth_device(x, idx, k=k, coef=1.0, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): smooth_device(x, idx, k=k, coef=1.0, mode=mode)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b = n * (k * 50 * 4 + k * 4 + 50 * 4)
for mode in ("gauss_seidel", "jacobi"):
    ms = t(mode)
    print(f"{mode}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s (algorithmic)  rpt={os.environ.get('VVB_SMOOTH_RPT', 'default')}")

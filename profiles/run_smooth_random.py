"""Smoother timing at scale with a random lower-triangular-ish neighbour graph (cost does not depend on
which rows are gathered, the Gauss-Seidel waits do depend on how close in index the neighbours are)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vivae_b200.device import smooth_device
n = int(sys.argv[1]); k = int(sys.argv[2])
x = torch.randn(n, 50, device="cuda")
idx = torch.randint(0, n, (n, k + 1), device="cuda", dtype=torch.int64)
idx[:, 0] = torch.arange(n, device="cuda")
for mode in ("gauss_seidel", "jacobi"):
    for _ in range(2): smooth_device(x, idx, k=k, coef=1.0, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); smooth_device(x, idx, k=k, coef=1.0, mode=mode); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1); b = n * (k * 200 + k * 4 + 200)
    print(f"n={n} k={k} {mode}: {ms:.2f} ms, {b / ms / 1e6:.0f} GB/s algorithmic")

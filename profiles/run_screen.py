"""Small driver used for the ncu captures in profiles/: one make_knn + smooth on synthetic cells."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_cells
from vivae_b200.device import knn_device, smooth_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
x = torch.from_numpy(synth_cells(n, 50)).cuda()
for _ in range(2):
    st = {}
    idx, dist = knn_device(x, 50, stats=st)
    out = smooth_device(x, idx, k=50, coef=1.0)
torch.cuda.synchronize()
print(st)

"""Small driver used for the ncu captures in profiles/: one make_knn + smooth on synthetic cells."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import synth_cells
from vivae_b200.device import knn_device, smooth_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 50
x = torch.from_numpy(synth_cells(n, d)).cuda()
for _ in range(2):
    st = {}
    idx, dist = knn_device(x, 50, stats=st)
    out = smooth_device(x, idx, k=50, coef=1.0)
torch.cuda.synchronize()
print(st)
if st.get('ms_screen'):
    print(f'screen: {2.0 * n * n * d / st["ms_screen"] / 1e9:.1f} algorithmic TFLOP/s, executed MACs/pair = {3 * ((d + 3 + 63) // 64) * 64 if d > 61 else 16 * ((d + 18) // 16 + 2 * ((d + 15) // 16))}')

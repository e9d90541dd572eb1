"""Config C3 (10M x 50, k=100) on one GPU: one graph build + smoother sweep, for the ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_pca_device
from vivae_b200.device import knn_device, smooth_device
dev = torch.device("cuda", 0)
x = synth_pca_device(10_000_000, 50, dev)
st = {}
idx, dist = knn_device(x, 100, stats=st)
out = smooth_device(x, idx, k=100, coef=1.0)
torch.cuda.synchronize()
print(st)

"""Quartet loss forward + backward at a bandwidth-sized batch (B = 2^20, D = 50, L = 2), as bench.py times it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vivae_b200 import _lib
lib = _lib.load()
B, d = 1 << 20, 50
dev = torch.device("cuda", 0)
x = torch.randn(B, d, device=dev)
z = torch.randn(B, 2, device=dev)
gu = torch.empty(B, 2, device=dev); gz = torch.empty(B, 2, device=dev)
loss = torch.empty((), device=dev); gout = torch.ones((), device=dev)
ws = torch.empty(int(lib.vvb_quartet_workspace_bytes(B)), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def step():
    _lib.check(lib.vvb_quartet_loss_fwd_device(x.data_ptr(), z.data_ptr(), B, d, 2, None, 1, 0, 0, loss.data_ptr(), gu.data_ptr(), ws.data_ptr(), st))
    _lib.check(lib.vvb_quartet_loss_bwd_device(gu.data_ptr(), gout.data_ptr(), B * 2, gz.data_ptr(), st))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"fwd+bwd {ms*1e3:.1f} us, {B*(4*(d+2)+4*6)/ms/1e6:.0f} GB/s, loss {loss.item():.6f}")

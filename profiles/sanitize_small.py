"""Small invocation of every kernel of libvivae_b200.so (written for compute-sanitizer, which is closed on this pool;
still useful as a quick all-kernel self-check):
   compute-sanitizer --tool memcheck  --error-exitcode 9 python profiles/sanitize_small.py
   compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/sanitize_small.py
Sizes are tiny (the tools slow kernels down 10-100x); results are still checked against each other."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vivae_b200 as vv
from vivae_b200.device import knn_device, smooth_device
from vivae_b200.sharded import _DeviceOps, ring_fold, shard_bounds

which = sys.argv[1:] or ["knn", "kloop", "smooth", "quartet", "ring"]
rng = np.random.default_rng(0)


def blobs(n, d):
    c = rng.standard_normal((5, d)) * 4
    return (c[rng.integers(0, 5, n)] + rng.standard_normal((n, d))).astype(np.float32)


if "knn" in which:
    x = torch.tensor(blobs(3000, 50), device="cuda")
    st = {}
    it, dt = knn_device(x, 20, algo="tensor", stats=st)
    ib, db = knn_device(x, 20, algo="brute")
    assert torch.equal(it, ib) and torch.equal(dt, db), "tensor vs brute"
    i2, _ = knn_device(x, 20, 500, 1000, algo="auto")
    assert torch.equal(i2, ib[500:1500])
    print("knn ok", st["tiles_swept"], st["rows_fallback"], flush=True)
if "kloop" in which:
    x = torch.tensor(blobs(1500, 200), device="cuda")
    it, dt = knn_device(x, 10, algo="tensor")
    ib, db = knn_device(x, 10, algo="brute")
    assert torch.equal(it, ib) and torch.equal(dt, db), "k-loop tensor vs brute"
    print("kloop ok", flush=True)
if "smooth" in which:
    xh = blobs(2000, 50)
    idx, dist = vv.make_knn(xh, k=15, verbose=False)
    a = vv.smooth(xh, [idx, dist], k=15, coef=0.5, n_iter=2)
    b = vv.smooth(xh.astype(np.float64), [idx, dist], k=15, coef=0.5, n_iter=1, mode="jacobi")
    assert np.isfinite(a).all() and np.isfinite(b).all()
    print("smooth ok", flush=True)
if "quartet" in which:
    xb = torch.tensor(blobs(515, 50), device="cuda")
    for hd, ld in (("euclidean", "euclidean"), ("cosine", "cosine")):
        z = torch.randn(515, 2, device="cuda", requires_grad=True)
        loss = vv.MDSLoss()(xb, z, distf_hd=hd, distf_ld=ld, n_sampling=2)
        loss.backward()
        assert torch.isfinite(loss) and torch.isfinite(z.grad).all()
    print("quartet ok", flush=True)
if "ring" in which:
    n, k, world = 2400, 30, 3
    xd = torch.tensor(blobs(n, 20), device="cuda")
    ops = _DeviceOps()
    wi, wd = ops.knn(xd, k)
    for rank in range(world):
        q0, nq = shard_bounds(n, world, rank)
        own = xd[q0:q0 + nq].contiguous()
        ri, rd = ops.init(nq, k, q0, xd.device)
        for s in range(1, world):
            s0, sn = shard_bounds(n, world, (rank + s) % world)
            ring_fold(ops, own, q0, [(s0, xd[s0:s0 + sn].contiguous())], ri, rd, with_own=(s == 1))
        assert torch.equal(ri, wi[q0:q0 + nq]) and torch.equal(ops.finish(rd), wd[q0:q0 + nq])
    print("ring ok", flush=True)
torch.cuda.synchronize()
print("all ok")

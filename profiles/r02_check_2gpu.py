"""Round 2, N > 1 check (run under torchrun, one rank per GPU): the cell plan prepared in row/cell shards over NCCL
(vvb_knn_plan_phase_device + three exchanges) and the host-sharded pipeline, against the single-GPU build."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_cells  # noqa: E402
from vivae_b200.device import knn_device, smooth_device  # noqa: E402
from vivae_b200.sharded import knn_smooth_host_sharded, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {}
for n, d, k in ((1_000_000, 50, 50), (200_000, 300, 30)):
    x = synth_cells(n, d, seed=3)
    xd = torch.from_numpy(x).to(dev)
    q0, nq = shard_bounds(n, world, rank)
    s_plain, s_group = {}, {}
    i0, d0 = knn_device(xd, k, q0, nq, stats=s_plain)                         # whole preparation on this GPU
    i1, d1 = knn_device(xd, k, q0, nq, stats=s_group, group=dist.group.WORLD)  # preparation split over the ranks
    same = bool(torch.equal(i0, i1) and torch.equal(d0, d1))
    for _ in range(2):
        s_group = {}
        knn_device(xd, k, q0, nq, stats=s_group, group=dist.group.WORLD)
        s_plain = {}
        knn_device(xd, k, q0, nq, stats=s_plain)
    hi, hd, hs = knn_smooth_host_sharded(torch.from_numpy(x[q0:q0 + nq]).pin_memory(), n, k, coef=1.0, n_iter=1,
                                         _knn_fn=lambda xx, kk, a, b, alg: knn_device(xx, kk, a, b, group=dist.group.WORLD))
    full_idx, _ = knn_device(xd, k)
    sm = smooth_device(xd, full_idx, k=k, coef=1.0, n_iter=1)
    pipe_same = bool(torch.equal(hi, i0.cpu()) and torch.equal(hs, sm[q0:q0 + nq].cpu()))
    flag = torch.tensor([int(same and pipe_same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out[f"{n}x{d}"] = {"identical_on_all_ranks": bool(flag.item()), "prepare_ms_whole": s_plain["ms_prepare"],
                       "prepare_ms_sharded": s_group["ms_prepare"], "screen_ms": s_group["ms_screen"],
                       "tiles_swept_whole_plan": s_plain["tiles_swept"], "tiles_swept_sharded_plan": s_group["tiles_swept"],
                       "parts": s_group["prepare_parts"]}
# ---- the Gauss-Seidel sweep over all GPUs at once (interleaved rows, pushes over NVLink peer memory)
from vivae_b200.sharded import release_peer_smoothers, smooth_peer_sharded  # noqa: E402

peer = {}
for n, d, k, dtype, order in ((1_000_000, 50, 50, torch.float32, "random"), (200_000, 38, 100, torch.float64, "sorted"),
                              (50_000, 7, 10, torch.float32, "random")):
    x = synth_cells(n, d, seed=5)
    if order == "sorted":
        x = x[np.argsort(x[:, 0])].copy()
    xd32 = torch.from_numpy(x).to(dev)
    idx, _ = knn_device(xd32, k)
    xd = xd32.to(dtype)
    res = {}
    for n_iter in (1, 2):
        want = smooth_device(xd, idx, k=k, coef=0.7, n_iter=n_iter)
        got = smooth_peer_sharded(xd, idx, k=k, coef=0.7, n_iter=n_iter)
        res[f"identical_n_iter{n_iter}"] = bool(torch.equal(want, got))
    for name, fn in (("replicated", lambda: smooth_device(xd, idx, k=k, coef=1.0, n_iter=1)),
                     ("peer", lambda: smooth_peer_sharded(xd, idx, k=k, coef=1.0, n_iter=1))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[f"ms_{name}"] = float(t.item())
    flag = torch.tensor([int(all(v for kk, v in res.items() if kk.startswith("identical")))], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["identical_on_all_ranks"] = bool(flag.item())
    peer[f"{n}x{d}_k{k}_{str(dtype).split('.')[-1]}_{order}"] = res
release_peer_smoothers()
if rank == 0:
    print(json.dumps({"world": world, "checks": out, "peer_smoother": peer}))
dist.barrier()
dist.destroy_process_group()

"""Correctness of the NCCL paths on real GPUs (run under torchrun, 2+ ranks):
   make_knn_ring == whole-matrix build for this rank's rows; smooth_jacobi_sharded == single-GPU Jacobi smooth;
   knn_smooth_sharded == single-GPU Gauss-Seidel pipeline.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 profiles/ring_check.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import vivae_b200 as vv
from vivae_b200.device import knn_device, smooth_device
from vivae_b200.sharded import knn_smooth_sharded, make_knn_ring, make_knn_sharded, shard_bounds, smooth_jacobi_sharded
from bench import synth_cells

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for n, d, k, ties in ((300_001, 50, 50, False), (40_000, 6, 20, True), (2_000, 40, 700, False)):
    if ties:
        x = np.random.default_rng(5).integers(0, 4, size=(n, d)).astype(np.float32)
    else:
        x = synth_cells(n, d)
    xd = torch.from_numpy(x).to(dev)
    q0, nq = shard_bounds(n, world, rank)
    whole_idx, whole_dist = knn_device(xd, k)
    for rep in range(2):  # twice: buffer reuse across calls
        ri, rd = make_knn_ring(xd[q0:q0 + nq].contiguous(), n, k)
        good = torch.equal(ri, whole_idx[q0:q0 + nq]) and torch.equal(rd, whole_dist[q0:q0 + nq])
        ok &= good
        print(f"rank {rank}: ring n={n} d={d} k={k} ties={ties} rep={rep}: {'ok' if good else 'MISMATCH'}", flush=True)
    si, sd = make_knn_sharded(xd, k)
    good = torch.equal(si, whole_idx) and torch.equal(sd, whole_dist)
    ok &= good
    kk = min(k, 50)
    idx_local, _ = make_knn_sharded(xd, k, gather=False)
    jac = smooth_jacobi_sharded(xd, idx_local, kk, coef=0.5, n_iter=2)
    want = smooth_device(xd, whole_idx, k=kk, coef=0.5, n_iter=2, mode="jacobi")
    good2 = torch.equal(jac, want)
    _, gs = knn_smooth_sharded(xd, k, coef=1.0, n_iter=1)
    good3 = torch.equal(gs, smooth_device(xd, whole_idx, k=k, coef=1.0, n_iter=1))
    ok &= good2 and good3
    print(f"rank {rank}: sharded graph {'ok' if good else 'MISMATCH'}, jacobi shards {'ok' if good2 else 'MISMATCH'}, "
          f"gauss-seidel pipeline {'ok' if good3 else 'MISMATCH'}", flush=True)
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if int(t.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)

"""world_size-2 and -3 `gloo` tests of the multi-GPU host logic (row shards, ragged all-gather,
replicated smoother).  The CUDA kernels are replaced by the CPU oracle through the `_knn_fn` /
`_smooth_fn` hooks: what is under test is the sharding and the collective plumbing."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from vivae_b200.sharded import all_gather_rows, knn_smooth_sharded, make_knn_sharded, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_knn(x, k, q0, nq, algo):
    i, d = orc.knn(x.numpy(), k, q0, q0 + nq)
    return torch.from_numpy(i), torch.from_numpy(d)


def _oracle_smooth(x, idx, k, coef, n_iter):
    return torch.from_numpy(orc.smooth(x.numpy(), idx.numpy(), k, coef, n_iter))


def _worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.from_numpy(np.random.default_rng(0).standard_normal((n, 6)).astype(np.float32))
        idx, dst = make_knn_sharded(x, k, _knn_fn=_oracle_knn)
        idx2, xs = knn_smooth_sharded(x, k, coef=1.0, n_iter=2, _knn_fn=_oracle_knn, _smooth_fn=_oracle_smooth)
        q0, nq = shard_bounds(n, world, rank)
        loc = torch.arange(q0, q0 + nq, dtype=torch.int64).reshape(-1, 1)
        rows = all_gather_rows(loc, n)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx.numpy(), dist=dst.numpy(), idx2=idx2.numpy(),
                 xs=xs.numpy(), rows=rows.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 301), (3, 100), (2, 5)])
def test_sharded_graph_and_smooth_equal_single_process(tmp_path, world, n):
    k = 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    x = np.random.default_rng(0).standard_normal((n, 6)).astype(np.float32)
    wi, wd = orc.knn(x, k)
    ws = orc.smooth(x, wi, k, 1.0, 2)
    for r in range(world):
        g = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(g["idx"], wi) and np.array_equal(g["dist"], wd)       # identical bytes on every rank
        assert np.array_equal(g["idx2"], wi) and np.array_equal(g["xs"], ws)
        assert np.array_equal(g["rows"].ravel(), np.arange(n))


def test_shard_bounds_cover_all_rows():
    for n in (1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for q0, c in spans:
                assert q0 == min(pos, n) and c >= 0
                pos += c

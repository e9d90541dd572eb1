"""world_size-2 and -3 `gloo` tests of the multi-GPU host logic (row shards, ragged all-gather,
replicated smoother).  The CUDA kernels are replaced by the CPU oracle through the `_knn_fn` /
`_smooth_fn` hooks: what is under test is the sharding and the collective plumbing."""

import os
import socket
from fractions import Fraction

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as orc
from vivae_b200.sharded import (all_gather_rows, col_bounds, knn_smooth_host_sharded, knn_smooth_sharded, make_knn_ring,
                                make_knn_sharded, shard_bounds, smooth_cols_sharded, smooth_jacobi_sharded)


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


def _oracle_smooth_cols(x, idx, c0, w, k, coef, n_iter, mode):
    """The oracle's sweep on a column slice (what vvb_smooth_cols_device computes): full-width result tensor with
    only the slice filled."""
    out = torch.full_like(x, float("nan"))
    if w:
        xs = np.ascontiguousarray(x.numpy()[:, c0:c0 + w])
        out[:, c0:c0 + w] = torch.from_numpy(orc.smooth(xs, idx.numpy(), k, coef, n_iter))
    return out


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
        # the smoother sharded by COLUMNS (the reference's sweep does not split by rows, but never mixes columns)
        cs = smooth_cols_sharded(x, idx, k, coef=0.6, n_iter=2, _cols_fn=_oracle_smooth_cols)
        # the host pipeline: every rank passes only its rows, gets its rows of the graph and of the smoothed matrix
        hi, hd, hs = knn_smooth_host_sharded(x[q0:q0 + nq].numpy().copy(), n, k, coef=1.0, n_iter=2, _knn_fn=_oracle_knn,
                                             _smooth_fn=_oracle_smooth, device=torch.device("cpu"))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), idx=idx.numpy(), dist=dst.numpy(), idx2=idx2.numpy(),
                 xs=xs.numpy(), cs=cs.numpy(), rows=rows.numpy(), hi=hi.numpy(), hd=hd.numpy(), hs=hs.numpy(), q0=q0, nq=nq)
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
        assert np.array_equal(g["cs"], orc.smooth(x, wi, k, 0.6, 2))  # column slices, gathered == the whole sweep
        q0, nq = int(g["q0"]), int(g["nq"])
        assert np.array_equal(g["hi"], wi[q0:q0 + nq]) and np.array_equal(g["hd"], wd[q0:q0 + nq])
        assert np.array_equal(g["hs"], ws[q0:q0 + nq])


def test_col_bounds_cover_all_columns_in_whole_pairs():
    for d in (1, 2, 7, 38, 50, 51, 2000):
        for world in (1, 2, 3, 8, 64):
            spans = [col_bounds(d, world, r) for r in range(world)]
            assert sum(w for _, w in spans) == d
            pos = 0
            for c0, w in spans:
                assert c0 == min(pos, d) and w >= 0 and (d % 2 or (c0 % 2 == 0 and w % 2 == 0))
                pos += w


def test_shard_bounds_cover_all_rows():
    for n in (1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for q0, c in spans:
                assert q0 == min(pos, n) and c >= 0
                pos += c


# ---------------------------------------------------------------------------- database-sharded ring build
class _CpuRingOps:
    """CPU restatements of the four device entry points the ring build calls (csrc/knn_merge.cu)."""

    @staticmethod
    def knn(xc, k, q0, nq, algo):
        return _oracle_knn(xc, k, q0, nq, algo)

    @staticmethod
    def init(nq, k, self0, device):
        run_idx = torch.full((nq, k + 1), -1, dtype=torch.int64)
        run_d2 = torch.full((nq, k + 1), float("inf"), dtype=torch.float32)
        run_idx[:, 0] = torch.arange(self0, self0 + nq)
        run_d2[:, 0] = 0.0
        return run_idx, run_d2

    @staticmethod
    def merge(xc, q0, nq, new_idx, keep_lo, keep_hi, global_off, run_idx, run_d2):
        x = xc.numpy()
        k = run_idx.shape[1] - 1
        for t in range(nq):
            ents = [(float(run_d2[t, c]), int(run_idx[t, c])) for c in range(1, k + 1) if run_idx[t, c] >= 0]
            for j in new_idx[t, 1:].tolist():
                if keep_lo <= j < keep_hi:
                    acc = np.float32(0.0)
                    for c in range(x.shape[1]):  # the contract's chain with an exactly rounded fmaf (exact rationals)
                        df = Fraction(float(np.float32(x[q0 + t, c] - x[j, c])))
                        acc = orc._round_to_f32(df * df + Fraction(float(acc)))
                    ents.append((float(acc), j + global_off))
            ents.sort()
            for c in range(1, k + 1):
                if c - 1 < len(ents):
                    run_d2[t, c], run_idx[t, c] = ents[c - 1][0], ents[c - 1][1]
                else:
                    run_d2[t, c], run_idx[t, c] = float("inf"), -1

    @staticmethod
    def finish(run_d2):
        return torch.from_numpy(np.sqrt(run_d2.numpy()))  # hardware sqrt: correctly rounded


def _ring_data(n, kind):
    rng = np.random.default_rng(7)
    if kind == "ties":  # integer lattice with duplicate rows: every tie-break rule is exercised across shards
        x = rng.integers(0, 3, size=(n, 3)).astype(np.float32)
    else:
        x = rng.standard_normal((n, 5)).astype(np.float32)
    return x


def _ring_worker(rank, world, port, n, k, kind, out_dir, per_sweep=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.from_numpy(_ring_data(n, kind))
        q0, nq = shard_bounds(n, world, rank)
        idx, dst = make_knn_ring(x[q0:q0 + nq].clone(), n, k, _ops=_CpuRingOps(), shards_per_sweep=per_sweep)
        np.savez(os.path.join(out_dir, f"ring{rank}.npz"), idx=idx.numpy(), dist=dst.numpy(), q0=q0, nq=nq)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,k,kind,per_sweep", [(2, 150, 7, "normal", None), (3, 100, 5, "ties", 1),
                                                      (3, 100, 5, "ties", None), (2, 90, 60, "ties", 1),
                                                      (3, 7, 4, "normal", 1), (4, 61, 9, "ties", 2),
                                                      (4, 61, 9, "normal", 1)])
def test_ring_build_equals_whole_matrix_build(tmp_path, world, n, k, kind, per_sweep):
    """Shards passed around the ring + running merge == one build over the whole matrix, byte for byte
    (k larger than a shard, ragged and short shards, ties and duplicate rows across shard borders; one shard per
    sweep, two, and as many as fit -- on the CPU: all of them)."""
    port = _free_port()
    mp.spawn(_ring_worker, args=(world, port, n, k, kind, str(tmp_path), per_sweep), nprocs=world, join=True)
    wi, wd = orc.knn(_ring_data(n, kind), k)
    for r in range(world):
        g = np.load(tmp_path / f"ring{r}.npz")
        q0, nq = int(g["q0"]), int(g["nq"])
        assert np.array_equal(g["idx"], wi[q0:q0 + nq])
        assert np.array_equal(g["dist"], wd[q0:q0 + nq])


# ---------------------------------------------------------------------------- row-sharded Jacobi smoother
def _oracle_jacobi_rows(x, idx_rows, k, coef, row_lo):
    """CPU stand-in for `smooth_rows_device`: one Jacobi sweep, rows [row_lo, row_lo+len(idx_rows))."""
    n = x.shape[0]
    full = np.tile(np.arange(n, dtype=np.int64)[:, None], (1, idx_rows.shape[1]))  # other rows: self only (unused)
    full[row_lo:row_lo + idx_rows.shape[0]] = idx_rows.numpy()
    out = orc.smooth(x.numpy(), full, k, coef, 1, mode=1)
    return torch.from_numpy(out[row_lo:row_lo + idx_rows.shape[0]].copy())


def _jacobi_worker(rank, world, port, n, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.from_numpy(np.random.default_rng(3).standard_normal((n, 5)).astype(np.float32))
        idx_local, _ = make_knn_sharded(x, k, gather=False, _knn_fn=_oracle_knn)
        out = smooth_jacobi_sharded(x, idx_local, k, coef=0.6, n_iter=3, _rows_fn=_oracle_jacobi_rows)
        np.save(os.path.join(out_dir, f"jac{rank}.npy"), out.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 203), (3, 50)])
def test_row_sharded_jacobi_smoother_equals_single_process(tmp_path, world, n):
    k = 6
    port = _free_port()
    mp.spawn(_jacobi_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    x = np.random.default_rng(3).standard_normal((n, 5)).astype(np.float32)
    wi, _ = orc.knn(x, k)
    want = orc.smooth(x, wi, k, 0.6, 3, mode=1)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"jac{r}.npy"), want)

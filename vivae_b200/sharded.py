"""Row-sharded kNN graph + smoothing over several GPUs (one process per GPU, torch.distributed).

Two ways to build the graph:
  * `make_knn_sharded`  -- every rank holds the whole matrix (2 GB at 10M x 50, 16 GB at 2M x 2000: it fits
    a 180 GB B200 many times over) and computes its query rows; no data-path collective.
  * `make_knn_ring`     -- every rank holds ONLY its rows; the shards travel around an NCCL ring
    (send to rank-1 / receive from rank+1 over NVLink, overlapped with the step's kernels) and each rank
    folds what passes by into running top-k lists.  For matrices beyond one GPU's memory; bit-identical.

The graph builder shards by query rows with no data-path collective: every rank holds the whole
matrix and computes rows [q0, q0+nq) of the graph; neighbour indices are global.  The only
exchange is the all-gather of the index shards, needed because the Gauss-Seidel smoother reads
the whole graph -- and does not itself shard by rows (row i depends on updated rows j < i), so it
runs replicated on every rank ("replicas only").
"""

from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "all_gather_rows", "all_gather_indices", "make_knn_sharded", "make_knn_ring", "ring_fold",
           "knn_smooth_sharded", "knn_smooth_host_sharded", "smooth_jacobi_sharded", "smooth_cols_sharded", "smooth_auto_sharded", "col_bounds"]


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous query-row shard of `rank`: (first row, row count); the last shards may be short or empty."""
    per = (n + world - 1) // world
    q0 = min(rank * per, n)
    return q0, min(per, n - q0)


def col_bounds(d: int, world: int, rank: int) -> Tuple[int, int]:
    """Column slice of `rank` for the column-sharded smoother: (first column, column count).  Slices are made of whole
    column PAIRS when d is even (the kernel then fetches two columns per load); trailing slices may be empty."""
    unit = 2 if d % 2 == 0 else 1
    units = d // unit
    per = (units + world - 1) // world
    c0 = min(rank * per, units)
    return c0 * unit, min(per, units - c0) * unit


def smooth_cols_sharded(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1, group=None,
                        mode: str = "gauss_seidel", _cols_fn: Optional[Callable] = None) -> torch.Tensor:
    """`smooth` of the whole matrix with the COLUMNS split over the ranks of `group`.

    The reference's sweep (knn.py:173-185) is sequential in the rows (row i reads rows j < i already updated), so it
    does not split by rows -- but it never mixes columns: element (i, c) depends on column c of row i's neighbours
    only.  Every rank therefore runs the complete sweep (all n_iter passes, same row order, same arithmetic) on its
    slice of the columns, and the slices are all-gathered once at the end.  Bit-identical to the single-GPU call.
    `x` (n, d) and `idx` (n, >= k) are the whole matrix and the whole graph on every rank."""
    if _cols_fn is None:
        from .device import smooth_cols_device

        _cols_fn = smooth_cols_device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n, d = x.shape
    c0, w = col_bounds(d, world, rank)
    out = _cols_fn(x, idx, c0, w, k, coef, n_iter, mode)
    if world == 1:
        return out
    wmax = col_bounds(d, world, 0)[1]
    mine = torch.zeros((n, wmax), dtype=x.dtype, device=x.device)
    mine[:, :w] = out[:, c0:c0 + w]
    flat = torch.empty((world * n, wmax), dtype=x.dtype, device=x.device)  # rank-major concatenation
    dist.all_gather_into_tensor(flat, mine, group=group)
    allc = flat.view(world, n, wmax)
    for r in range(world):
        r0, rw = col_bounds(d, world, r)
        if rw and r != rank:
            out[:, r0:r0 + rw] = allc[r, :, :rw]
    return out


def smooth_auto_sharded(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1, group=None,
                        mode: str = "gauss_seidel") -> torch.Tensor:
    """The smoother of the multi-GPU pipelines: column shards when a slice still keeps a warp busy (at least 64
    columns per rank), otherwise the whole sweep on every rank.  Measured on 2 B200 at 1M x 50, k=50: 3.06 ms in
    two column shards against 2.59 ms replicated -- at that width the sweep costs one load instruction per neighbour
    whether the row has 50 columns or 26, and it is bound by those instructions and their latency, not by bytes."""
    from .device import smooth_device

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1 and x.shape[1] >= 64 * world:
        return smooth_cols_sharded(x, idx, k, coef, n_iter, group, mode)
    return smooth_device(x, idx, k, coef, n_iter, mode)


def all_gather_rows(local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """Concatenate the ranks' row shards (laid out by `shard_bounds`) into the full (n, ...) tensor on every rank."""
    world = dist.get_world_size(group)
    per = (n + world - 1) // world
    if local.shape[0] < per:  # ragged tail: pad to the common shard height
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        local = pad
    full = torch.empty((per * world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full[:n]


def all_gather_indices(idx_local: torch.Tensor, n: int, group=None) -> torch.Tensor:
    """All-gather of neighbour-index shards: the indices cross NVLink as int32 (n < 2^31: half the bytes of the
    reference's int64) and are widened back on arrival -- the smoother reads int64 like `knn.py:137` returns."""
    if n <= 0x7FFFFFFF:
        return all_gather_rows(idx_local.to(torch.int32), n, group).to(torch.int64)
    return all_gather_rows(idx_local, n, group)


def make_knn_sharded(x: torch.Tensor, k: int, group=None, algo: str = "auto", gather: bool = True,
                     _knn_fn: Optional[Callable] = None):
    """Exact kNN graph with query rows sharded over the ranks of `group`.

    `x` is the whole (n, d) matrix, resident on this rank's device.  Returns `(idx, dist)` for all rows
    (`gather=True`, identical on every rank) or for this rank's shard only."""
    if _knn_fn is None:
        from .device import knn_device

        _knn_fn = knn_device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = x.shape[0]
    q0, nq = shard_bounds(n, world, rank)
    idx, dst = _knn_fn(x, k, q0, nq, algo)
    if gather and world > 1:
        idx = all_gather_indices(idx, n, group)
        dst = all_gather_rows(dst, n, group)
    return idx, dst


_COLS = object()  # marker: the default smoother of the multi-GPU pipelines (column-sharded, CUDA)


class _DeviceOps:
    """The CUDA entry points the ring build is made of (tests substitute CPU restatements under gloo)."""

    def __init__(self):
        from . import device

        self.knn, self.init, self.merge, self.finish = (device.knn_device, device.knn_merge_init, device.knn_merge,
                                                        device.knn_merge_finish)


def ring_fold(ops, x_local: torch.Tensor, q0: int, shard: torch.Tensor, s0: int, run_idx: torch.Tensor,
              run_d2: torch.Tensor, algo: str = "auto") -> None:
    """One ring step: fold database rows [s0, s0+len(shard)) into the running lists of query rows
    [q0, q0+len(x_local)).  `s0 == q0` means the shard IS the rank's own rows (step 0)."""
    nq, sn, k = x_local.shape[0], shard.shape[0], run_idx.shape[1] - 1
    if nq == 0 or sn == 0:
        return
    if s0 == q0:
        xc, qoff, keep_lo, keep_hi, goff = x_local, 0, 0, nq, q0
    elif s0 < q0:
        xc, qoff, keep_lo, keep_hi, goff = torch.cat([shard, x_local]), sn, 0, sn, s0
    else:
        xc, qoff, keep_lo, keep_hi, goff = torch.cat([x_local, shard]), 0, nq, nq + sn, s0 - nq
    kn = min(k, xc.shape[0] - 1)
    if kn >= 1:
        new_idx, _ = ops.knn(xc, kn, qoff, nq, algo)
        ops.merge(xc, qoff, nq, new_idx, keep_lo, keep_hi, goff, run_idx, run_d2)


def make_knn_ring(x_local: torch.Tensor, n: int, k: int, group=None, algo: str = "auto", _ops=None):
    """Exact kNN graph of an (n, d) matrix whose rows are spread over the ranks: `x_local` holds this rank's rows
    [q0, q0+nq) = `shard_bounds(n, world, rank)`.  Returns `(idx, dist)` of those rows, global indices.

    `world` ring steps.  In step s the rank holds the shard of rank (rank+s) % world; it forwards it to rank-1
    while the graph builder runs on [shard ; own rows] stacked in global row order (so the builder's tie-break by
    local index is the contract's tie-break by global index), and the merge kernel folds the shard's rows of the
    result into the running lists (`csrc/knn_merge.cu` explains why that is exact).  Step 0 ranks the own rows."""
    ops = _ops or _DeviceOps()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if k < 1 or k > n - 1:
        raise ValueError("`k` must be between 1 and (n-1)")
    q0, nq = shard_bounds(n, world, rank)
    if x_local.shape[0] != nq:
        raise ValueError(f"rank {rank} must hold rows [{q0}, {q0 + nq}) of the matrix, got {x_local.shape[0]} rows")
    per = (n + world - 1) // world
    d = x_local.shape[1]
    run_idx, run_d2 = ops.init(nq, k, q0, x_local.device)
    cur = torch.zeros((per, d), dtype=x_local.dtype, device=x_local.device)
    cur[:nq] = x_local
    nxt = torch.empty_like(cur)
    for s in range(world):
        s0, sn = shard_bounds(n, world, (rank + s) % world)
        reqs = []
        if s + 1 < world:  # pass the shard on while this step computes on it
            to, frm = (rank - 1) % world, (rank + 1) % world
            if group is not None:
                to, frm = dist.get_global_rank(group, to), dist.get_global_rank(group, frm)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, cur, to, group), dist.P2POp(dist.irecv, nxt, frm, group)])
        ring_fold(ops, x_local, q0, cur[:sn], s0, run_idx, run_d2, algo)
        for r in reqs:
            r.wait()
        cur, nxt = nxt, cur
    return run_idx, ops.finish(run_d2)


def knn_smooth_sharded(x: torch.Tensor, k: int, coef: float = 1.0, n_iter: int = 1, group=None, algo: str = "auto",
                       _knn_fn: Optional[Callable] = None, _smooth_fn: Optional[Callable] = None):
    """make_knn (sharded by query rows) -> all-gather of the indices -> smooth (`smooth_auto_sharded`: column shards
    for wide rows, replicated otherwise; tests pass `_smooth_fn` to run it on the CPU).  Returns (idx, smoothed x)."""
    if _smooth_fn is None:
        _smooth_fn = _COLS
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    idx, _ = make_knn_sharded(x, k, group=group, algo=algo, gather=False, _knn_fn=_knn_fn)
    if world > 1:
        idx = all_gather_indices(idx, x.shape[0], group)
    if _smooth_fn is _COLS:
        return idx, smooth_auto_sharded(x, idx, k, coef, n_iter, group)
    return idx, _smooth_fn(x, idx, k, coef, n_iter)


def knn_smooth_host_sharded(x_shard, n: int, k: int, coef: float = 1.0, n_iter: int = 1, group=None, algo: str = "auto",
                            out=None, _knn_fn: Optional[Callable] = None, _smooth_fn: Optional[Callable] = None,
                            device=None):
    """The whole path with HOST buffers, one process per GPU: every rank passes ONLY its rows of the matrix.

    `x_shard`: this rank's rows [q0, q0+nq) = `shard_bounds(n, world, rank)` as a host array (NumPy or CPU
    tensor, float32; page-locked memory makes the upload one DMA).  Steps: upload of the shard -> NCCL all-gather
    of the matrix -> graph rows of the shard -> all-gather of the indices (int32 over NVLink) -> Gauss-Seidel
    smoother (`smooth_auto_sharded`: column shards for wide rows, replicated otherwise) -> download of the shard's graph rows and smoothed rows.
    Returns host tensors `(idx int64 (nq, k+1), dist float32 (nq, k+1), smoothed (nq, d))`; with `out=` (a tuple of
    three page-locked CPU tensors of those shapes) the downloads land there.  world = 1 is the plain pipeline."""
    if _knn_fn is None:
        from .device import knn_device

        _knn_fn = knn_device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    q0, nq = shard_bounds(n, world, rank)
    xs = x_shard if isinstance(x_shard, torch.Tensor) else torch.from_numpy(x_shard)
    if xs.shape[0] != nq:
        raise ValueError(f"rank {rank} must pass rows [{q0}, {q0 + nq}) of the matrix, got {xs.shape[0]} rows")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    x_loc = xs.to(device, non_blocking=True)
    x = all_gather_rows(x_loc, n, group) if world > 1 else x_loc
    idx_loc, dist_loc = _knn_fn(x, k, q0, nq, algo)
    idx = all_gather_indices(idx_loc, n, group) if world > 1 else idx_loc
    if _smooth_fn is None:
        sm = smooth_auto_sharded(x, idx, k, coef, n_iter, group)
    else:
        sm = _smooth_fn(x, idx, k, coef, n_iter)
    sm_loc = sm[q0:q0 + nq]
    if out is None:
        out = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=(device.type == "cuda"))
                    for t in (idx_loc, dist_loc, sm_loc))
    out[0].copy_(idx_loc, non_blocking=True)
    out[1].copy_(dist_loc, non_blocking=True)
    out[2].copy_(sm_loc, non_blocking=True)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    return out


def smooth_jacobi_sharded(x: torch.Tensor, idx_local: torch.Tensor, k: int, coef: float = 0.1, n_iter: int = 1, group=None,
                          _rows_fn: Optional[Callable] = None) -> torch.Tensor:
    """`smooth(..., mode='jacobi')` with the rows of every sweep split over the ranks.

    `x` is the whole (n, d) matrix on every rank; `idx_local` holds the neighbour lists of this rank's rows
    (`shard_bounds`), e.g. straight from `make_knn_sharded(..., gather=False)` -- the graph itself is never
    gathered.  Each sweep: a rank computes its rows from the current `x`, then the new rows are all-gathered
    (n*d*4 bytes per sweep, against n*(k+1)*8 for the graph).  This is the order-independent variant, NOT the
    reference's in-place sweep (that one runs replicated: `knn_smooth_sharded`)."""
    if _rows_fn is None:
        from .device import smooth_rows_device

        _rows_fn = smooth_rows_device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = x.shape[0]
    q0, nq = shard_bounds(n, world, rank)
    if idx_local.shape[0] != nq:
        raise ValueError(f"rank {rank} must hold the neighbour lists of rows [{q0}, {q0 + nq})")
    if n_iter < 1 or n_iter > 10000:
        raise ValueError("`n_iter` must be at least 1 and at most 10000")
    cur = x
    for _ in range(n_iter):
        new_rows = _rows_fn(cur, idx_local, k, coef, q0)
        cur = all_gather_rows(new_rows, n, group) if world > 1 else new_rows
    return cur

"""Row-sharded kNN graph + smoothing over several GPUs (one process per GPU, torch.distributed).

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

__all__ = ["shard_bounds", "all_gather_rows", "make_knn_sharded", "knn_smooth_sharded"]


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous query-row shard of `rank`: (first row, row count); the last shards may be short or empty."""
    per = (n + world - 1) // world
    q0 = min(rank * per, n)
    return q0, min(per, n - q0)


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
        idx = all_gather_rows(idx, n, group)
        dst = all_gather_rows(dst, n, group)
    return idx, dst


def knn_smooth_sharded(x: torch.Tensor, k: int, coef: float = 1.0, n_iter: int = 1, group=None, algo: str = "auto",
                       _knn_fn: Optional[Callable] = None, _smooth_fn: Optional[Callable] = None):
    """make_knn (sharded) -> all-gather of the indices -> smooth (replicated).  Returns (idx, smoothed x)."""
    if _smooth_fn is None:
        from .device import smooth_device

        _smooth_fn = smooth_device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    idx, _ = make_knn_sharded(x, k, group=group, algo=algo, gather=False, _knn_fn=_knn_fn)
    if world > 1:
        idx = all_gather_rows(idx, x.shape[0], group)
    return idx, _smooth_fn(x, idx, k, coef, n_iter)

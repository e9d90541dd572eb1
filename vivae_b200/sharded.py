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
           "knn_smooth_sharded", "knn_smooth_host_sharded", "smooth_jacobi_sharded", "smooth_cols_sharded", "smooth_auto_sharded", "smooth_peer_sharded", "PeerSmoother", "col_bounds"]


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


class PeerSmoother:
    """The reference's Gauss-Seidel sweep split over the GPUs of `group` by INTERLEAVED rows, the exchange fused into
    the kernel (`csrc/smooth_peer.cu`): every rank keeps a complete copy of the sweep's output in memory its peers
    write over NVLink (CUDA IPC) -- one 64-bit word {value, sweep epoch} per element, so a value and its "published"
    mark arrive in one store and nothing needs a fence; the owner of a row pushes it into all copies, rows that need a
    neighbour still in flight re-read its words.  This object owns the buffers of one (n, d, dtype) and the mappings
    of the peers' buffers.  Bit-identical to `smooth_device`."""

    def __init__(self, n: int, d: int, dtype: torch.dtype, device: torch.device, group=None):
        import ctypes

        from . import _lib

        self.lib = _lib.load()
        self._check = _lib.check
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n, self.d, self.dtype, self.device = int(n), int(d), dtype, device
        self.itemsize = torch.empty((), dtype=dtype).element_size()
        nbytes = self.n * self.d * 8 * (self.itemsize // 4)  # one word per float32, two per float64
        self.local, self.opened = [], []
        with torch.cuda.device(device):
            ptr, h = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
            self._check(self.lib.vvb_peer_alloc(nbytes, ctypes.byref(ptr), h))
            self.local.append(ptr.value)
            mine = torch.frombuffer(bytearray(h), dtype=torch.uint8).to(device)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine, group=group)
            allh = allh.cpu()
            self.ptrs = [0] * self.world  # every rank's buffer as mapped into this process
            err = None
            for r in range(self.world):
                if r == self.rank:
                    self.ptrs[r] = self.local[0]
                    continue
                p_ = ctypes.c_void_p()
                raw = (ctypes.c_ubyte * 64)(*allh[r].tolist())
                if err is None and self.lib.vvb_peer_open(raw, ctypes.byref(p_)) != 0:
                    err = _lib.last_error()
                self.ptrs[r] = p_.value or 0
                if p_.value:
                    self.opened.append(p_.value)
            # every rank must agree on whether the mapping worked (a rank that gave up would leave the others waiting)
            ok = torch.tensor([0.0 if err else 1.0], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if float(ok.item()) < 1.0:
                for q_ in self.opened:
                    self.lib.vvb_peer_close(q_)
                dist.barrier(group=group)
                for q_ in self.local:
                    self.lib.vvb_peer_free(q_)
                self.local, self.opened = [], []
                raise RuntimeError(f"peer mapping of the smoother's buffers failed on some rank ({err or 'another rank'})")
            self.table = (ctypes.c_void_p * self.world)(*self.ptrs)
            self.ws = torch.zeros(int(self.lib.vvb_smooth_peer_workspace_bytes()), dtype=torch.uint8, device=device)
            self.token = torch.zeros(1, dtype=torch.float32, device=device)
        self.epoch = 0

    def _barrier(self):
        dist.all_reduce(self.token, group=self.group)  # stream-ordered: no rank passes it before every rank's stream got here

    def sweep(self, x: torch.Tensor, idx: torch.Tensor, k: int, coef: float, n_iter: int = 1) -> torch.Tensor:
        if x.shape != (self.n, self.d) or x.dtype != self.dtype or not x.is_contiguous() or not idx.is_contiguous():
            raise ValueError("PeerSmoother: x / idx do not match the buffers")
        if n_iter < 1 or n_iter > 10000:
            raise ValueError("`n_iter` must be at least 1 and at most 10000")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            cur = x
            for _ in range(n_iter):
                self.epoch += 1
                self._barrier()  # every rank has unpacked the previous sweep: the copies may be overwritten
                self._check(self.lib.vvb_smooth_peer_device(cur.data_ptr(), self.itemsize, self.n, self.d, idx.data_ptr(),
                                                            idx.shape[1], int(k), float(coef), self.epoch, self.rank,
                                                            self.world, self.table, self.ws.data_ptr(), stream))
                self._barrier()  # every rank's pushes have landed in this rank's copy
                out = torch.empty_like(x)
                self._check(self.lib.vvb_smooth_peer_unpack_device(self.local[0], self.itemsize, self.n * self.d, self.epoch,
                                                                   out.data_ptr(), self.ws.data_ptr(), stream))
                cur = out
            bad = int(self.ws[:4].view(torch.int32).item())
        if bad == 1:
            raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
        if bad:
            raise RuntimeError("peer smoother: a row never arrived in this rank's copy")
        return cur

    def close(self):
        """Unmap the peers' buffers and free the local one (all ranks, after a barrier: nobody may still push)."""
        if not self.local:
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p_ in self.opened:
                self.lib.vvb_peer_close(p_)
            dist.barrier(group=self.group)
            for p_ in self.local:
                self.lib.vvb_peer_free(p_)
        self.local, self.opened = [], []


_PEER_SMOOTHERS: dict = {}
_PEER_BROKEN = False


def smooth_peer_sharded(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1,
                        group=None) -> torch.Tensor:
    """`smooth` (the reference's in-place sweep) of the whole matrix on all GPUs of `group` at once -- see
    `PeerSmoother`.  `x` (n, d) and `idx` (n, >= k) are the whole matrix and graph on every rank; every rank gets the
    whole smoothed matrix.  The buffers are kept per (n, d, dtype, group) until `release_peer_smoothers()`."""
    key = (x.shape[0], x.shape[1], x.dtype, x.device.index, id(group))
    ps = _PEER_SMOOTHERS.get(key)
    if ps is None:
        for old in list(_PEER_SMOOTHERS.values()):  # one set of buffers at a time: they are as large as the matrix
            old.close()
        _PEER_SMOOTHERS.clear()
        ps = _PEER_SMOOTHERS[key] = PeerSmoother(x.shape[0], x.shape[1], x.dtype, x.device, group)
    return ps.sweep(x.contiguous(), idx.contiguous(), k, coef, n_iter)


def release_peer_smoothers() -> None:
    for ps in list(_PEER_SMOOTHERS.values()):
        ps.close()
    _PEER_SMOOTHERS.clear()


def smooth_auto_sharded(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1, group=None,
                        mode: str = "gauss_seidel") -> torch.Tensor:
    """The smoother of the multi-GPU pipelines: column shards when a slice still keeps a warp busy (at least 64 columns
    per rank), else the whole sweep on every rank.  The interleaved-row sweep over NVLink peer memory (`PeerSmoother`;
    k <= 128, d <= 128, up to 8 ranks) is bit-identical but measured SLOWER than the replicated sweep on 2 B200 (8.5 ms
    against 2.6 ms at 1M x 50, `profiles/r02_peer_smoother_2gpu.json`): a row waiting on a row still in flight on the
    other GPU waits an NVLink round trip, and every gather reads 8-byte {value, epoch} words.  It is therefore opt-in
    (`VVB_SMOOTH_PEER=1`).  Measured on 2 B200 at 1M x 50, k=50: 3.06 ms in
    two column shards against 2.59 ms replicated -- at that width the sweep costs one load instruction per neighbour
    whether the row has 50 columns or 26, and it is bound by those instructions and their latency, not by bytes."""
    import os
    import warnings

    from .device import smooth_device

    global _PEER_BROKEN
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n, d = x.shape
    if (world > 1 and world <= 8 and mode == "gauss_seidel" and k <= 128 and d <= 128 and n >= 4096 and x.is_cuda
            and not _PEER_BROKEN and os.environ.get("VVB_SMOOTH_PEER", "0") == "1"):
        try:
            return smooth_peer_sharded(x, idx, k, coef, n_iter, group)
        except (RuntimeError, NotImplementedError) as exc:  # e.g. no peer access between the GPUs: stay replicated
            _PEER_BROKEN = True
            warnings.warn(f"vivae_b200: peer-memory smoother unavailable ({exc}); running the sweep on every rank")
    if world > 1 and d >= 64 * world:
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


def ring_fold(ops, x_local: torch.Tensor, q0: int, shards, run_idx: torch.Tensor, run_d2: torch.Tensor,
              algo: str = "auto", with_own: bool = False) -> None:
    """One sweep of the ring build: fold the database shards `shards` = [(first global row, rows tensor), ...] into the
    running lists of query rows [q0, q0+len(x_local)).  The builder runs on the shards and the own rows stacked in
    GLOBAL row order (its tie-break by local index is then the contract's tie-break by global index); every contiguous
    range of global indices is folded by one merge call.  `with_own`: the rank's own rows are folded in as well (the
    sweep covers own x own anyway, so no sweep is spent on the own rows alone).  A shard that starts at `q0` IS the
    rank's own rows (a one-rank "ring")."""
    nq, k = x_local.shape[0], run_idx.shape[1] - 1
    shards = [(s0, t) for s0, t in shards if t.shape[0] > 0 and s0 != q0]
    if nq == 0 or (not shards and not with_own):
        return
    pieces = sorted(shards + [(q0, x_local)], key=lambda p_: p_[0])
    xc = torch.cat([t for _, t in pieces]) if len(pieces) > 1 else x_local
    keeps, qoff, pos = [], 0, 0
    for s0, t in pieces:
        if s0 == q0:
            qoff = pos
            if with_own:
                keeps.append((pos, pos + nq, q0 - pos))
        else:
            keeps.append((pos, pos + t.shape[0], s0 - pos))
        pos += t.shape[0]
    kn = min(k, xc.shape[0] - 1)
    if kn >= 1:
        new_idx, _ = ops.knn(xc, kn, qoff, nq, algo)
        for keep_lo, keep_hi, goff in keeps:  # (rows of xc to keep, global index of xc row 0 for that range)
            ops.merge(xc, qoff, nq, new_idx, keep_lo, keep_hi, goff, run_idx, run_d2)


def _ring_shards_per_sweep(per: int, d: int, device: torch.device, world: int) -> int:
    """How many foreign shards one sweep of the ring build may hold: as many as fit in about a third of the free device
    memory next to the own rows (matrix rows, fp16 hi/lo operand rows, cell-sorted order), at least one.  A sweep pruned
    by the cell bounds costs about the same per query block whatever the size of its database, so fewer, larger sweeps
    are cheaper: with everything in memory the ring is one sweep, i.e. the replicated layout."""
    if device.type != "cuda":
        return world - 1
    free, _ = torch.cuda.mem_get_info(device)
    ka = (d + 3 + 63) // 64
    per_row = 4 * d + 512 * ka + 64
    return max(1, min(world - 1, int(free / 3 // (per_row * max(per, 1))) - 1))


def make_knn_ring(x_local: torch.Tensor, n: int, k: int, group=None, algo: str = "auto", _ops=None,
                  shards_per_sweep: Optional[int] = None):
    """Exact kNN graph of an (n, d) matrix whose rows are spread over the ranks: `x_local` holds this rank's rows
    [q0, q0+nq) = `shard_bounds(n, world, rank)`.  Returns `(idx, dist)` of those rows, global indices.

    `world` ring steps.  In step s the rank holds the shard of rank (rank+s) % world and forwards it to rank-1.  The
    graph builder runs on [shards ; own rows] stacked in global row order (so the builder's tie-break by local index is
    the contract's tie-break by global index) once `shards_per_sweep` foreign shards have arrived (default: as many
    as fit in memory, `_ring_shards_per_sweep`; 1 = a sweep per ring step, overlapped with the next transfer), and the
    merge kernel folds the shards' rows of the result into the running lists (`csrc/knn_merge.cu` explains why that is
    exact).  The own rows are ranked by the first sweep, which covers them anyway."""
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
    if world == 1:
        ring_fold(ops, x_local, q0, [], run_idx, run_d2, algo, with_own=True)
        return run_idx, ops.finish(run_d2)
    m = shards_per_sweep if shards_per_sweep else _ring_shards_per_sweep(per, d, x_local.device, world)
    m = max(1, min(int(m), world - 1))
    if not shards_per_sweep and m >= world - 1:
        # Everything fits next to the own rows: ONE sweep ranks them against the whole matrix in global row order, so
        # there is nothing to merge (running lists exist to carry results from one sweep to the next) and nothing to
        # pass on -- the shards arrive with one all-gather instead of N - 1 ring rounds, and the cell plan is prepared
        # across the ranks like in the replicated layout.  Round 2b ran this case through the ring machinery (N - 1
        # waited transfers, a stacked copy, N merge calls that each re-ranked k rows per query): 1.29x / 1.48x the
        # replicated layout at 1M x 50 on 2 / 4 GPUs.  An explicit `shards_per_sweep` keeps the ring schedule.
        x_all = all_gather_rows(x_local, n, group)
        if _ops is None:
            return ops.knn(x_all, k, q0, nq, algo, group=(group if group is not None else dist.group.WORLD))
        return ops.knn(x_all, k, q0, nq, algo)
    cur = torch.zeros((per, d), dtype=x_local.dtype, device=x_local.device)
    cur[:nq] = x_local
    nxt = torch.empty_like(cur)
    held, first = [], True
    for s in range(world):
        s0, sn = shard_bounds(n, world, (rank + s) % world)
        reqs = []
        if s + 1 < world:  # pass the shard on while this step computes on it
            to, frm = (rank - 1) % world, (rank + 1) % world
            if group is not None:
                to, frm = dist.get_global_rank(group, to), dist.get_global_rank(group, frm)
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, cur, to, group), dist.P2POp(dist.irecv, nxt, frm, group)])
        if s > 0:
            # (a held shard is copied out of the exchange buffers unless it is swept right away)
            held.append((s0, cur[:sn] if m == 1 else cur[:sn].clone()))
            if len(held) == m or s == world - 1:
                ring_fold(ops, x_local, q0, held, run_idx, run_d2, algo, with_own=first)
                held, first = [], False
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

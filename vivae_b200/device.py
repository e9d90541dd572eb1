"""Device-resident entry points (torch CUDA tensors in, torch CUDA tensors out).

Thin wrappers over the `*_device` functions of the C ABI: torch is used only for
device memory and the current stream.  These are additive to the reference API;
`vv.make_knn` / `vv.smooth` (host NumPy in/out) are the drop-in surface.
"""

from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

__all__ = ["knn_device", "smooth_device", "smooth_cols_device", "smooth_rows_device", "knn_merge_init", "knn_merge", "knn_merge_finish"]


def _require_cuda(name: str, t: torch.Tensor, dtype) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"`{name}` must be a CUDA tensor (vivae_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"`{name}` must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


_XCH_DTYPES = {_lib.XCH_I32: torch.int32, _lib.XCH_U32: torch.int32, _lib.XCH_F64: torch.float64}
_ORD_FLIP = -2147483648  # flips the top bit: unsigned order of the uint32 images == signed order of the flipped int32


def _xch_view(ws: torch.Tensor, e) -> torch.Tensor:
    """Typed view of the table a plan phase reported (a device pointer INTO the workspace tensor)."""
    dt = _XCH_DTYPES[e.dtype]
    off = e.ptr - ws.data_ptr()
    nbytes = e.count * torch.empty((), dtype=dt).element_size()
    if off < 0 or off + nbytes > ws.numel():
        raise RuntimeError("knn plan: exchange table outside the workspace")
    return ws[off:off + nbytes].view(dt)


def _plan_phase(lib, x, k, nq, ws, phase, part, parts):
    n, d = x.shape
    xch = (_lib.Exchange * 2)()
    m = ctypes.c_int(0)
    rc = lib.vvb_knn_plan_phase_device(x.data_ptr(), n, d, k, nq, ws.data_ptr(), ws.numel(), phase, part, parts,
                                       torch.cuda.current_stream().cuda_stream, xch, ctypes.byref(m))
    if rc == _lib.ENOTSUP:
        return None
    _lib.check(rc)
    return [(_xch_view(ws, xch[i]), xch[i].op) for i in range(m.value)]


def _plan_over_group(lib, x, k, nq, ws, group) -> bool:
    """Phases 0..3 of the preparation on this rank's workspace, the three exchanges over `group` in between
    (include/vivae_b200.h: vvb_knn_plan_phase_device).  False: the shape is outside the tensor path."""
    import torch.distributed as dist

    from .sharded import all_gather_rows, shard_bounds

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0]
    for phase in range(4):
        tables = _plan_phase(lib, x, k, nq, ws, phase, rank, world)
        if tables is None:
            return False
        for view, op in tables:
            if op == _lib.XCH_GATHER_ROWS:
                r0, rn = shard_bounds(n, world, rank)
                view.copy_(all_gather_rows(view[r0:r0 + rn], n, group))
            elif op == _lib.XCH_SUM:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
            else:  # XCH_MAX_ORD
                view.bitwise_xor_(_ORD_FLIP)
                dist.all_reduce(view, op=dist.ReduceOp.MAX, group=group)
                view.bitwise_xor_(_ORD_FLIP)
    return True


def _plan_emulated(lib, x, k, nq, ws, parts: int) -> bool:
    """The same protocol with all `parts` ranks played by this process (tests; one GPU): one workspace per part,
    the collectives restated with torch ops.  Leaves the finished plan of part 0 in `ws`."""
    from .sharded import shard_bounds

    n = x.shape[0]
    wss = [ws] + [_lib.raw_empty(ws.numel(), torch.uint8, ws.device) for _ in range(parts - 1)]
    for phase in range(4):
        per_part = [_plan_phase(lib, x, k, nq, wss[p], phase, p, parts) for p in range(parts)]
        if per_part[0] is None:
            return False
        for t in range(len(per_part[0])):
            views, op = [pp[t][0] for pp in per_part], per_part[0][t][1]
            if op == _lib.XCH_GATHER_ROWS:
                full = torch.empty_like(views[0])
                for p in range(parts):
                    r0, rn = shard_bounds(n, parts, p)
                    full[r0:r0 + rn] = views[p][r0:r0 + rn]
            elif op == _lib.XCH_SUM:
                full = torch.stack(views).sum(0)
            else:
                full = torch.stack([v.bitwise_xor(_ORD_FLIP) for v in views]).max(0).values.bitwise_xor(_ORD_FLIP)
            for v in views:
                v.copy_(full)
    return True


def knn_device(x: torch.Tensor, k: int, q0: int = 0, nq: Optional[int] = None, algo: str = "auto",
               stats: Optional[dict] = None, group=None, _emulate_parts: int = 0):
    """Exact (k+1)-NN of rows [q0, q0+nq) of `x` against all rows of `x`, on `x`'s device.

    Returns `(idx int64 (nq, k+1), dist float32 (nq, k+1))` CUDA tensors; enqueued on the
    current stream (synchronises only when `stats` is requested).

    `group`: a `torch.distributed` process group whose ranks all call this with the SAME matrix and their own
    query rows (`sharded.make_knn_sharded`): the preparation of the cell plan is then split across the ranks (pivot
    assignment by rows, cell extents by cells, three small NCCL exchanges) instead of being repeated on every GPU.
    Every rank must pass the same `nq` class (the workspace layout depends on min(nq, 1,048,576))."""
    lib = _lib.load()
    x = _require_cuda("x", x, torch.float32)
    n, d = x.shape
    nq = n - q0 if nq is None else nq
    if k < 1 or k > n - 1:
        raise ValueError("`k` must be between 1 and (n-1)")
    if algo not in _lib.KNN_ALGOS:
        raise ValueError(f"`algo` must be one of {sorted(_lib.KNN_ALGOS)}")
    a = _lib.KNN_ALGOS[algo]
    world = 1
    if group is not None:
        import torch.distributed as dist

        world = dist.get_world_size(group)
    with torch.cuda.device(x.device):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vvb_make_knn_workspace_bytes(n, d, k, nq, a, ctypes.byref(nbytes)))
        ws = _lib.raw_empty(max(int(nbytes.value), 256), torch.uint8, x.device)
        idx = _lib.raw_empty((nq, k + 1), torch.int64, x.device)
        dist_ = _lib.raw_empty((nq, k + 1), torch.float32, x.device)
        st = _lib.KnnStats() if stats is not None else None
        stream = torch.cuda.current_stream().cuda_stream
        planned, ms_plan = False, 0.0
        if (world > 1 or _emulate_parts > 1) and algo != "brute":
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if stats is not None else None
            if ev:
                ev[0].record()
            planned = (_plan_over_group(lib, x, k, nq, ws, group) if world > 1 else
                       _plan_emulated(lib, x, k, nq, ws, _emulate_parts))
            if ev:
                ev[1].record()
                ev[1].synchronize()
                ms_plan = ev[0].elapsed_time(ev[1])
        if nq == 0:  # an empty shard took part in the exchanges above and has nothing to build
            if stats is not None:
                stats.update(st.as_dict())
            return idx, dist_
        if planned:
            _lib.check(lib.vvb_make_knn_planned_device(x.data_ptr(), n, d, k, q0, nq, idx.data_ptr(), dist_.data_ptr(),
                                                       ws.data_ptr(), ws.numel(), stream,
                                                       ctypes.byref(st) if st is not None else None))
        else:
            _lib.check(lib.vvb_make_knn_device(x.data_ptr(), n, d, k, q0, nq, a, idx.data_ptr(), dist_.data_ptr(),
                                               ws.data_ptr(), ws.numel(), stream,
                                               ctypes.byref(st) if st is not None else None))
        ws.record_stream(torch.cuda.current_stream())
    if stats is not None:
        stats.update(st.as_dict())
        stats["ms_prepare"] += ms_plan
        stats["prepare_parts"] = world if planned and world > 1 else (_emulate_parts if planned else 1)
    return idx, dist_


def smooth_device(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1,
                  mode: str = "gauss_seidel", validate: bool = True) -> torch.Tensor:
    """`smooth` on device-resident data: `x` (n, d) float32/float64, `idx` (n, >=k) int64.

    The kernel never dereferences an out-of-range neighbour index (it substitutes the row itself and raises a
    flag).  `validate=True` reads that flag back (one 4-byte copy, synchronises the stream) and raises the
    reference's `ValueError` (knn.py:58-61) like the host path does; `validate=False` keeps the call asynchronous
    and returns `(out, flag)` with `flag` a 1-element int32 CUDA tensor the caller checks when it chooses."""
    lib = _lib.load()
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError("`x` must be float32 or float64")
    x = _require_cuda("x", x, x.dtype)
    idx = _require_cuda("idx", idx, torch.int64)
    n, d = x.shape
    if mode not in _lib.SMOOTH_MODES:
        raise ValueError(f"`mode` must be one of {sorted(_lib.SMOOTH_MODES)}")
    with torch.cuda.device(x.device):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vvb_smooth_workspace_bytes(x.element_size(), n, d, n_iter, ctypes.byref(nbytes)))
        ws = _lib.raw_empty(int(nbytes.value), torch.uint8, x.device)
        out = _lib.raw_empty(x.shape, x.dtype, x.device)
        _lib.check(lib.vvb_smooth_device(x.data_ptr(), x.element_size(), n, d, idx.data_ptr(), idx.shape[1], int(k),
                                         float(coef), int(n_iter), _lib.SMOOTH_MODES[mode], out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        ws.record_stream(torch.cuda.current_stream())
        flag = ws[:4].view(torch.int32)
        if not validate:
            return out, flag.clone()
        if int(flag.item()):
            raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
    return out


def smooth_cols_device(x: torch.Tensor, idx: torch.Tensor, col0: int, dcols: int, k: int = 50, coef: float = 0.1,
                       n_iter: int = 1, mode: str = "gauss_seidel", out: Optional[torch.Tensor] = None,
                       validate: bool = True) -> torch.Tensor:
    """Columns [col0, col0+dcols) of `smooth_device(x, idx, ...)`, written into the same columns of `out` (a full
    (n, d) tensor, allocated when None; the other columns are left alone).  The smoother never mixes columns, so the
    slice is bit-identical to the full call's -- this is how several GPUs share one Gauss-Seidel sweep."""
    lib = _lib.load()
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError("`x` must be float32 or float64")
    x = _require_cuda("x", x, x.dtype)
    idx = _require_cuda("idx", idx, torch.int64)
    n, d = x.shape
    if mode not in _lib.SMOOTH_MODES:
        raise ValueError(f"`mode` must be one of {sorted(_lib.SMOOTH_MODES)}")
    if col0 < 0 or dcols < 0 or col0 + dcols > d:
        raise ValueError("column slice outside the matrix")
    with torch.cuda.device(x.device):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vvb_smooth_workspace_bytes(x.element_size(), n, d, n_iter, ctypes.byref(nbytes)))
        ws = _lib.raw_empty(int(nbytes.value), torch.uint8, x.device)
        if out is None:
            out = _lib.raw_empty(x.shape, x.dtype, x.device)
        if out.shape != x.shape or out.dtype != x.dtype or not out.is_contiguous() or out.data_ptr() == x.data_ptr():
            raise ValueError("`out` must be a contiguous tensor of x's shape and dtype, distinct from x")
        _lib.check(lib.vvb_smooth_cols_device(x.data_ptr(), x.element_size(), n, d, int(col0), int(dcols), idx.data_ptr(),
                                              idx.shape[1], int(k), float(coef), int(n_iter), _lib.SMOOTH_MODES[mode],
                                              out.data_ptr(), ws.data_ptr(), ws.numel(),
                                              torch.cuda.current_stream().cuda_stream))
        ws.record_stream(torch.cuda.current_stream())
        if validate and int(ws[:4].view(torch.int32).item()):
            raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
    return out


def smooth_rows_device(x: torch.Tensor, idx_rows: torch.Tensor, k: int, coef: float, row_lo: int) -> torch.Tensor:
    """Rows [row_lo, row_lo + len(idx_rows)) of ONE Jacobi sweep over `x` (n, d): `idx_rows` holds the neighbour
    lists of just those rows.  Returns the new rows.  (A Gauss-Seidel sweep does not split by rows.)"""
    lib = _lib.load()
    if x.dtype not in (torch.float32, torch.float64):
        raise ValueError("`x` must be float32 or float64")
    x = _require_cuda("x", x, x.dtype)
    idx_rows = _require_cuda("idx_rows", idx_rows, torch.int64)
    n, d = x.shape
    rows = idx_rows.shape[0]
    with torch.cuda.device(x.device):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vvb_smooth_workspace_bytes(x.element_size(), n, d, 1, ctypes.byref(nbytes)))
        ws = _lib.raw_empty(int(nbytes.value), torch.uint8, x.device)
        out = _lib.raw_empty((rows, d), x.dtype, x.device)
        _lib.check(lib.vvb_smooth_rows_device(x.data_ptr(), x.element_size(), n, d, idx_rows.data_ptr(),
                                              idx_rows.shape[1], int(k), float(coef), _lib.SMOOTH_MODES["jacobi"],
                                              int(row_lo), int(row_lo) + rows, out.data_ptr(), ws.data_ptr(),
                                              ws.numel(), torch.cuda.current_stream().cuda_stream))
        bad = int(ws[:4].view(torch.int32).item())  # device-side index validation (synchronises)
    if bad:
        raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
    return out


# ------------------------------------------------------------------ running top-k merge (database-sharded build)
def knn_merge_init(nq: int, k: int, self0: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    """Empty running lists for query rows [self0, self0+nq): `(run_idx int64, run_d2 float32)`, both (nq, k+1)."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("knn_merge_init: a CUDA device is required (vivae_b200 has no CPU path)")
    with torch.cuda.device(device):
        run_idx = _lib.raw_empty((nq, k + 1), torch.int64, device)
        run_d2 = _lib.raw_empty((nq, k + 1), torch.float32, device)
        _lib.check(lib.vvb_knn_merge_init_device(run_idx.data_ptr(), run_d2.data_ptr(), nq, k, self0,
                                                 torch.cuda.current_stream().cuda_stream))
    return run_idx, run_d2


def knn_merge(xc: torch.Tensor, q0: int, nq: int, new_idx: torch.Tensor, keep_lo: int, keep_hi: int, global_off: int,
              run_idx: torch.Tensor, run_d2: torch.Tensor) -> None:
    """Fold the neighbour lists `new_idx` (nq, kn+1; local rows of the stacked matrix `xc`, whose rows
    [q0, q0+nq) are the queries) into the running lists, in place (include/vivae_b200.h: vvb_knn_merge_device)."""
    lib = _lib.load()
    xc = _require_cuda("xc", xc, torch.float32)
    new_idx = _require_cuda("new_idx", new_idx, torch.int64)
    _require_cuda("run_idx", run_idx, torch.int64)
    _require_cuda("run_d2", run_d2, torch.float32)
    if not (run_idx.is_contiguous() and run_d2.is_contiguous()):
        raise ValueError("running lists must be contiguous")
    if new_idx.shape[0] != nq or run_idx.shape[0] != nq or run_d2.shape != run_idx.shape:
        raise ValueError("list shapes do not match the number of query rows")
    with torch.cuda.device(xc.device):
        _lib.check(lib.vvb_knn_merge_device(xc.data_ptr(), xc.shape[0], xc.shape[1], q0, nq, new_idx.data_ptr(),
                                            new_idx.shape[1] - 1, keep_lo, keep_hi, global_off, run_idx.data_ptr(),
                                            run_d2.data_ptr(), run_idx.shape[1] - 1,
                                            torch.cuda.current_stream().cuda_stream))


def knn_merge_finish(run_d2: torch.Tensor) -> torch.Tensor:
    """Distances (correctly rounded square roots) of the running squared distances."""
    lib = _lib.load()
    run_d2 = _require_cuda("run_d2", run_d2, torch.float32)
    out = _lib.raw_empty(run_d2.shape, run_d2.dtype, run_d2.device)
    with torch.cuda.device(run_d2.device):
        _lib.check(lib.vvb_knn_merge_finish_device(run_d2.data_ptr(), run_d2.numel(), out.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
    return out

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

__all__ = ["knn_device", "smooth_device", "smooth_rows_device", "knn_merge_init", "knn_merge", "knn_merge_finish"]


def _require_cuda(name: str, t: torch.Tensor, dtype) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"`{name}` must be a CUDA tensor (vivae_b200 has no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"`{name}` must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


def knn_device(x: torch.Tensor, k: int, q0: int = 0, nq: Optional[int] = None, algo: str = "auto",
               stats: Optional[dict] = None):
    """Exact (k+1)-NN of rows [q0, q0+nq) of `x` against all rows of `x`, on `x`'s device.

    Returns `(idx int64 (nq, k+1), dist float32 (nq, k+1))` CUDA tensors; enqueued on the
    current stream (synchronises only when `stats` is requested)."""
    lib = _lib.load()
    x = _require_cuda("x", x, torch.float32)
    n, d = x.shape
    nq = n - q0 if nq is None else nq
    if k < 1 or k > n - 1:
        raise ValueError("`k` must be between 1 and (n-1)")
    if algo not in _lib.KNN_ALGOS:
        raise ValueError(f"`algo` must be one of {sorted(_lib.KNN_ALGOS)}")
    a = _lib.KNN_ALGOS[algo]
    with torch.cuda.device(x.device):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vvb_make_knn_workspace_bytes(n, d, k, nq, a, ctypes.byref(nbytes)))
        ws = _lib.raw_empty(max(int(nbytes.value), 256), torch.uint8, x.device)
        idx = _lib.raw_empty((nq, k + 1), torch.int64, x.device)
        dist = _lib.raw_empty((nq, k + 1), torch.float32, x.device)
        st = _lib.KnnStats() if stats is not None else None
        _lib.check(lib.vvb_make_knn_device(x.data_ptr(), n, d, k, q0, nq, a, idx.data_ptr(), dist.data_ptr(),
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream,
                                           ctypes.byref(st) if st is not None else None))
        ws.record_stream(torch.cuda.current_stream())
    if stats is not None:
        stats.update(st.as_dict())
    return idx, dist


def smooth_device(x: torch.Tensor, idx: torch.Tensor, k: int = 50, coef: float = 0.1, n_iter: int = 1,
                  mode: str = "gauss_seidel") -> torch.Tensor:
    """`smooth` on device-resident data: `x` (n, d) float32/float64, `idx` (n, >=k) int64."""
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

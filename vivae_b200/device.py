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

__all__ = ["knn_device", "smooth_device"]


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
        ws = torch.empty(max(int(nbytes.value), 256), dtype=torch.uint8, device=x.device)
        idx = torch.empty((nq, k + 1), dtype=torch.int64, device=x.device)
        dist = torch.empty((nq, k + 1), dtype=torch.float32, device=x.device)
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
        ws = torch.empty(int(nbytes.value), dtype=torch.uint8, device=x.device)
        out = torch.empty_like(x)
        _lib.check(lib.vvb_smooth_device(x.data_ptr(), x.element_size(), n, d, idx.data_ptr(), idx.shape[1], int(k),
                                         float(coef), int(n_iter), _lib.SMOOTH_MODES[mode], out.data_ptr(),
                                         ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        ws.record_stream(torch.cuda.current_stream())
    return out

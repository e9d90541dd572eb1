"""k-nearest-neighbour graph builder and kNN smoother -- host side.

Drop-in for `/root/reference/src/vivae/knn.py` (`make_knn` 87-144, `smooth` 147-185,
`ensure_valid_knn` 24-63, `correct_knn_for_duplicates` 66-84): same names, argument
meaning, defaults, return layout and `ValueError`s.  The numerics run in
`libvivae_b200.so` (sm_100a CUDA) through the C ABI of `include/vivae_b200.h`;
there is no CPU implementation in this package.

Differences from the reference, all deliberate:
  * `make_knn` is EXACT (the reference's pynndescent graph is approximate and
    seeded); `random_state` is accepted and ignored.
  * additive keyword-only arguments: `make_knn(..., algo=, stats=, compact_cache=)`,
    `smooth(..., mode=)`.
  * capacity limits of the CUDA kernels surface as `NotImplementedError` (never as one of
    the reference's `ValueError`s): `make_knn` k > 1984, `smooth` k > 1024.  The tensor-core
    path covers k <= 152 and d <= 4093; beyond that `algo='auto'` uses the FP32 brute-force kernel.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import _lib

__all__ = ["make_knn", "smooth", "ensure_valid_knn", "correct_knn_for_duplicates"]

_FLOATS = (float, np.float32, np.float64)
_INTS = (int, np.int32, np.int64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def correct_knn_for_duplicates(knn: list) -> list:
    """Move the self index into column 0 of every row where a duplicate point took
    its place (reference: knn.py:66-84).  Indices only; distances are untouched and
    `knn[0]` is edited in place, exactly like the reference."""
    idx = knn[0]
    n = idx.shape[0]
    rows = np.flatnonzero(idx[:, 0] != np.arange(n, dtype=idx.dtype))
    if rows.size:
        hit = idx[rows] == rows[:, None].astype(idx.dtype)
        if not hit.any(axis=1).all():
            # the reference fails here too ("setting an array element with a sequence")
            raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
        pos = hit.argmax(axis=1)
        first = idx[rows, 0].copy()
        idx[rows, 0] = rows.astype(idx.dtype)
        idx[rows, pos] = first
    return knn


def ensure_valid_knn(n: int, knn: list) -> list:
    """Validate a k-NNG object (reference: knn.py:24-63); returns the duplicate-corrected
    object or raises the reference's `ValueError`s.  The two O(n) Python scans of the
    reference (knn.py:58, 80) are vectorised."""
    if len(knn) != 2 or not isinstance(knn[0], np.ndarray) or not isinstance(knn[1], np.ndarray):
        raise ValueError("`knn` object must be a list of 2 np.ndarray objects (index matrix and distance matrix)")
    first = knn[0][0, 0]
    if isinstance(first, _FLOATS) and float(first).is_integer():
        knn = [knn[0].astype(int), knn[1]]  # e.g. loaded from the float64 .npy cache
    if knn[0].shape[0] != n or knn[1].shape[0] != n:
        raise ValueError(
            "Neighbour and distance matrices of `knn` must have as many rows as the number of points in the reference coordinate matrix"
        )
    if not isinstance(knn[0][0, 0], _INTS):
        raise ValueError(
            "Neighbour matrix (`knn[0]`) must be an `np.ndarray` of `int`, `np.int32` or `np.int64` or coercible to integer"
        )
    if not isinstance(knn[1][0, 0], _FLOATS):
        raise ValueError("Distance matrix (`knn[1]`) must be an `np.ndarray` of `float`, `np.float32` or `np.float64`")
    self_first, dist0_zero = _scan_first_column(n, knn)
    if not self_first:
        knn = correct_knn_for_duplicates(knn)
    if not dist0_zero:
        raise ValueError("Neighbourhoods in `knn` must include self and neighbour indices must be 0-indexed")
    return knn


def _scan_first_column(n: int, knn: list):
    """(every idx[i, 0] == i, every dist[i, 0] == 0): the two O(n) scans of knn.py:58-61.  Row-major int64 /
    float matrices are scanned by the library's threaded host routine (9.5 -> ~1.5 ms at 1M rows, a tenth of an
    end-to-end smooth call); anything else by NumPy."""
    idx, dist = knn[0], knn[1]
    if (n >= 4096 and idx.dtype == np.int64 and dist.dtype in (np.float32, np.float64) and idx.ndim == 2
            and dist.ndim == 2 and idx.strides[1] == 8 and dist.strides[1] == dist.itemsize
            and idx.strides[0] > 0 and idx.strides[0] % 8 == 0
            and dist.strides[0] > 0 and dist.strides[0] % dist.itemsize == 0 and os.path.isfile(_lib.LIB_PATH)):
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.load().vvb_knn_validate_host(_ptr(idx), idx.strides[0] // 8, _ptr(dist), dist.itemsize,
                                                     dist.strides[0] // dist.itemsize, n, ctypes.byref(a),
                                                     ctypes.byref(b)))
        return bool(a.value), bool(b.value)
    return bool(np.array_equal(idx[:, 0], np.arange(n))), bool(np.all(dist[:, 0] == 0.0))


def _knn_host(x: np.ndarray, k: int, q0: int, nq: int, algo: str, stats: Optional[dict]):
    """One call of `vvb_make_knn_host` (HOST buffers in and out)."""
    lib = _lib.load()
    n, d = x.shape
    idx = _lib.result_empty((nq, k + 1), np.int64)  # page-locked when large: the D2H copy is then one DMA
    dist = _lib.result_empty((nq, k + 1), np.float32)
    st = _lib.KnnStats()
    if algo not in _lib.KNN_ALGOS:
        raise ValueError(f"`algo` must be one of {sorted(_lib.KNN_ALGOS)}")
    _lib.check(lib.vvb_make_knn_host(_ptr(x), n, d, k, q0, nq, _lib.KNN_ALGOS[algo], _ptr(idx), _ptr(dist),
                                     ctypes.byref(st)))
    if stats is not None:
        stats.update(st.as_dict())
    return idx, dist


def _sidecar_path(fname: str) -> str:
    return fname + ".vvb.npz"


def _save_compact_sidecar(fname: str, idx: np.ndarray, dist: np.ndarray) -> None:
    """Compact copy of the cache next to the reference-format file: int32 indices + float32 distances
    (8 bytes per entry instead of 16; at 10M x 101 that is 8 GB instead of 16 GB)."""
    if idx.shape[0] > np.iinfo(np.int32).max:
        return
    with open(_sidecar_path(fname), "wb") as f:
        np.savez(f, idx=idx.astype(np.int32), dist=np.asarray(dist, dtype=np.float32),
                 shape=np.asarray(idx.shape, dtype=np.int64))


def _load_compact_sidecar(fname: str):
    """`[idx int64, dist float32]` from the sidecar, or None when it is missing, older than the reference-format
    file, or does not describe the same graph (then the caller reads the reference-format file)."""
    side = _sidecar_path(fname)
    try:
        if not os.path.isfile(side) or os.path.getmtime(side) < os.path.getmtime(fname):
            return None
        with np.load(side) as z:
            idx, dist, shape = z["idx"], z["dist"], tuple(int(v) for v in z["shape"])
        if idx.shape != shape or dist.shape != shape or idx.dtype != np.int32 or dist.dtype != np.float32:
            return None
        return [idx.astype(np.int64), dist]
    except Exception:
        return None


def make_knn(
    x: Optional[np.ndarray] = None,
    fname: Optional[str] = None,
    k: int = 100,
    as_tuple: bool = False,
    random_state: Optional[int] = None,
    verbose: bool = True,
    *,
    algo: str = "auto",
    stats: Optional[dict] = None,
    compact_cache: bool = True,
) -> list:
    """Construct or load a k-nearest-neighbour graph (reference: knn.py:87-144).

    Returns `[idx, dist]` (or a tuple): `idx` int64 (n, k+1), `dist` float32 (n, k+1),
    column 0 is the point itself at distance 0, the rest ascending true Euclidean
    distances.  The graph is exact: neighbours ranked by the FP32 key (d2, index).

    Args:
        x: row-wise coordinates.
        fname: cache file; loaded if it exists (knn.py:122-131), else written (knn.py:142-143)
            in the reference's format (`np.save(fname, [idx, dist])`: one float64 (2, n, k+1) array).
        k: neighbour count (self excluded).
        as_tuple: return a tuple instead of a list.
        random_state: accepted for compatibility; the exact builder has no randomness.
        verbose: print the reference's progress messages.
        algo: 'auto' (tensor-core screening + FP32 re-rank with certified fallback),
            'brute' (FP32 brute force) or 'tensor'.
        stats: optional dict that receives the `vvb_knn_stats` counters.
        compact_cache: with `fname`, also keep a compact sidecar `<fname>.vvb.npz` (int32 indices, float32
            distances: half the bytes of the reference's float64 pair and no float->int pass on load).  The
            reference-format file is always written and stays the source of truth: the sidecar is used on load
            only when it is at least as new as `fname` and describes the same (n, k+1) graph.

    Raises `NotImplementedError` for k > 1984 (capacity of the brute-force kernel's candidate lists).
    """
    if k < 1 or k > (x.shape[0] - 1):
        raise ValueError("`k` must be between 1 and (n-1)")

    if fname is not None and os.path.isfile(fname):
        if verbose:
            print("Loading k-NNG")
        knn = _load_compact_sidecar(fname) if compact_cache else None
        if knn is None:
            res = np.load(fname, allow_pickle=True)
            knn = [np.array(res[0], dtype=np.int64), np.array(res[1])]
    else:
        if verbose:
            print("Constructing k-NNG")
        xf = np.ascontiguousarray(x, dtype=np.float32)  # pynndescent coerces to float32 as well
        if xf.ndim != 2:
            raise ValueError("`x` must be a 2-dimensional array")
        idx, dist = _knn_host(xf, int(k), 0, xf.shape[0], algo, stats)
        knn = [idx, dist]  # self is pinned to column 0 on the device: no duplicate fix-up needed
        if fname is not None:
            np.save(fname, knn)  # the reference's format (knn.py:142-143): one float64 (2, n, k+1) array
            if compact_cache:
                _save_compact_sidecar(fname, idx, dist)
    if as_tuple:
        knn = tuple(knn)
    return knn


def smooth(
    x: np.ndarray,
    knn: list,
    k: int = 50,
    coef: Optional[float] = 0.1,
    n_iter: int = 1,
    *,
    mode: str = "gauss_seidel",
) -> np.ndarray:
    """Denoise by moving each point towards the mean of its `k` nearest neighbours
    (reference: knn.py:147-185).

    `mode='gauss_seidel'` reproduces the reference bit for bit: its loop updates the
    working array in place, so row i sees rows j < i already updated in the same
    sweep.  `mode='jacobi'` (every row reads the previous sweep) is a faster,
    order-independent variant that does NOT match the reference.

    Raises `NotImplementedError` for k > 1024 (a warp keeps a row's neighbour list in registers).
    """
    if k < 1 or k > (x.shape[0] - 1):
        raise ValueError("`k` must be between 1 and (n-1)")
    if coef < 0.0 or coef > 1.0:
        raise ValueError("`coef` must be more than 0. and at most 1.")
    if n_iter < 1 or n_iter > 10000:
        raise ValueError("`n_iter` must be at least 1 and at most 10000")
    knn = ensure_valid_knn(n=x.shape[0], knn=knn)
    if mode not in _lib.SMOOTH_MODES:
        raise ValueError(f"`mode` must be one of {sorted(_lib.SMOOTH_MODES)}")
    if knn[0].shape[1] < k:
        raise ValueError("`knn` has fewer than `k` neighbour columns")

    xin = x if x.dtype in (np.float32, np.float64) else x.astype(np.float64)
    xin = np.ascontiguousarray(xin)
    idx = np.ascontiguousarray(knn[0], dtype=np.int64)  # index range is validated on the device
    lib = _lib.load()
    out = _lib.result_empty(xin.shape, xin.dtype)
    _lib.check(lib.vvb_smooth_host(_ptr(xin), xin.dtype.itemsize, xin.shape[0], xin.shape[1], _ptr(idx),
                                   idx.shape[1], int(k), float(coef), int(n_iter), _lib.SMOOTH_MODES[mode], _ptr(out)))
    return out

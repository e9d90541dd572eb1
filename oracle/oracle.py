"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

* ctypes bindings to `oracle/liboracle.so` (built by `oracle/Makefile` from
  `oracle/vivae_oracle.c`; see that file's header for the parity status of each
  function and the reference file:line each one restates), and
* small pure-NumPy / pure-Python restatements used to cross-check the C code on
  tiny inputs.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this module.  The product package
`vivae_b200` never does.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc only; no GPU, no reference needed)."""
    src = os.path.join(_HERE, "vivae_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        c_i64, c_int, c_dbl, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p
        L.vv_oracle_version.restype = c_int
        L.vv_oracle_num_threads.restype = c_int
        L.vv_oracle_knn_f32.restype = c_int
        L.vv_oracle_knn_f32.argtypes = [vp, c_i64, c_int, c_int, c_i64, c_i64, vp, vp]
        for nm in ("vv_oracle_smooth_f32", "vv_oracle_smooth_f64"):
            f = getattr(L, nm)
            f.restype = c_int
            f.argtypes = [vp, c_i64, c_int, vp, c_i64, c_int, c_dbl, c_int, c_int, vp]
        L.vv_oracle_quartet_pass_f64.restype = c_dbl
        L.vv_oracle_quartet_pass_f64.argtypes = [vp, vp, c_i64, c_int, c_int, vp, c_int, c_int, c_dbl, vp]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().vv_oracle_num_threads())


def set_threads(n: int) -> None:
    """OpenMP threads of the oracle's kNN (torchrun sets OMP_NUM_THREADS=1 in its workers)."""
    L = lib()
    L.vv_oracle_set_threads.restype = None
    L.vv_oracle_set_threads.argtypes = [ctypes.c_int]
    L.vv_oracle_set_threads(int(n))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- kNN
def knn(x: np.ndarray, k: int, q0: int = 0, q1: int | None = None):
    """Exact (k+1)-NN incl. self of rows [q0,q1) against all rows (C, OpenMP).

    Layout follows /root/reference/src/vivae/knn.py:135-141: idx int64 (m,k+1),
    dist float32 (m,k+1), column 0 = self / 0.0.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    q1 = n if q1 is None else q1
    idx = np.empty((q1 - q0, k + 1), dtype=np.int64)
    dist = np.empty((q1 - q0, k + 1), dtype=np.float32)
    rc = lib().vv_oracle_knn_f32(_ptr(x), n, d, k, q0, q1, _ptr(idx), _ptr(dist))
    if rc != 0:
        raise ValueError(f"vv_oracle_knn_f32 failed with code {rc}")
    return idx, dist


def _round_to_f32(v: Fraction) -> np.float32:
    """Round an exact rational to the nearest-even float32 (no double rounding)."""
    if v == 0:
        return np.float32(0.0)
    sign = -1 if v < 0 else 1
    a = abs(v)
    # exponent e with 2^23 <= a / 2^e < 2^24, clamped at the subnormal exponent -149
    e = a.numerator.bit_length() - a.denominator.bit_length() - 24
    while a / Fraction(2) ** e >= 2**24:
        e += 1
    while a / Fraction(2) ** e < 2**23:
        e -= 1
    e = max(e, -149)
    q = a / Fraction(2) ** e
    fl = q.numerator // q.denominator
    rem = q - fl
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (fl & 1)):
        fl += 1
    return np.float32(sign * float(fl) * 2.0**e)  # fl < 2^24+1 and the product are exact in float64


def knn_tiny(x: np.ndarray, k: int):
    """Pure-Python exact restatement of the kNN key for tiny inputs.

    d2 = fold over c of fmaf(t, t, acc), t = fl32(x_ic - x_jc), evaluated with
    exact rational arithmetic and one correctly rounded float32 rounding per
    fmaf -- independent of libm / compiler behaviour.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    idx = np.empty((n, k + 1), dtype=np.int64)
    dist = np.empty((n, k + 1), dtype=np.float32)
    for i in range(n):
        keys = []
        for j in range(n):
            if j == i:
                continue
            acc = np.float32(0.0)
            for c in range(d):
                t = np.float32(x[i, c] - x[j, c])
                tf = Fraction(float(t))
                acc = _round_to_f32(tf * tf + Fraction(float(acc)))
            keys.append((float(acc), j))
        keys.sort()
        idx[i, 0], dist[i, 0] = i, 0.0
        for m in range(k):
            idx[i, m + 1] = keys[m][1]
            dist[i, m + 1] = np.sqrt(np.float32(keys[m][0]))
    return idx, dist


# ------------------------------------------------------------------------ smooth
def smooth(x: np.ndarray, idx: np.ndarray, k: int, coef: float, n_iter: int = 1, mode: int = 0):
    """C restatement of /root/reference/src/vivae/knn.py:173-185 (mode 0 = in-place
    Gauss-Seidel sweep = the reference; mode 1 = Jacobi, the non-parity variant)."""
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    x = np.ascontiguousarray(x)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n, d = x.shape
    out = np.empty_like(x)
    fn = lib().vv_oracle_smooth_f32 if x.dtype == np.float32 else lib().vv_oracle_smooth_f64
    rc = fn(_ptr(x), n, d, _ptr(idx), idx.shape[1], k, float(coef), n_iter, mode, _ptr(out))
    if rc != 0:
        raise ValueError(f"vv_oracle_smooth failed with code {rc}")
    return out


def smooth_numpy(x: np.ndarray, idx: np.ndarray, k: int, coef: float, n_iter: int = 1, mode: int = 0):
    """NumPy restatement (row loop in Python; small inputs only)."""
    y = np.array(x, copy=True)
    c = y.dtype.type(coef)
    for _ in range(n_iter):
        src = y.copy() if mode == 1 else y
        for i in range(y.shape[0]):
            acc = src[idx[i, 0]].copy()
            for jj in range(1, k):
                acc = acc + src[idx[i, jj]]
            target = acc / y.dtype.type(k)
            y[i] = src[i] + (target - src[i]) * c
    return y


# ----------------------------------------------------------------------- quartet
METRICS = {"euclidean": 0, "cosine": 1}


def quartet_loss(x, z, distf_hd="euclidean", distf_ld="euclidean", perms=None, want_grad=True):
    """Loss and d(loss)/dz of /root/reference/src/vivae/losses.py:266-349 in float64.

    `perms`: list with one entry per sampling pass; entry None = identity
    (losses.py:331), otherwise an int64 permutation of range(B) (losses.py:329).
    """
    x = np.ascontiguousarray(x, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.float64)
    B, D = x.shape
    L = z.shape[1]
    perms = [None] if perms is None else list(perms)
    ns = len(perms)
    nq = B // 4
    grad = np.zeros((B, L), dtype=np.float64) if want_grad else None
    total = 0.0
    for p in perms:
        pp = None if p is None else np.ascontiguousarray(p, dtype=np.int64)
        scale = (B / (nq * nq) / ns) if nq else 0.0
        total += lib().vv_oracle_quartet_pass_f64(
            _ptr(x), _ptr(z), B, D, L, None if pp is None else _ptr(pp),
            METRICS[distf_hd], METRICS[distf_ld], scale, None if grad is None else _ptr(grad))
    return total / ns, grad


def quartet_loss_numpy(x, z, distf_hd="euclidean", distf_ld="euclidean", perms=None):
    """Vectorised float64 NumPy restatement of the loss value (no gradient)."""
    x = np.asarray(x, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64)
    B = x.shape[0]
    nq = B // 4
    perms = [None] if perms is None else list(perms)

    def dist(a, b, metric):
        if metric == "euclidean":
            return np.sqrt(np.maximum(((a - b) ** 2).sum(1), 1e-9))
        na = np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-8)
        nb = np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-8)
        return 1.0 - ((a / na) * (b / nb)).sum(1)

    pairs = [(0, 1), (1, 2), (2, 3), (0, 2), (0, 3), (1, 3)]
    total = 0.0
    for p in perms:
        p = np.arange(B) if p is None else np.asarray(p)
        xq = [x[p][a * nq:(a + 1) * nq] for a in range(4)]
        zq = [z[p][a * nq:(a + 1) * nq] for a in range(4)]
        dx = [dist(xq[a], xq[b], distf_hd) for a, b in pairs]
        dz = [dist(zq[a], zq[b], distf_ld) for a, b in pairs]
        X, S = sum(dx), sum(dz)
        cost = sum((dx[i] / X - dz[i] / S) ** 2 for i in range(6))
        with np.errstate(invalid="ignore", divide="ignore"):
            total += (cost.mean() if nq else np.nan) * B / nq if nq else np.nan
    return total / len(perms)

"""ctypes binding of the C ABI in include/vivae_b200.h (libvivae_b200.so).

There is no CPU fallback: if the shared library is missing, or a compute entry
point is called without a CUDA device, the call raises.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# VVB_LIB_PATH: a differently configured build of the same library (kernel experiments: profiles/)
LIB_PATH = os.environ.get("VVB_LIB_PATH") or os.path.join(_HERE, "libvivae_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, EINVAL, ECUDA, ENOMEM, ENOTSUP = 0, -1, -2, -3, -4
KNN_ALGOS = {"auto": 0, "brute": 1, "tensor": 2}
SMOOTH_MODES = {"gauss_seidel": 0, "jacobi": 1}
DIST_FUNCS = {"euclidean": 0, "cosine": 1}


class KnnStats(ctypes.Structure):
    """Mirror of `vvb_knn_stats`."""

    _fields_ = [
        ("rows", ctypes.c_int64),
        ("rows_fallback", ctypes.c_int64),
        ("algo_used", ctypes.c_int32),
        ("candidates", ctypes.c_int32),
        ("ms_prepare", ctypes.c_float),
        ("ms_screen", ctypes.c_float),
        ("ms_rerank", ctypes.c_float),
        ("ms_fallback", ctypes.c_float),
        ("kernel_launches", ctypes.c_int64),
        ("tiles_swept", ctypes.c_int64),
        ("tiles_dense", ctypes.c_int64),
        ("cells", ctypes.c_int32),
        ("rows_refined", ctypes.c_int32),
        ("products", ctypes.c_int32),
        ("tiles_pre", ctypes.c_int32),
    ]

    def as_dict(self) -> dict:
        return {name: getattr(self, name) for name, _ in self._fields_}


class Exchange(ctypes.Structure):
    """Mirror of `vvb_exchange`."""

    _fields_ = [("ptr", ctypes.c_void_p), ("count", ctypes.c_size_t), ("dtype", ctypes.c_int32), ("op", ctypes.c_int32)]


XCH_I32, XCH_U32, XCH_F64 = 0, 1, 2
XCH_GATHER_ROWS, XCH_SUM, XCH_MAX_ORD = 0, 1, 2

# every symbol include/vivae_b200.h declares: (restype, argtypes)
_i64, _int, _dbl, _vp, _sz = ctypes.c_int64, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t
SYMBOLS = {
    "vvb_version": (_int, []),
    "vvb_last_error": (ctypes.c_char_p, []),
    "vvb_launch_count": (_i64, []),
    "vvb_release_cached_memory": (_int, []),
    "vvb_knn_validate_host": (_int, [_vp, _i64, _vp, _int, _i64, _i64, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "vvb_host_alloc": (_int, [_sz, ctypes.POINTER(ctypes.c_void_p)]),
    "vvb_host_free": (_int, [_vp]),
    "vvb_make_knn_host": (_int, [_vp, _i64, _int, _int, _i64, _i64, _int, _vp, _vp, ctypes.POINTER(KnnStats)]),
    "vvb_make_knn_workspace_bytes": (_int, [_i64, _int, _int, _i64, _int, ctypes.POINTER(_sz)]),
    "vvb_make_knn_device": (_int, [_vp, _i64, _int, _int, _i64, _i64, _int, _vp, _vp, _vp, _sz, _vp,
                                   ctypes.POINTER(KnnStats)]),
    "vvb_knn_plan_phase_device": (_int, [_vp, _i64, _int, _int, _i64, _vp, _sz, _int, _int, _int, _vp, _vp,
                                         ctypes.POINTER(_int)]),
    "vvb_make_knn_planned_device": (_int, [_vp, _i64, _int, _int, _i64, _i64, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vvb_knn_screen_debug_capacity": (_int, []),
    "vvb_knn_screen_debug_host": (_int, [_vp, _i64, _int, _int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vvb_knn_merge_init_device": (_int, [_vp, _vp, _i64, _int, _i64, _vp]),
    "vvb_knn_merge_device": (_int, [_vp, _i64, _int, _i64, _i64, _vp, _int, _i64, _i64, _i64, _vp, _vp, _int, _vp]),
    "vvb_knn_merge_finish_device": (_int, [_vp, _i64, _vp, _vp]),
    "vvb_smooth_host": (_int, [_vp, _int, _i64, _int, _vp, _i64, _int, _dbl, _int, _int, _vp]),
    "vvb_smooth_workspace_bytes": (_int, [_int, _i64, _int, _int, ctypes.POINTER(_sz)]),
    "vvb_smooth_device": (_int, [_vp, _int, _i64, _int, _vp, _i64, _int, _dbl, _int, _int, _vp, _vp, _sz, _vp]),
    "vvb_smooth_cols_device": (_int, [_vp, _int, _i64, _int, _int, _int, _vp, _i64, _int, _dbl, _int, _int, _vp, _vp, _sz, _vp]),
    "vvb_peer_alloc": (_int, [_sz, ctypes.POINTER(ctypes.c_void_p), _vp]),
    "vvb_peer_open": (_int, [_vp, ctypes.POINTER(ctypes.c_void_p)]),
    "vvb_peer_close": (_int, [_vp]),
    "vvb_peer_free": (_int, [_vp]),
    "vvb_smooth_peer_workspace_bytes": (_sz, []),
    "vvb_smooth_peer_device": (_int, [_vp, _int, _i64, _int, _vp, _i64, _int, _dbl, _int, _int, _int, _vp, _vp, _vp]),
    "vvb_smooth_peer_unpack_device": (_int, [_vp, _int, _i64, _int, _vp, _vp, _vp]),
    "vvb_smooth_rows_device": (_int, [_vp, _int, _i64, _int, _vp, _i64, _int, _dbl, _int, _i64, _i64, _vp, _vp, _sz, _vp]),
    "vvb_vae_step_device": (_int, [_vp, _vp, _vp, _vp, _int, _int, _vp, _vp, _int, _int, _int, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_sz),
                                   ctypes.POINTER(_int), ctypes.POINTER(_int), _vp]),
    "vvb_quartet_workspace_bytes": (_sz, [_i64]),
    "vvb_quartet_loss_fwd_device": (_int, [_vp, _vp, _i64, _int, _int, _vp, _int, _int, _int, _vp, _vp, _vp, _vp]),
    "vvb_quartet_loss_bwd_device": (_int, [_vp, _vp, _i64, _vp, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libvivae_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libvivae_b200.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(vivae_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the header disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def last_error() -> str:
    return load().vvb_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Translate a VVB_E* status into the reference's exception conventions."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == ENOTSUP:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


_graph_replayed = 0  # kernels of this library launched through CUDA-graph replays


def note_graph_replay(kernels: int) -> None:
    global _graph_replayed
    _graph_replayed += int(kernels)


def launch_count() -> int:
    """Kernels of this library launched so far: direct launches (counted in C) plus the library's
    kernel nodes inside replayed CUDA graphs (counted at capture time, added per replay)."""
    return int(load().vvb_launch_count()) + _graph_replayed


class no_uninitialised_fill:
    """Context: suspend the NaN fill that deterministic mode adds to every `torch.empty` (11-13 extra kernels in
    each training step, inside autograd's own allocations).  A debugging aid of torch, not part of determinism:
    results are unchanged."""

    def __enter__(self):
        import torch

        self.det = torch.utils.deterministic
        self.prev = self.det.fill_uninitialized_memory
        self.det.fill_uninitialized_memory = False
        return self

    def __exit__(self, *exc):
        self.det.fill_uninitialized_memory = self.prev
        return False


def raw_empty(shape, dtype, device):
    """`torch.empty` without the fill kernel that `torch.use_deterministic_algorithms(True)` (the package
    default, like the reference's) adds to every uninitialised allocation: on a 1.6 GB kNN workspace that fill
    costs 0.25 ms per call, and inside the captured training step it is four extra kernels per batch.  Only for
    buffers the kernels fully write before anything reads them."""
    import torch

    det = torch.utils.deterministic
    if not (torch.are_deterministic_algorithms_enabled() and det.fill_uninitialized_memory):
        return torch.empty(shape, dtype=dtype, device=device)
    det.fill_uninitialized_memory = False
    try:
        return torch.empty(shape, dtype=dtype, device=device)
    finally:
        det.fill_uninitialized_memory = True


# ---------------------------------------------------------------------------------------------------------
# Result arrays in page-locked memory.  A D2H copy into a fresh pageable NumPy array is bound by the host
# (page faults, then a staging memcpy: ~20 GB/s on the bench host); into page-locked memory it is one DMA at
# PCIe speed (54 GB/s).  Pinning is slow (~0.5 s per GB), so blocks are recycled: when the last array that
# views a block is garbage-collected the block goes back to the pool, and the next call of the same size reuses
# it.  The first request of a size gets ordinary memory (a one-off call never pays for pinning); repeated
# calls of the same shape reach the steady state after two or three calls.  `VVB_PINNED_RESULTS=0` switches this off; `VVB_PINNED_MAX_MB` (default 4096) caps the pool -- beyond it
# results are ordinary NumPy arrays.
class _PinnedBlock:
    """Owner of one page-locked block; NumPy arrays made from it keep it alive through `.base`."""

    def __init__(self, ptr: int, cap: int, nbytes: int):
        self.ptr, self.cap, self.nbytes = ptr, cap, nbytes

    @property
    def __array_interface__(self):
        return {"data": (self.ptr, False), "shape": (self.nbytes,), "typestr": "|u1", "version": 3}

    def __del__(self):
        try:
            _pinned_pool.give_back(self.ptr, self.cap)
        except Exception:  # interpreter shutdown
            pass


class _PinnedPool:
    def __init__(self):
        import threading

        self.lock = threading.RLock()  # re-entrant: a garbage collection inside take() may run a block's __del__
        self.free = []   # (cap, ptr)
        self.total = 0   # bytes pinned by this pool (in use + free)
        self.seen = {}   # size class -> requests that found no free block

    def _limit(self) -> int:
        return int(os.environ.get("VVB_PINNED_MAX_MB", "4096")) << 20

    def take(self, nbytes: int):
        """A free block of this size class, or -- from the second request of the class on -- a newly pinned
        one; None when the caller should use ordinary memory.  The very first request of a size is never
        pinned: a one-off call would pay the pinning and never get it back."""
        lib = load()
        cap = (nbytes + (1 << 21) - 1) >> 21 << 21
        with self.lock:
            for c, p in self.free:
                if c == cap:
                    self.free.remove((c, p))
                    return _PinnedBlock(p, cap, nbytes)
            self.seen[cap] = self.seen.get(cap, 0) + 1
            if self.seen[cap] < 2:
                return None
            while self.total + cap > self._limit() and self.free:  # make room: drop idle blocks, largest first
                c, p = max(self.free)
                self.free.remove((c, p))
                lib.vvb_host_free(p)
                self.total -= c
            if self.total + cap > self._limit():
                return None
            ptr = ctypes.c_void_p()
            if lib.vvb_host_alloc(cap, ctypes.byref(ptr)) != OK or not ptr.value:
                return None
            self.total += cap
            return _PinnedBlock(ptr.value, cap, nbytes)

    def give_back(self, ptr: int, cap: int) -> None:
        with self.lock:
            self.free.append((cap, ptr))

    def release(self) -> None:
        lib = load()
        with self.lock:
            for c, p in self.free:
                lib.vvb_host_free(p)
                self.total -= c
            self.free = []


_pinned_pool = _PinnedPool()


def result_empty(shape, dtype):
    """Uninitialised result array: page-locked (pool above) when large and allowed, else `np.empty`."""
    import numpy as np

    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes >= (8 << 20) and os.environ.get("VVB_PINNED_RESULTS", "1") != "0":
        block = _pinned_pool.take(nbytes)
        if block is not None:
            return np.asarray(block).view(dtype).reshape(shape)
    return np.empty(shape, dtype=dtype)

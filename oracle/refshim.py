"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference ViVAE.

Only `oracle/make_golden.py` (run in the build container, where
`/root/reference` is mounted) uses this module.  It is never imported by the
product package, by the `-m gpu` tests, by `smoke()` or by `bench.py`
(`/root/reference` does not exist on the GPU box).

What it does (recipe: SURVEY.md Appendix B):
  1. stubs `pynndescent` (only the cache-miss branch of `make_knn` touches it,
     /root/reference/src/vivae/knn.py:135),
  2. stubs `matplotlib` sub-modules needed at import time by
     /root/reference/src/vivae/geometry.py:35-36 and plotting.py:35-40,
  3. makes `importlib.metadata.version("vivae")` return "0.0.2"
     (/root/reference/src/vivae/__init__.py:53, pyproject.toml:10),
  4. forces the CPU backend (`VIVAE_CUDA=0`, __init__.py:14),
  5. puts /root/reference/src on sys.path and imports `vivae`.
"""

from __future__ import annotations

import importlib
import importlib.metadata
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def load_reference():
    """Return the reference `vivae` module (CPU backend)."""
    if not os.path.isdir(REFERENCE_SRC):
        raise RuntimeError(f"reference not mounted at {REFERENCE_SRC}")
    os.environ["VIVAE_CUDA"] = "0"
    os.environ.setdefault("VIVAE_DETERMINISTIC", "1")

    try:
        import pynndescent  # noqa: F401
    except Exception:
        sys.modules["pynndescent"] = types.ModuleType("pynndescent")

    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = types.ModuleType("matplotlib")
        for sub in ("cm", "pyplot", "collections", "patches", "colors", "lines", "figure", "axes"):
            m = types.ModuleType(f"matplotlib.{sub}")
            setattr(mpl, sub, m)
            sys.modules[f"matplotlib.{sub}"] = m

        class _Stub:  # placeholder for PatchCollection / Polygon / ...
            def __init__(self, *a, **k):
                raise RuntimeError("matplotlib is stubbed in the oracle shim")

        for sub, names in {
            "collections": ("PatchCollection", "LineCollection"),
            "patches": ("Polygon", "Circle", "Patch"),
            "lines": ("Line2D",),
            "figure": ("Figure",),
            "axes": ("Axes",),
            "colors": ("Normalize", "ListedColormap"),
        }.items():
            for nm in names:
                setattr(sys.modules[f"matplotlib.{sub}"], nm, _Stub)
        sys.modules["matplotlib"] = mpl

    _real_version = importlib.metadata.version

    def _version(name):
        if name == "vivae":
            return "0.0.2"
        return _real_version(name)

    importlib.metadata.version = _version
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    try:
        return importlib.import_module("vivae")
    finally:
        importlib.metadata.version = _real_version

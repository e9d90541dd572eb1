"""Steady-state host-path times of vv.make_knn / vv.smooth at 1M x 50 for one VVB_COPY_THREADS setting
(run once per setting: the value is read when the library is first used)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import synth_cells
import vivae_b200 as vv
x = synth_cells(1_000_000, 50)
ts = []
def step():
    t0 = time.perf_counter(); knn = vv.make_knn(x, k=50, verbose=False); t1 = time.perf_counter()
    out = vv.smooth(x, knn, k=50, coef=1.0, n_iter=1); t2 = time.perf_counter()
    return 1e3 * (t1 - t0), 1e3 * (t2 - t1)
for _ in range(4): step()
for _ in range(5): ts.append(step())
print(json.dumps({"threads": os.environ.get("VVB_COPY_THREADS", "6"), "chunk_kb": os.environ.get("VVB_COPY_CHUNK_KB", "4096"), "make_knn_ms": round(float(np.median([a for a, _ in ts])), 2),
                  "smooth_ms": round(float(np.median([b for _, b in ts])), 2)}))

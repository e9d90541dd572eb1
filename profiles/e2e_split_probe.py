"""Host-path split of the end-to-end step: seconds of vv.make_knn and vv.smooth (host NumPy in/out) at 1M x 50."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import synth_cells
import vivae_b200 as vv
x = synth_cells(1_000_000, 50)
res = {}
for rep in range(6):
    t0 = time.perf_counter(); knn = vv.make_knn(x, k=50, verbose=False); t1 = time.perf_counter()
    out = vv.smooth(x, knn, k=50, coef=1.0, n_iter=1); t2 = time.perf_counter()
    res[f"rep{rep}"] = {"make_knn_ms": 1e3 * (t1 - t0), "smooth_ms": 1e3 * (t2 - t1)}
os.environ["VVB_SMOOTH_PIPELINE"] = "1"
t1 = time.perf_counter(); out2 = vv.smooth(x, knn, k=50, coef=1.0, n_iter=1); t2 = time.perf_counter()
res["smooth_pipelined_ms"] = 1e3 * (t2 - t1)
res["equal"] = bool(np.array_equal(out, out2))
print(json.dumps(res))

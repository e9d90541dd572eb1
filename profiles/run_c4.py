"""Config C4 of BASELINE.json / SURVEY.md 8(d), end to end through the public API on one GPU:
synthetic bone-marrow-shaped 250k cells x 50 PCs (12 clusters + 3 linear trajectories between centres),
make_knn(k=100) -> smooth(k=50) -> ViVAE.fit (quartet loss, batch 512) -> transform ->
encoder_indicatrices(radius=1e-3, n_steps=30, n_polygon=200).  Prints one JSON line of stage times.

    python profiles/run_c4.py [--cells 250000] [--epochs 100]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vivae_b200 as vv


def c4_cells(n, d=50, seed=3):
    rng = np.random.default_rng(seed)
    spec = (5.0 / np.sqrt(1.0 + np.arange(d))).astype(np.float32)
    centres = rng.standard_normal((12, d)).astype(np.float32) * spec
    lab = rng.integers(0, 15, size=n)
    x = np.empty((n, d), dtype=np.float32)
    noise = 0.3 * spec * rng.standard_normal((n, d)).astype(np.float32)
    for c in range(12):
        x[lab == c] = centres[c]
    for t, (a, b) in enumerate(((0, 1), (1, 2), (3, 4))):  # trajectories: points along the segment between two centres
        m = lab == 12 + t
        u = rng.random(m.sum()).astype(np.float32)[:, None]
        x[m] = (1 - u) * centres[a] + u * centres[b]
    return x + noise


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=250_000)
    ap.add_argument("--epochs", type=int, default=100)
    a = ap.parse_args()
    x = c4_cells(a.cells)
    vv.make_knn(x[:40_000], k=100, verbose=False)  # CUDA context, buffers
    res = {"cells": a.cells, "epochs": a.epochs}

    def timed(name, f):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        res[name + "_s"] = time.perf_counter() - t
        return out

    st = {}
    knn = timed("make_knn_k100", lambda: vv.make_knn(x, k=100, verbose=False, stats=st))
    xs = timed("smooth_k50", lambda: vv.smooth(x, knn, k=50, coef=1.0, n_iter=1))
    model = vv.ViVAE(input_dim=50, latent_dim=2, random_state=1)
    timed("fit", lambda: model.fit(xs, n_epochs=a.epochs, batch_size=512, lam_mds=100.0, verbose=False))
    res["fit_s_per_epoch"] = res["fit_s"] / a.epochs
    emb = timed("transform", lambda: model.transform(xs))
    ei = timed("encoder_indicatrices", lambda: model.encoder_indicatrices(xs, radius=1e-3, n_steps=30, n_polygon=200))
    res["indicatrices"] = len(ei.zp)
    res["knn_rows_fallback"] = st["rows_fallback"]
    res["embedding_finite"] = bool(np.isfinite(emb).all())
    # the same run with both geometric terms (closed-form Jacobians, still one CUDA graph per step)
    model2 = vv.ViVAE(input_dim=50, latent_dim=2, random_state=1)
    e2 = max(1, a.epochs // 10)
    timed("fit_geom", lambda: model2.fit(xs, n_epochs=e2, batch_size=512, lam_mds=100.0, lam_geom=1.0, lam_egeom=1.0, verbose=False))
    res["fit_geom_s_per_epoch"] = res["fit_geom_s"] / e2
    res["graph_used"] = model2._graph is not None and model2._graph[1] is not None
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""CPU study (NumPy, no GPU) of the filter-and-refine plan for wide rows:
how many (256-query block x 128-row database tile) pairs survive a contractive 128-dimensional projection
filter on C5-shaped data (log1p Poisson counts, 40 clusters, d = 2000), at a size NumPy handles in seconds."""
import json
import numpy as np

rng = np.random.default_rng(4)
n, d, k, nc, r = 20_000, 2000, 50, 40, 128
centres = rng.gamma(0.35, 1.0, size=(nc, d)) * 1.5
lab = rng.integers(0, nc, n)
depth = np.exp(0.3 * rng.standard_normal((n, 1)))
x = np.log1p(rng.poisson(centres[lab] * depth)).astype(np.float32)
order = np.argsort(lab, kind="stable")          # stands in for the cell-sorted order
x, lab = x[order], lab[order]
xc = x - x.mean(0)
sq = (xc * xc).sum(1)
d2 = sq[:, None] + sq[None, :] - 2.0 * (xc @ xc.T)
np.fill_diagonal(d2, np.inf)
kth = np.sqrt(np.maximum(np.partition(d2, k - 1, axis=1)[:, k - 1], 0))          # true k-th neighbour distance
same = lab[:, None] == lab[None, :]
res = {"n": n, "d": d, "k": k, "kth_distance_median": float(np.median(kth)),
       "same_cluster_distance_median": float(np.sqrt(np.median(d2[same & np.isfinite(d2)]))),
       "other_cluster_distance_median": float(np.sqrt(np.median(d2[~same])))}
for name, basis in (("random orthonormal", None), ("span of cluster means", "means")):
    if basis is None:
        q, _ = np.linalg.qr(rng.standard_normal((d, r)))
    else:
        m = np.stack([xc[lab == c].mean(0) for c in range(nc)])
        u, s, vt = np.linalg.svd(m, full_matrices=False)
        extra, _ = np.linalg.qr(rng.standard_normal((d, r)))
        q, _ = np.linalg.qr(np.concatenate([vt[: nc - 1].T, extra], axis=1)[:, :r])
    z = (xc @ q).astype(np.float32)
    zs = (z * z).sum(1)
    z2 = np.maximum(zs[:, None] + zs[None, :] - 2.0 * (z @ z.T), 0)
    assert np.all(z2 <= d2 * (1 + 1e-4) + 1e-3)                                  # contraction: a valid lower bound
    bound = (1.05 * kth) ** 2                                                    # pass (i) gives an upper bound: 5 % slack
    hit = z2 <= bound[:, None]
    qb, tb = (n + 255) // 256, (n + 127) // 128
    surv = 0
    for a in range(qb):
        rows = hit[a * 256:(a + 1) * 256]
        surv += int(np.add.reduceat(rows.any(0), np.arange(0, n, 128)).astype(bool).sum())
    res[name] = {"tiles_surviving": surv, "tiles_total": qb * tb, "fraction": surv / (qb * tb),
                 "z_same_cluster_median": float(np.sqrt(np.median(z2[same]))),
                 "z_other_cluster_median": float(np.sqrt(np.median(z2[~same])))}
print(json.dumps(res, indent=1))

# pair level and what breaks the tile test (last basis = span of cluster means)
pair_frac = float(hit.mean())
per_db_row = hit.mean(0)                      # fraction of queries a database row passes the filter for
hubs = np.argsort(-per_db_row)[: n // 100]
print(json.dumps({"pairs_passing_filter": pair_frac, "own_cluster_share": float(same.mean()),
                  "db_rows_passing_for_over_half_of_the_queries": int((per_db_row > 0.5).sum()),
                  "top_1pct_db_rows_mean_pass_rate": float(per_db_row[hubs].mean()),
                  "their_mean_centred_norm": float(np.sqrt(sq[hubs]).mean()), "all_rows_mean_centred_norm": float(np.sqrt(sq).mean())}, indent=1))

# the same tile test with the database (and the queries) ordered by Voronoi cells of 64 pivot rows in Z-space
piv = z[rng.choice(n, 64, replace=False)]
cell = np.argmin(((z[:, None, :] - piv[None, :, :]) ** 2).sum(-1) if n <= 4000 else
                 (zs[:, None] + (piv * piv).sum(1)[None, :] - 2.0 * (z @ piv.T)), axis=1)
o2 = np.argsort(cell, kind="stable")
hit2 = hit[o2][:, o2]
surv2 = 0
for a in range(qb):
    rows = hit2[a * 256:(a + 1) * 256]
    surv2 += int(np.add.reduceat(rows.any(0), np.arange(0, n, 128)).astype(bool).sum())
print(json.dumps({"tiles_surviving_with_z_space_cells": surv2, "tiles_total": qb * tb, "fraction": surv2 / (qb * tb)}))

#!/usr/bin/env python
"""Benchmark of the ViVAE hot path on B200: exact kNN (k=50) + smooth, cells/s; ViVAE.fit s/epoch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1]): synthetic 1,000,000 cells x 50 PCs (30-cluster mixture with a
PCA-like decaying spectrum, seed 1), exact kNN k=50, then smooth(k=50, coef=1, n_iter=1).  One "step" = one full pass
(graph + smoothing) over all cells.

Timed quantities
  value : cells/s with the matrix already resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e   : cells/s with HOST buffers in and out, H2D / D2H copies inside the timed region.  N = 1: the drop-in API
          (`vv.make_knn` + `vv.smooth`, NumPy in / NumPy out).  N > 1: `vv.sharded.knn_smooth_host_sharded` -- every
          rank uploads ONLY its rows from page-locked memory, the matrix is assembled by an NCCL all-gather, and the
          smoother is part of the path; the same function is also timed at N = 1 (`e2e_sharded_api`) so that the
          per-N numbers of one path can be compared.
  roofline : dominant kernel (tensor-core screening).  `frac` = EXECUTED MMA flops / measured dense peak (what the
          tensor pipe really did: tiles swept x 256 x 128 pairs x executed MACs); the ALGORITHMIC rate 2*nq*n*d/t --
          which exceeds the peak when the exact cell bounds skip most of the cross product -- is under `algorithmic`.
  config_c3 / config_c5 : the north-star configurations (10M x 50, k=100 at every N; 2M x 2000, k=50 at N >= 4),
          device-generated data, a few steps each, beside the headline without changing it.
  fit   : `ViVAE.fit` s/epoch at the headline's 1M rows, with the reference's fit (restated operation for operation
          in oracle/ref_fit.py) timed on this box's host cores and on this GPU in eager mode beside it.

N > 1 (launched under torchrun): the query rows are sharded across ranks (strong scaling: total work fixed); each
rank holds the whole matrix, the preparation of the cell plan is split across the ranks, the index shards are
all-gathered over NCCL (int32) and the Gauss-Seidel smoother -- which does not shard by rows but never mixes columns
 -- runs in COLUMN shards when the rows are wide (>= 64 columns per rank) and replicated on every
rank otherwise (at 50 columns a slice costs as many load instructions as the whole row); bit-identical either way.

`--impl reference` times the reference's CPU path for the same metric on the host cores: the oracle port (oracle/,
C + OpenMP) on a bounded sample -- the reference's own kNN (pynndescent) is not installable here and its smoother is
the single-threaded Python loop the oracle's C function restates.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- synthetic inputs
def synth_cells(n: int, d: int, seed: int = 1, n_clusters: int = 30) -> np.ndarray:
    """C2-syn: Gaussian mixture, per-dimension scale 5/sqrt(1+c)."""
    rng = np.random.default_rng(seed)
    spec = (5.0 / np.sqrt(1.0 + np.arange(d))).astype(np.float32)
    centres = rng.standard_normal((n_clusters, d)).astype(np.float32) * spec
    out = np.empty((n, d), dtype=np.float32)
    chunk = 1 << 18
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lab = rng.integers(0, n_clusters, size=e - s)
        out[s:e] = centres[lab] + 0.3 * spec * rng.standard_normal((e - s, d), dtype=np.float32)
    return out


def synth_pca_device(n: int, d: int, dev, seed: int = 2, n_clusters: int = 30):
    """C3-syn: the same mixture as C2-syn, drawn on the device (identical on every rank: same seed)."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    spec = 5.0 / torch.sqrt(1.0 + torch.arange(d, device=dev, dtype=torch.float32))
    centres = torch.randn(n_clusters, d, device=dev, generator=g) * spec
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    chunk = 1 << 20
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        lab = torch.randint(0, n_clusters, (e0 - s0,), device=dev, generator=g)
        out[s0:e0] = centres[lab] + 0.3 * spec * torch.randn(e0 - s0, d, device=dev, generator=g)
    return out


def synth_overlap_device(n: int, d: int, dev, seed: int = 7, n_clusters: int = 30):
    """The headline mixture with the within-cluster spread raised from 0.3 to 1.0 of the between-centre spread: the
    clusters overlap completely (one blob with the same decaying spectrum), so the cell bounds separate far fewer
    tiles -- the second workload VERDICT r1 asked for, so that the headline is not a property of 93 % pruning."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    spec = 5.0 / torch.sqrt(1.0 + torch.arange(d, device=dev, dtype=torch.float32))
    centres = torch.randn(n_clusters, d, device=dev, generator=g) * spec
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    chunk = 1 << 20
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        lab = torch.randint(0, n_clusters, (e0 - s0,), device=dev, generator=g)
        out[s0:e0] = centres[lab] + 1.0 * spec * torch.randn(e0 - s0, d, device=dev, generator=g)
    return out


def synth_cells_device(n: int, d: int, dev, seed: int = 4, n_clusters: int = 40):
    """C5-syn, generated on the device (a 2M x 2000 float32 matrix is 16 GB per rank: too large to draw on the host
    of every rank): non-negative log1p-transformed counts, ~65 % zeros, 40 clusters with Gamma-distributed
    expression programmes."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    shape = torch.full((n_clusters, d), 0.35, device=dev)
    centres = torch._standard_gamma(shape, generator=g) * 1.5  # many near-silent genes, a few strong ones
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    chunk = 1 << 16
    for s0 in range(0, n, chunk):
        e0 = min(n, s0 + chunk)
        lab = torch.randint(0, n_clusters, (e0 - s0,), device=dev, generator=g)
        depth = torch.exp(0.3 * torch.randn(e0 - s0, 1, device=dev, generator=g))
        out[s0:e0] = torch.log1p(torch.poisson(centres[lab] * depth, generator=g))
    return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            j = json.load(f)
        return {"hbm_gbs": j["hbm_gbs"], "tf": j["bf16_tflops"], "tf_sustained": j.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU arms
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc

    return orc


def cpu_baseline(x: np.ndarray, k: int, knn_queries: int, smooth_rows: int, keep: Optional[dict] = None,
                 q_first: int = 0) -> dict:
    """The oracle port of the reference path, timed on the host cores on a bounded sample: brute-force kNN of
    `knn_queries` query rows (starting at `q_first`) against the FULL matrix (all threads) plus the single-threaded
    smooth sweep over `smooth_rows` rows on a RANDOM neighbour graph of those rows (a timing sample: the sweep's cost
    does not depend on which rows are neighbours); per-cell costs are added."""
    orc = _oracle()
    n, d = x.shape
    knn_queries = min(knn_queries, n - q_first)
    smooth_rows = min(smooth_rows, n)
    orc.set_threads(os.cpu_count() or 1)  # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    t0 = time.perf_counter()
    o_idx, o_dist = orc.knn(x, k, q_first, q_first + knn_queries)
    t_knn = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    idx = rng.integers(0, smooth_rows, size=(smooth_rows, k), dtype=np.int64)
    idx[:, 0] = np.arange(smooth_rows)
    xs = np.ascontiguousarray(x[:smooth_rows])
    t0 = time.perf_counter()
    o_smooth = orc.smooth(xs, idx, k, 1.0, 1)
    t_sm = time.perf_counter() - t0
    if keep is not None:  # the checker's outputs, for the parity gates printed beside the numbers
        keep.update(knn_idx=o_idx, knn_dist=o_dist, smooth_x=xs, smooth_idx=idx, smooth_out=o_smooth, q_first=q_first)
    per_cell = t_knn / knn_queries + t_sm / smooth_rows
    return {"value": 1.0 / per_cell, "unit": "cells/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"kNN: {knn_queries} query rows vs all {n} rows ({t_knn:.2f} s, OpenMP x{orc.num_threads()}); "
                      f"smooth: first {smooth_rows} rows on a random neighbour graph, k={k}, 1 thread ({t_sm:.2f} s); "
                      "per-cell costs added",
            "seconds": t_knn + t_sm}


def ref_fit_cpu(x: np.ndarray, rows: int, batch: int) -> dict:
    """The reference's `fit` (oracle/ref_fit.py: restated operation for operation) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch

    import ref_fit

    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    det = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(False)
    try:
        r = ref_fit.time_ref_fit(np.ascontiguousarray(x[:rows]), batch_size=batch, epochs=1, device="cpu")
    finally:
        torch.use_deterministic_algorithms(det)
    r["kind"] = "port (oracle/ref_fit.py: the reference's fit restated op for op; reference model.py:130-177, losses.py:266-349)"
    return r


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x = synth_cells(args.n, args.d)
    vals = []
    base = None
    t_all0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(x, args.k, args.cpu_knn_queries, args.cpu_smooth_rows)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    fit = None if args.skip_fit else ref_fit_cpu(x, min(args.n, args.ref_fit_rows), args.fit_batch)
    line = {
        "impl": "reference", "metric": "cells/sec exact kNN(k=50)+smooth", "value": v, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.n / v,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fit": fit,
        "wall_s": time.perf_counter() - t_all0,
        "note": "ms_per_step is extrapolated to the full workload from the bounded sample (brute force is linear in "
                "queries); the smoother leg runs on a random neighbour graph of the first rows (timing only)",
    }
    print(json.dumps(line))


def workload_config(args, n_gpus: int) -> dict:
    return {"workload": f"synthetic {args.n} cells x {args.d} PCs, exact kNN k={args.k} + smooth(k={args.k}, coef=1, n_iter=1)",
            "n": args.n, "d": args.d, "k": args.k,
            "parallelism": (f"graph: query-row shards x{n_gpus} (cell plan prepared in row / cell shards), indices all-gathered "
                            f"over NCCL; smoother: " + (f"column shards x{n_gpus}" if args.d >= 64 * n_gpus else "replicated")
                            if getattr(args, "exchange", "replicated") != "ring" else
                            f"query-row AND database-row shards x{n_gpus} (NCCL ring exchange + running merge), smoother replicated"),
            "l2": f"inputs larger than L2 (matrix {args.n * args.d * 4 / 1e6:.0f} MB, graph "
                  f"{args.n * (args.k + 1) * 12 / 1e6:.0f} MB vs 126 MB L2)"}


def executed_macs_per_pair(dd: int, products: int = 3) -> int:
    """MACs the tensor pipe executes per (query, database row) pair: `products` fp16 products (three for the hi/lo
    split, one in the single-product first tier of wide rows), K padded to the MMA's k-step (16) for d + 3 <= 64, to
    the 64-column atom otherwise."""
    if dd + 3 <= 64:
        return 16 * ((dd + 3 + 15) // 16 + 2 * ((dd + 15) // 16))
    return products * 64 * ((dd + 3 + 63) // 64)


def screen_roofline(st: dict, nq: int, n: int, d: int, peaks: dict, sustained: bool = False) -> dict:
    """Roofline object of the screening kernel from the stats of one call (see the module docstring)."""
    ms = st.get("ms_screen") or st.get("ms_fallback")
    peak = (peaks["tf_sustained"] or peaks["tf"]) if sustained else peaks["tf"]
    alg = 2.0 * nq * n * d
    if st.get("algo_used") != 2:
        ach = alg / (ms / 1e3) / 1e12
        return {"kernel": "knn_brute_ffma", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None, "peak_source": peaks["source"], "ms": ms}
    macs = executed_macs_per_pair(d, st.get("products") or 3)
    # tiles of the threshold pre-pass run ONE product (d + 3 columns padded to the k-step of 16) since round 2
    tiles_pre = int(st.get("tiles_pre") or 0)
    macs_pre = 16 * ((d + 3 + 15) // 16)
    # (the few refine sweeps run three products: counted as one)
    ex = (float(st["tiles_swept"] - tiles_pre) * macs + float(tiles_pre) * macs_pre) * 256 * 128 * 2
    ach = ex / (ms / 1e3) / 1e12
    return {"kernel": "knn_screen_tcgen05", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak, "traffic": None, "peak_source": peaks["source"] + (" (sustained)" if sustained else " (burst)"),
            "ms": ms, "what": "EXECUTED MMA flops (tiles swept x 256 x 128 pairs x executed MACs x 2) / duration",
            "executed_macs_per_pair": macs, "algorithmic_macs_per_pair": d,
            "tiles_pre_pass_one_product": tiles_pre, "executed_macs_per_pair_pre_pass": macs_pre,
            "tiles_swept": st["tiles_swept"], "tiles_dense": st["tiles_dense"], "cells": st["cells"],
            "algorithmic": {"flops": alg, "tflops": alg / (ms / 1e3) / 1e12, "x_peak": alg / (ms / 1e3) / 1e12 / peak,
                            "note": "2*nq*n*d over the duration; NOT a utilisation: the exact cell bounds skipped "
                                    f"{100.0 * (1.0 - st['tiles_swept'] / max(1, st['tiles_dense'])):.1f} % of the tiles "
                                    "(tiles_swept includes the threshold pre-pass)"}}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import vivae_b200 as vv
    from vivae_b200 import _lib
    from vivae_b200.device import knn_device, smooth_device
    from vivae_b200.sharded import (all_gather_indices, knn_smooth_host_sharded, make_knn_ring, shard_bounds,
                                    smooth_auto_sharded)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: int) -> int:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    def knn_step(xd, kk, q0_, nq_, stats=None):
        """Graph rows [q0_, q0_+nq_) of the replicated matrix; at N > 1 the cell plan is prepared in row shards."""
        return knn_device(xd, kk, q0_, nq_, algo=args.algo, stats=stats, group=(dist.group.WORLD if world > 1 else None))

    def pipeline(xd, kk, q0_, nq_, x_local=None, stats=None):
        n_ = xd.shape[0]
        if args.exchange == "ring" and x_local is not None:
            idx, dist_ = make_knn_ring(x_local, n_, kk, algo=args.algo)
        else:
            idx, dist_ = knn_step(xd, kk, q0_, nq_, stats)
        if world > 1:
            idx = all_gather_indices(idx, n_)  # the smoother needs the whole graph (NCCL all-gather, int32)
            out = smooth_auto_sharded(xd, idx, k=kk, coef=1.0, n_iter=1)  # column shards only for wide rows
        else:
            out = smooth_device(xd, idx, k=kk, coef=1.0, n_iter=1)
        return idx, dist_, out

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        sync_all()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        return max_over_ranks(e0.elapsed_time(e1)) / steps, sum_over_ranks(_lib.launch_count() - l0)

    n, d, k = args.n, args.d, args.k
    device_data = args.device_data or n * d > 1_000_000_000
    if device_data:
        x_host = None
        x_dev = synth_cells_device(n, d, dev)
        args.skip_e2e = args.skip_cpu = args.skip_fit = True  # these legs need the host matrix
    else:
        x_host = synth_cells(n, d)
        x_dev = torch.from_numpy(x_host).to(dev)
    q0, nq = shard_bounds(n, world, rank)
    x_local = x_dev[q0:q0 + nq].contiguous() if args.exchange == "ring" else None

    # ---- headline: warm-up, then EXACTLY `steps` steps between CUDA events on the launching stream
    with ClockSampler(local) as clocks:
        ms_step, launches = timed(lambda: pipeline(x_dev, k, q0, nq, x_local), args.steps, args.warmup)
    value = n / (ms_step / 1e3)

    # ---- per-kernel split of one more (untimed) step, for the roofline of the dominant kernel
    st = {}
    sync_all()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    idx_s, dist_s = knn_step(x_dev, k, q0, nq, st)
    e1.record()
    torch.cuda.synchronize()
    ms_knn = e0.elapsed_time(e1)
    idx_full = all_gather_indices(idx_s, n) if world > 1 else idx_s
    torch.cuda.synchronize()
    e1.record()
    smooth_auto_sharded(x_dev, idx_full, k=k, coef=1.0, n_iter=1)
    e2.record()
    torch.cuda.synchronize()
    ms_smooth = max_over_ranks(e1.elapsed_time(e2))
    roofline = screen_roofline(st, nq, n, d, peaks)
    cap = os.path.join(ROOT, "profiles", "r02c_c2_raw.csv")
    if not os.path.isfile(cap):
        cap = os.path.join(ROOT, "profiles", "r02b_c2_raw.csv")
    if world == 1 and st.get("algo_used") == 2 and (n, d, k) == (1_000_000, 50, 50) and os.path.isfile(cap):
        import csv
        with open(cap) as f:
            rows_ = list(csv.reader(f))
        col = {h: i for i, h in enumerate(rows_[0])}
        gb = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        srow = next((r for r in rows_[2:] if "knn_screen" in r[col["Kernel Name"]]), None)
        if srow is not None:
            roofline["traffic"] = sum(float(srow[col[m]]) * gb[rows_[1][col[m]]]
                                      for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            roofline["traffic_note"] = ("dram__bytes_read + dram__bytes_write of this launch (1,000,000 query rows, one "
                                        "GPU) from the committed ncu --set full capture profiles/" + os.path.basename(cap))
    dense_roof = None
    if st.get("algo_used") == 2 and world == 1 and not args.skip_dense:
        # the same kernel with the partition switched off (every tile swept): the dense tensor roofline
        os.environ["VVB_TC_CELLS"] = "1"
        try:
            for _ in range(2):
                sd = {}
                knn_device(x_dev, k, q0, nq, algo=args.algo, stats=sd)
                torch.cuda.synchronize()
        finally:
            del os.environ["VVB_TC_CELLS"]
        dense_roof = screen_roofline(sd, nq, n, d, peaks)
        dense_roof["kernel"] += " (VVB_TC_CELLS=1: no tile skipped)"
    sbytes = float(n) * (k * d * 4 + k * 4 + d * 4)
    smooth_roof = {"kernel": "smooth_sweep_gs" + (f" (column shards x{world} + all-gather)" if world > 1 and d >= 64 * world else
                                                  " (replicated on every rank)" if world > 1 else ""), "bound": "hbm", "achieved": sbytes / (ms_smooth / 1e3) / 1e9,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": sbytes / (ms_smooth / 1e3) / 1e9 / peaks["hbm_gbs"],
                   "ms": ms_smooth, "algorithmic_bytes": sbytes, "traffic": None}

    # ---- end to end with host buffers
    e2e = e2e_sharded = None
    if not args.skip_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        # (a) the sharded host API: every rank passes only its rows, page-locked; same function at every N
        xs_pin = torch.from_numpy(x_host[q0:q0 + nq]).pin_memory()
        outs = (torch.empty((nq, k + 1), dtype=torch.int64).pin_memory(), torch.empty((nq, k + 1), dtype=torch.float32).pin_memory(),
                torch.empty((nq, d), dtype=torch.float32).pin_memory())

        def step_sharded():
            o = knn_smooth_host_sharded(xs_pin, n, k, coef=1.0, n_iter=1, algo=args.algo, out=outs,
                                        _knn_fn=lambda xx, kk, a, b, alg: knn_step(xx, kk, a, b))
            return float(o[2][0, 0]) if nq else 0.0

        for _ in range(3):
            step_sharded()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_sharded()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        h2d = nq * d * 4
        d2h = nq * (k + 1) * 12 + nq * d * 4
        e2e_sharded = {"value": n / dt, "unit": "cells/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": dt * 1e3, "steps": e2e_steps, "bytes_are": "per rank",
                       "path": "vv.sharded.knn_smooth_host_sharded: upload of the rank's rows (page-locked) -> NCCL all-gather "
                               "of the matrix -> graph -> all-gather of indices -> smooth -> download of the rank's graph "
                               "rows and smoothed rows"}
        del xs_pin, outs
        if world == 1:
            # (b) the drop-in API (what a user of the reference calls): NumPy in, NumPy out
            def step_dropin():
                knn = vv.make_knn(x_host, k=k, verbose=False, algo=args.algo)
                out = vv.smooth(x_host, knn, k=k, coef=1.0, n_iter=1)
                return float(out[0, 0])

            for _ in range(3):  # warm-up: device buffers, bounce buffers and the pool of page-locked result blocks
                step_dropin()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                step_dropin()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            # bytes that cross PCIe: neighbour indices travel as int32 (widened / narrowed in the staging copy)
            e2e = {"value": n / dt, "unit": "cells/s", "h2d_bytes_per_step": int(2 * n * d * 4 + n * (k + 1) * 4),
                   "d2h_bytes_per_step": int(n * (k + 1) * 8 + n * d * 4), "ms_per_step": dt * 1e3, "steps": e2e_steps,
                   "path": "vv.make_knn + vv.smooth (host NumPy in/out; the graph goes down and up again between the calls)"}
        else:
            e2e = e2e_sharded

    # ---- quartet loss at a bandwidth-sized batch (rank 0): achieved GB/s of forward + backward, C ABI calls
    quartet = None
    if rank == 0 and not args.skip_fit:
        lib = _lib.load()
        Bq = min(n, 1 << 20)
        xq_ = x_dev[:Bq].contiguous()
        zq_ = torch.randn(Bq, 2, device=dev)
        gu = torch.empty(Bq, 2, device=dev)
        gz = torch.empty(Bq, 2, device=dev)
        loss_ = torch.empty((), device=dev)
        gout = torch.ones((), device=dev)
        wsq = torch.empty(int(lib.vvb_quartet_workspace_bytes(Bq)), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream().cuda_stream

        def q_step():
            _lib.check(lib.vvb_quartet_loss_fwd_device(xq_.data_ptr(), zq_.data_ptr(), Bq, d, 2, None, 1, 0, 0,
                                                       loss_.data_ptr(), gu.data_ptr(), wsq.data_ptr(), stream))
            _lib.check(lib.vvb_quartet_loss_bwd_device(gu.data_ptr(), gout.data_ptr(), Bq * 2, gz.data_ptr(), stream))

        for _ in range(3):
            q_step()
        torch.cuda.synchronize()
        q0e, q1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0e.record()
        for _ in range(20):
            q_step()
        q1e.record()
        torch.cuda.synchronize()
        msq = q0e.elapsed_time(q1e) / 20
        qbytes = float(Bq) * (4 * (d + 2) + 4 * (2 + 2 * 2))  # fwd reads x,z; writes unit grad; bwd reads it, writes dz
        quartet = {"batch": Bq, "ms_fwd_bwd": msq, "achieved_gbs": qbytes / (msq / 1e3) / 1e9,
                   "peak_gbs": peaks["hbm_gbs"], "frac": qbytes / (msq / 1e3) / 1e9 / peaks["hbm_gbs"],
                   "algorithmic_bytes": qbytes, "how": "vvb_quartet_loss_fwd_device + bwd_device, 20 calls between CUDA events"}
        del xq_, zq_, gu, gz, wsq

    # ---- fit (the metric's second half: s/epoch), rank 0 only, at the headline's row count
    fit = None
    if rank == 0 and not args.skip_fit:
        rows = min(n, args.fit_rows)
        model = vv.ViVAE(input_dim=d, random_state=1)
        xf = x_host[:rows]
        model.fit(xf, n_epochs=1, batch_size=args.fit_batch, lam_mds=100.0, verbose=False)  # upload, capture, warm-up
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        for _ in range(args.fit_epochs):  # the epochs fit() runs after its one-off set-up (reference: model.py:258-270)
            model.train_epoch(lam_mds=100.0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.fit_epochs
        fit = {"rows": rows, "batch_size": args.fit_batch, "s_per_epoch": dt, "cells_per_s": rows / dt,
               "epochs_timed": args.fit_epochs,
               "quartet_kernel_launches_per_epoch": (_lib.launch_count() - l0) / args.fit_epochs}
        del model
        if world == 1 and not args.skip_cpu:
            # the reference's fit beside it, same box: host cores, and this GPU in eager mode (dispatch-bound)
            rr = min(rows, args.ref_fit_rows)
            fit["reference_cpu"] = ref_fit_cpu(x_host, rr, args.fit_batch)
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import ref_fit
            det = torch.are_deterministic_algorithms_enabled()
            torch.use_deterministic_algorithms(False)
            try:
                fit["reference_eager_gpu"] = ref_fit.time_ref_fit(np.ascontiguousarray(x_host[:rr]), batch_size=args.fit_batch,
                                                                 epochs=1, device=f"cuda:{local}")
            finally:
                torch.use_deterministic_algorithms(det)
            fit["speedup_vs_reference_cpu"] = fit["cells_per_s"] / fit["reference_cpu"]["cells_per_s"]
            fit["speedup_vs_reference_eager_gpu"] = fit["cells_per_s"] / fit["reference_eager_gpu"]["cells_per_s"]

    # ---- CPU baseline (N = 1 only) and parity gates (every N: the oracle is the checker, rank 0's own shard)
    base = None
    gates = None
    if rank == 0 and not args.skip_cpu:
        kept = {}
        qn = args.cpu_knn_queries if world == 1 else min(args.cpu_knn_queries, 1024)
        sr = args.cpu_smooth_rows if world == 1 else min(args.cpu_smooth_rows, 50_000)
        b = cpu_baseline(x_host, k, min(qn, nq), sr, keep=kept, q_first=q0)
        if world == 1:
            base = b
        qn = kept["knn_idx"].shape[0]
        g_idx, g_dist = idx_s[:qn].cpu().numpy(), dist_s[:qn].cpu().numpy()
        rows_equal = (g_idx == kept["knn_idx"]).all(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(g_dist - kept["knn_dist"]) / np.where(kept["knn_dist"] > 0, kept["knn_dist"], 1.0)
        g_sm = smooth_device(torch.from_numpy(kept["smooth_x"]).to(dev), torch.from_numpy(kept["smooth_idx"]).to(dev), k=k,
                             coef=1.0, n_iter=1).cpu().numpy()
        gates = {"knn_rows_checked": int(qn), "knn_rows_exact": float(rows_equal.mean()),
                 "knn_dist_max_rel_err": float(rel.max()), "knn_rows_fallback": int(st.get("rows_fallback", 0)),
                 "knn_rows_refined": int(st.get("rows_refined", 0)),
                 "smooth_rows_checked": int(g_sm.shape[0]),
                 "smooth_max_abs_err": float(np.abs(g_sm - kept["smooth_out"]).max()),
                 "checked": f"rank 0's first {qn} graph rows of this run and a {g_sm.shape[0]}-row smoother sample vs the CPU oracle"}

    # ---- the same graph with the DATABASE sharded too: every rank holds only its rows, shards travel around an NCCL
    # ring (north star; `make_knn_ring`).  Measured beside the replicated layout at every N > 1 and compared bit for bit.
    ring = None
    if world > 1 and not args.skip_ring and args.exchange != "ring":
        xl = x_dev[q0:q0 + nq].contiguous()
        ms_ring, _ = timed(lambda: make_knn_ring(xl, n, k, algo=args.algo), 5, 3)
        r_idx, r_dist = make_knn_ring(xl, n, k, algo=args.algo)
        same = torch.tensor([int(torch.equal(r_idx, idx_s) and torch.equal(r_dist, dist_s))], device=dev)
        # the ring schedule proper (what a matrix beyond one GPU's memory pays): one foreign shard per sweep
        ms_ring1, _ = timed(lambda: make_knn_ring(xl, n, k, algo=args.algo, shards_per_sweep=1), 5, 3)
        r1_idx, r1_dist = make_knn_ring(xl, n, k, algo=args.algo, shards_per_sweep=1)
        same &= int(torch.equal(r1_idx, idx_s) and torch.equal(r1_dist, dist_s))
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        ring = {"what": "make_knn_ring: each rank holds only its n/N rows; graph only, no smoother.  ms_knn: the default "
                        "schedule -- every shard fits next to the own rows here, so they arrive with one all-gather and ONE "
                        "sweep ranks the own rows (no running lists, no merges); ms_knn_one_shard_per_sweep: the ring "
                        "schedule (isend/irecv of the shard overlapped with the step's sweep, running top-k merge)",
                "ms_knn": ms_ring, "ms_knn_one_shard_per_sweep": ms_ring1,
                "ms_knn_replicated_layout": max_over_ranks(ms_knn), "steps": 5, "warmup": 3,
                "identical_to_replicated_layout": bool(same.item()),
                "note": "a sweep pruned by the cell bounds costs about the same per query block whatever the database "
                        "size, so the number of sweeps is what counts (DESIGN.md section 6)"}
        del xl, r_idx, r_dist, r1_idx, r1_dist

    # ---- the north-star configurations beside the headline (device-generated data, a few steps)
    del idx_s, dist_s, idx_full
    x_dev = None
    torch.cuda.empty_cache()
    sub = {}

    def sub_config(name, xs, kk, steps, with_smooth=True, label="synthetic"):
        n_, d_ = xs.shape
        a0, an = shard_bounds(n_, world, rank)
        fn = (lambda: pipeline(xs, kk, a0, an)) if with_smooth else (lambda: knn_step(xs, kk, a0, an))
        ms, _ = timed(fn, steps, 1)
        s_ = {}
        knn_step(xs, kk, a0, an, s_)
        torch.cuda.synchronize()
        ms_kn = max_over_ranks(s_["ms_prepare"] + s_["ms_screen"] + s_["ms_rerank"] + s_["ms_fallback"])
        rf = screen_roofline(s_, an, n_, d_, peaks, sustained=True)
        sub[name] = {"workload": f"{label} {n_} x {d_}, exact kNN k={kk}" + (f" + smooth(k={kk}, coef=1, n_iter=1)" if with_smooth else " (graph only)"),
                     "n_gpus": world, "steps": steps, "warmup": 1, "ms_per_step": ms, "value": n_ / (ms / 1e3), "unit": "cells/s",
                     "data": "synthetic (generated on the device, identical on every rank)", "ms_knn_rank0": ms_kn,
                     "knn_stats_rank0": s_, "roofline": rf}

    if not args.skip_c3 and args.c3_rows > 0:
        xs = synth_pca_device(args.c3_rows, 50, dev)
        sub_config("config_c3", xs, 100, args.sub_steps)
        if world > 1 and not args.skip_ring:
            # the ring build at the north star's size too ("ring <= 1.15 x replicated at 10M")
            a0, an = shard_bounds(args.c3_rows, world, rank)
            xl = xs[a0:a0 + an].contiguous()
            i_rep, d_rep = knn_step(xs, 100, a0, an)
            ms_rep, _ = timed(lambda: knn_step(xs, 100, a0, an), 2, 1)
            ms_ring3, _ = timed(lambda: make_knn_ring(xl, args.c3_rows, 100, algo=args.algo), 2, 2)
            r_idx, r_dist = make_knn_ring(xl, args.c3_rows, 100, algo=args.algo)
            same = torch.tensor([int(torch.equal(r_idx, i_rep) and torch.equal(r_dist, d_rep))], device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            sub["config_c3"]["ring"] = {"ms_knn": ms_ring3, "ms_knn_replicated_layout": ms_rep,
                                        "ratio": ms_ring3 / ms_rep, "steps": 2, "warmup": 2,
                                        "identical_to_replicated_layout": bool(same.item()),
                                        "what": "make_knn_ring (every rank holds only its n/N rows) beside the replicated "
                                                "layout, graph only, max over ranks"}
            del xl, r_idx, r_dist, i_rep, d_rep
        del xs
        torch.cuda.empty_cache()
    if not args.skip_overlap:
        # second, less separable workload at the headline's size (same n, d, k, same pipeline)
        xs = synth_overlap_device(args.n, args.d, dev)
        sub_config("config_overlap", xs, args.k, args.sub_steps,
                   label="synthetic overlapping mixture (within-cluster spread 1.0 of the between-centre spread)")
        so = sub["config_overlap"]["knn_stats_rank0"]
        if so.get("tiles_dense"):
            sub["config_overlap"]["tiles_swept_frac_rank0"] = so.get("tiles_swept", 0) / so["tiles_dense"]
        del xs
        torch.cuda.empty_cache()
    if not args.skip_c5 and args.c5_rows > 0:
        if world >= 4:
            xs = synth_cells_device(args.c5_rows, 2000, dev)
            sub_config("config_c5", xs, 50, 1, with_smooth=False)
            del xs
            torch.cuda.empty_cache()
        else:
            sub["config_c5"] = {"skipped": "BASELINE quotes 2M x 2000 at 8 GPUs; measured here at N >= 4 only "
                                           "(one GPU needs about a minute per step)"}

    if rank == 0:
        line = {
            "metric": "cells/sec exact kNN(k=50)+smooth", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 (screening operands f16 hi/lo, f32 accumulate; re-rank f32)",
            "data": "synthetic (generated on the device)" if device_data else "synthetic",
            "config": workload_config(args, world), "clocks": clocks.summary(),
            "e2e": e2e, "e2e_sharded_api": e2e_sharded, "gpu_launches": launches, "roofline": roofline,
            "roofline_dense": dense_roof, "roofline_smooth": smooth_roof, "quartet": quartet,
            "knn_stats": st, "ms_knn": ms_knn, "ms_smooth": ms_smooth, "cpu_baseline": base, "fit": fit,
            "parity_gates": gates, "ring": ring, **sub,
            "parity": {"knn": "bit-exact vs FP32 brute-force oracle (tests/test_gpu_parity.py, tests/test_gpu_configs.py)",
                       "smooth": "bit-exact vs reference smooth()", "quartet": "1e-4 relative"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", "--n", dest="n", type=int, default=1_000_000)  # use --cells under torchrun
    ap.add_argument("--dims", "--d", dest="d", type=int, default=50)
    ap.add_argument("--neighbours", "--k", dest="k", type=int, default=50)
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--exchange", default="replicated", choices=["replicated", "ring"],
                    help="N>1: every rank holds the matrix (default) or only its rows, shards passed around an NCCL ring")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-fit", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-dense", action="store_true")
    ap.add_argument("--skip-c3", action="store_true")
    ap.add_argument("--skip-c5", action="store_true")
    ap.add_argument("--skip-overlap", action="store_true")
    ap.add_argument("--skip-ring", action="store_true")
    ap.add_argument("--c3-rows", type=int, default=10_000_000)
    ap.add_argument("--c5-rows", type=int, default=2_000_000)
    ap.add_argument("--sub-steps", type=int, default=2)
    ap.add_argument("--device-data", action="store_true")
    ap.add_argument("--fit-rows", type=int, default=1_000_000)
    ap.add_argument("--ref-fit-rows", type=int, default=100_000)
    ap.add_argument("--fit-batch", type=int, default=512)
    ap.add_argument("--fit-epochs", type=int, default=2)
    ap.add_argument("--cpu-knn-queries", type=int, default=4096)
    ap.add_argument("--cpu-smooth-rows", type=int, default=200_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus != int(os.environ.get("WORLD_SIZE", "1")) and int(os.environ.get("WORLD_SIZE", "1")) == 1 and args.gpus > 1:
            # convenience: `python bench.py --gpus N` re-launches itself under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        run_ours(args)


if __name__ == "__main__":
    main()

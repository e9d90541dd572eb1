"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the
C ABI (`vivae_b200` -> ctypes -> libvivae_b200.so), against the CPU oracle and the
golden vectors recorded from the unmodified reference.

Bars (BASELINE.json north_star): neighbour indices bit-exact; smoothed values
bit-exact in Gauss-Seidel mode (bar: 1e-5 relative); quartet loss and gradient within
1e-4 relative."""

import glob
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

import oracle as orc
from conftest import GOLDEN

import vivae_b200 as vv

pytestmark = pytest.mark.gpu

SMOOTH = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "smooth_*.npz")))
QUARTET = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "quartet_*.npz")))


def clustered(n, d, n_clusters, seed, scale=5.0):
    """PCA-like synthetic cells: Gaussian mixture with a decaying per-dimension spectrum."""
    rng = np.random.default_rng(seed)
    spec = scale / np.sqrt(1.0 + np.arange(d))
    centres = rng.standard_normal((n_clusters, d)) * spec
    lab = rng.integers(0, n_clusters, size=n)
    return (centres[lab] + 0.3 * spec * rng.standard_normal((n, d))).astype(np.float32)


def tie_data(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 4, size=(n, d)).astype(np.float32)
    dup = rng.choice(n, size=max(2, n // 100), replace=False)
    x[dup] = x[dup[0]]  # 1 % exact duplicates of one point
    return x


def assert_knn_equal(got, want):
    gi, gd = got
    wi, wd = want
    assert gi.dtype == np.int64 and gd.dtype == np.float32
    bad = np.flatnonzero((gi != wi).any(axis=1))
    assert bad.size == 0, f"{bad.size} rows differ, first {bad[:5]}: got {gi[bad[0]]} want {wi[bad[0]]}"
    assert np.array_equal(gd, wd), f"distances differ, max abs {np.abs(gd - wd).max()}"


# ============================================================================ kNN
@pytest.mark.parametrize("algo", ["brute", "auto"])
@pytest.mark.parametrize("n,d,k,kind", [
    (3000, 50, 50, "clustered"),
    (2000, 38, 100, "clustered"),
    (1500, 7, 10, "ties"),
    (600, 3, 50, "ties"),
    (257, 100, 17, "clustered"),   # d not a multiple of 16
    (129, 1, 5, "gauss"),          # d = 1
    (64, 5, 63, "gauss"),          # k = n-1
    (2, 4, 1, "gauss"),            # smallest legal problem
])
def test_knn_bit_exact_vs_oracle(algo, n, d, k, kind):
    if kind == "clustered":
        x = clustered(n, d, 8, seed=n + d)
    elif kind == "ties":
        x = tie_data(n, d, seed=n + d)
    else:
        x = np.random.default_rng(n + d).standard_normal((n, d)).astype(np.float32)
    st = {}
    got = vv.make_knn(x, k=k, verbose=False, algo=algo, stats=st)
    assert_knn_equal(got, orc.knn(x, k))
    assert st["rows"] == n and st["kernel_launches"] >= 1
    assert np.array_equal(got[0][:, 0], np.arange(n)) and np.all(got[1][:, 0] == 0)


@pytest.mark.parametrize("algo", ["brute", "auto"])
def test_knn_large_offset_data_and_float64_input(algo):
    # un-centred data far from the origin (worst case for a GEMM-expansion distance) and
    # float64 input (R/reticulate passes doubles): coerced to float32 like pynndescent does
    x = clustered(4000, 50, 6, seed=3).astype(np.float64) + 1000.0
    got = vv.make_knn(x, k=30, verbose=False, algo=algo)
    assert_knn_equal(got, orc.knn(x.astype(np.float32), 30))


@pytest.mark.parametrize("algo", ["brute", "auto"])
def test_knn_many_duplicates_more_than_k(algo):
    # 200 copies of one point with k = 20: ties at distance 0 must resolve to the LOWEST indices
    x = clustered(1000, 10, 3, seed=9)
    x[300:500] = x[300]
    got = vv.make_knn(x, k=20, verbose=False, algo=algo)
    assert_knn_equal(got, orc.knn(x, 20))
    assert np.array_equal(got[0][450, :4], [450, 300, 301, 302])


@pytest.mark.parametrize("algo", ["brute", "auto"])
def test_knn_query_row_shard_through_c_abi(algo):
    from vivae_b200.knn import _knn_host

    x = clustered(5000, 50, 5, seed=12)
    want = orc.knn(x, 25)
    for q0, nq in ((0, 1000), (1000, 3000), (4000, 1000), (4999, 1), (17, 0)):
        gi, gd = _knn_host(x, 25, q0, nq, algo, None)
        assert gi.shape == (nq, 26)
        assert np.array_equal(gi, want[0][q0:q0 + nq]) and np.array_equal(gd, want[1][q0:q0 + nq])


def test_knn_medium_scale_properties_and_sampled_oracle():
    """200k x 50, k=50 (too big for a full CPU oracle in seconds): size-independent
    properties on all rows + bit-exact check of a query sample against the oracle."""
    n, d, k = 200_000, 50, 50
    x = clustered(n, d, 30, seed=1)
    st = {}
    idx, dist = vv.make_knn(x, k=k, verbose=False, stats=st)
    assert np.array_equal(idx[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist, axis=1) >= 0)                       # sortedness
    assert idx.min() >= 0 and idx.max() < n
    srt = np.sort(idx, axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])                        # no repeated neighbour
    rng = np.random.default_rng(0)
    rows = rng.choice(n, 64, replace=False)
    d_chk = np.sqrt(((x[rows, None, :].astype(np.float64) - x[idx[rows]].astype(np.float64)) ** 2).sum(-1))
    np.testing.assert_allclose(dist[rows], d_chk, rtol=2e-6, atol=1e-6)
    for q0 in (0, 77_777, n - 256):
        wi, wd = orc.knn(x, k, q0, q0 + 256)
        assert np.array_equal(idx[q0:q0 + 256], wi) and np.array_equal(dist[q0:q0 + 256], wd)
    print("knn stats:", st)


def test_full_size_config_1M_x_50_properties_sampled_oracle_and_smooth():
    """BASELINE.json configs[1] at full size (1M cells x 50 PCs, k=50, smooth k=50): size-independent properties on
    every row, bit-exact oracle check of three query slices, the whole smoothed matrix bit-exact against the oracle's
    sweep, and the device-resident entry point equal to the host one."""
    from vivae_b200.device import knn_device

    n, d, k = 1_000_000, 50, 50
    x = clustered(n, d, 30, seed=1)
    st = {}
    idx, dist = vv.make_knn(x, k=k, verbose=False, stats=st)
    assert st["algo_used"] == 2 and st["rows_fallback"] <= 1000 and st["tiles_swept"] < 0.25 * st["tiles_dense"], st
    assert np.array_equal(idx[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist, axis=1) >= 0)
    assert idx.min() >= 0 and idx.max() < n
    srt = np.sort(idx[::7], axis=1)
    assert np.all(srt[:, 1:] != srt[:, :-1])
    for q0 in (0, 500_123, n - 128):
        wi, wd = orc.knn(x, k, q0, q0 + 128)
        assert np.array_equal(idx[q0:q0 + 128], wi) and np.array_equal(dist[q0:q0 + 128], wd)
    di, dd = knn_device(torch.from_numpy(x).cuda(), k)
    assert np.array_equal(di.cpu().numpy(), idx) and np.array_equal(dd.cpu().numpy(), dist)
    del di, dd
    out = vv.smooth(x, [idx, dist], k=k, coef=1.0, n_iter=1)
    assert np.array_equal(out, orc.smooth(x, idx, k, 1.0, 1))


def test_make_knn_cache_write_then_load(tmp_path):
    x = clustered(500, 8, 3, seed=4)
    f = str(tmp_path / "graph.npy")
    buf = io.StringIO()
    with redirect_stdout(buf):
        a = vv.make_knn(x, fname=f, k=12)
        b = vv.make_knn(x, fname=f, k=12)
    assert buf.getvalue().split() == ["Constructing", "k-NNG", "Loading", "k-NNG"]
    raw = np.load(f)
    assert raw.dtype == np.float64 and raw.shape == (2, 500, 13)     # the reference's cache format
    assert np.array_equal(a[0], b[0]) and np.allclose(a[1], b[1])
    out = vv.smooth(x, b, k=10, coef=1.0)                            # loaded object is a valid `knn`
    assert out.shape == x.shape


# ========================================================================= smooth
@pytest.mark.parametrize("name", SMOOTH)
def test_smooth_bit_exact_vs_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    out = vv.smooth(g["x"], [g["idx"].astype(np.int64), g["dist"]], k=int(g["k"]), coef=float(g["coef"]),
                    n_iter=int(g["n_iter"]))
    assert out.dtype == g["out"].dtype
    assert np.array_equal(out, g["out"]), f"max abs diff {np.abs(out - g['out']).max()}"


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ["random", "sorted"])
def test_smooth_bit_exact_vs_oracle_100k(dtype, order):
    """100k x 50, k=50.  'sorted' orders rows along coordinate 0 (trajectory-like data), the
    adversarial case for the dependency depth of the Gauss-Seidel sweep."""
    n, d, k = 100_000, 50, 50
    x = clustered(n, d, 20, seed=5)
    if order == "sorted":
        x = x[np.argsort(x[:, 0])]
    idx, dist = vv.make_knn(x, k=k, verbose=False)
    xin = x.astype(dtype)
    for coef, n_iter in ((1.0, 1), (0.1, 2)):
        out = vv.smooth(xin, [idx, dist], k=k, coef=coef, n_iter=n_iter)
        want = orc.smooth(xin, idx, k, coef, n_iter)
        assert out.dtype == dtype
        assert np.array_equal(out, want), f"{(out != want).sum()} elements differ, max abs {np.abs(out - want).max()}"
    assert np.array_equal(xin, x.astype(dtype))  # input not mutated


def test_smooth_jacobi_mode_and_wide_rows():
    x = clustered(3000, 300, 5, seed=8)  # d > 128: several column passes per row
    idx, dist = vv.make_knn(x, k=20, verbose=False)
    for mode, m in (("gauss_seidel", 0), ("jacobi", 1)):
        out = vv.smooth(x, [idx, dist], k=15, coef=0.7, n_iter=2, mode=mode)
        assert np.array_equal(out, orc.smooth(x, idx, 15, 0.7, 2, mode=m))
    with pytest.raises(ValueError):
        vv.smooth(x, [idx, dist], k=15, mode="nope")


def test_smooth_is_identity_for_coef_zero_and_idempotent_on_constant_data():
    x = clustered(2000, 12, 3, seed=2)
    knn = vv.make_knn(x, k=10, verbose=False)
    assert np.array_equal(vv.smooth(x, knn, k=10, coef=0.0), x)
    c = np.full((500, 6), 3.25, dtype=np.float32)
    idx = np.stack([np.roll(np.arange(500), -s) for s in range(8)], axis=1)
    assert np.array_equal(vv.smooth(c, [idx, np.zeros((500, 8), np.float32)], k=8, coef=1.0, n_iter=3), c)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mode,m", [("gauss_seidel", 0), ("jacobi", 1)])
def test_smooth_pipelined_host_path_equals_single_launch_and_oracle(monkeypatch, dtype, mode, m):
    """vvb_smooth_host with one sweep and a large neighbour matrix runs the sweep in row chunks while the
    neighbour rows of the next chunk upload and finished rows download; same bytes as the single launch."""
    n, d, k = 160_000, 24, 40  # 160k x 41 int64 = 52 MB: above the pipeline threshold, ragged last chunk
    x = clustered(n, d, 12, seed=41).astype(dtype)
    idx, dist = vv.make_knn(x.astype(np.float32), k=k, verbose=False)
    single = vv.smooth(x, [idx, dist], k=k, coef=0.6, n_iter=1, mode=mode)
    monkeypatch.setenv("VVB_SMOOTH_PIPELINE", "1")  # opt-in: neutral on the bench host (host-memory bound)
    piped = vv.smooth(x, [idx, dist], k=k, coef=0.6, n_iter=1, mode=mode)
    assert piped.dtype == dtype and np.array_equal(piped, single)
    assert np.array_equal(piped, orc.smooth(x, idx, k, 0.6, 1, mode=m))
    bad = idx.copy()
    bad[n - 5, 7] = n + 3  # in the last chunk
    with pytest.raises(ValueError, match="must include self"):
        vv.smooth(x, [bad, dist], k=k, coef=0.6)
    monkeypatch.delenv("VVB_SMOOTH_PIPELINE")


def test_results_live_in_recycled_page_locked_blocks(monkeypatch):
    """Large result arrays are NumPy views of page-locked blocks (direct DMA instead of a staged copy); a block
    returns to the pool when its arrays die and is reused by the next call; VVB_PINNED_RESULTS=0 gives plain
    arrays with the same contents; pinned arrays are accepted as inputs (smooth reads the graph directly)."""
    import gc

    x = clustered(60_000, 50, 10, seed=3)

    def owner(a):
        while isinstance(a, np.ndarray) and a.base is not None:
            a = a.base
        return a

    # idle blocks left by earlier tests of the same size class would be handed out first: start from an empty pool
    import gc as _gc

    _gc.collect()
    from vivae_b200 import _lib as _vl

    _vl._pinned_pool.release()
    _vl._pinned_pool.seen.clear()  # (and the very first request of a size class is never pinned)
    first = vv.make_knn(x, k=50, verbose=False)
    idx, dist = vv.make_knn(x, k=50, verbose=False)
    assert type(idx) is np.ndarray and idx.flags.c_contiguous and idx.flags.writeable and idx.dtype == np.int64
    blk = owner(idx)
    assert type(blk).__name__ == "_PinnedBlock" and type(owner(dist)).__name__ == "_PinnedBlock"
    ptr = blk.ptr
    sm = vv.smooth(x, [idx, dist], k=50, coef=1.0)
    want_sm = orc.smooth(x, idx, 50, 1.0, 1)
    assert np.array_equal(sm, want_sm) and np.array_equal(first[0], idx) and np.array_equal(first[1], dist)
    del first
    keep_idx, keep_dist = idx.copy(), dist.copy()
    del idx, dist, blk
    gc.collect()
    idx2, dist2 = vv.make_knn(x, k=50, verbose=False)
    assert owner(idx2).ptr == ptr, "freed block was not recycled"
    assert np.array_equal(idx2, keep_idx) and np.array_equal(dist2, keep_dist)
    monkeypatch.setenv("VVB_PINNED_RESULTS", "0")
    idx3, dist3 = vv.make_knn(x, k=50, verbose=False)
    assert not isinstance(owner(idx3), type(owner(idx2))) and np.array_equal(idx3, keep_idx) and np.array_equal(dist3, keep_dist)
    assert np.array_equal(vv.smooth(x, [idx3, dist3], k=50, coef=1.0), want_sm)
    wi, wd = orc.knn(x, 50, 1000, 1400)
    assert np.array_equal(keep_idx[1000:1400], wi) and np.array_equal(keep_dist[1000:1400], wd)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_smooth_rows_entry_point_jacobi_shards_equal_whole_sweep(dtype):
    """vvb_smooth_rows_device: row ranges of a Jacobi sweep computed separately (what each rank of
    smooth_jacobi_sharded does) equal the whole-matrix sweep; a Gauss-Seidel range is refused; bad indices raise."""
    from vivae_b200.device import smooth_rows_device
    from vivae_b200.sharded import shard_bounds, smooth_jacobi_sharded

    n, d, k = 20_000, 30, 25
    x = clustered(n, d, 7, seed=9).astype(dtype)
    idx, dist = vv.make_knn(x.astype(np.float32), k=k, verbose=False)
    want = orc.smooth(x, idx, k, 0.5, 1, mode=1)
    xd, idd = torch.tensor(x, device="cuda"), torch.tensor(idx, device="cuda")
    parts = []
    for rank in range(3):
        q0, nq = shard_bounds(n, 3, rank)
        parts.append(smooth_rows_device(xd, idd[q0:q0 + nq].contiguous(), k, 0.5, q0))
    assert np.array_equal(torch.cat(parts).cpu().numpy(), want)
    # single-process call of the sharded driver (world = 1): two sweeps
    two = smooth_jacobi_sharded(xd, idd, k, coef=0.5, n_iter=2)
    assert np.array_equal(two.cpu().numpy(), orc.smooth(x, idx, k, 0.5, 2, mode=1))
    bad = idd[:100].clone()
    bad[5, 3] = n + 1
    with pytest.raises(ValueError, match="must include self"):
        smooth_rows_device(xd, bad, k, 0.5, 0)
    lib = vv._lib.load()
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    rc = lib.vvb_smooth_rows_device(xd.data_ptr(), xd.element_size(), n, d, idd.data_ptr(), idd.shape[1], k, 0.5, 0, 0, 10,
                                    ws.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == vv._lib.ENOTSUP


# ======================================================================== quartet
# `big`: VVB_QUARTET_BIG forces the kernel -- "0" eight lanes per quartet (training-size batches), "1" 32 quartets per warp
# (bandwidth-sized batches; the launcher picks it from 2^17-ish quartets on): both must meet every fixture.
@pytest.mark.parametrize("big", ["0", "1"])
@pytest.mark.parametrize("name", QUARTET)
def test_quartet_loss_and_grad_vs_reference_golden(monkeypatch, name, big):
    monkeypatch.setenv("VVB_QUARTET_BIG", big)
    g = np.load(os.path.join(GOLDEN, name))
    ns = int(g["n_sampling"])
    dev = torch.device("cuda")
    x = torch.tensor(g["x"], device=dev)
    z = torch.tensor(g["z"], device=dev, requires_grad=True)
    perms = [None] + [torch.tensor(g["perms"][i], device=dev) for i in range(ns - 1)]
    loss = vv.quartet_loss(x, z, str(g["hd"]), str(g["ld"]), perms)
    (loss * 3.0).backward()  # upstream gradient != 1 (lam_mds multiplies outside the op, network.py:286)
    grad = z.grad.cpu().numpy() / 3.0
    # vs the float32 reference (the 1e-4 bar) ...
    np.testing.assert_allclose(loss.item(), float(g["loss_f32"]), rtol=1e-4)
    scale = np.abs(g["grad_f32"]).max()
    np.testing.assert_allclose(grad, g["grad_f32"], rtol=1e-4, atol=1e-4 * scale)
    # ... and vs the float64 oracle, live
    ol, og = orc.quartet_loss(g["x"], g["z"], str(g["hd"]), str(g["ld"]), [None] + list(g["perms"]))
    np.testing.assert_allclose(loss.item(), ol, rtol=1e-4)
    np.testing.assert_allclose(grad, og, rtol=1e-4, atol=1e-4 * scale)


@pytest.mark.parametrize("big", ["0", "1"])
def test_quartet_small_batches_nan_and_zero_grad_rows(monkeypatch, big):
    monkeypatch.setenv("VVB_QUARTET_BIG", big)
    dev = torch.device("cuda")
    x = torch.randn(3, 5, device=dev)
    z = torch.randn(3, 2, device=dev, requires_grad=True)
    loss = vv.MDSLoss()(x, z)
    assert torch.isnan(loss)  # B < 4: mean over zero quartets, as in the reference
    loss.backward()
    assert torch.all(z.grad == 0)
    x = torch.randn(10, 5, device=dev)
    z = torch.randn(10, 2, device=dev, requires_grad=True)
    vv.MDSLoss()(x, z).backward()
    assert torch.all(z.grad[8:] == 0) and torch.any(z.grad[:8] != 0)
    with pytest.raises(ValueError, match="Invalid distance function"):
        vv.MDSLoss()(x, z, distf_hd="manhattan")


@pytest.mark.parametrize("big", ["0", "1"])
def test_quartet_is_deterministic_and_resampling_uses_global_rng(monkeypatch, big):
    monkeypatch.setenv("VVB_QUARTET_BIG", big)
    dev = torch.device("cuda")
    x = torch.randn(1024, 50, device=dev)
    z0 = torch.randn(1024, 2, device=dev)
    outs = []
    for _ in range(3):
        torch.manual_seed(5)
        z = z0.clone().requires_grad_(True)
        loss = vv.MDSLoss()(x, z, n_sampling=4)
        loss.backward()
        outs.append((loss.item(), z.grad.clone()))
    assert outs[0][0] == outs[1][0] == outs[2][0]
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][1], outs[2][1])
    # same draws as the reference: pass i>0 consumes torch.randperm(B) from the global generator
    torch.manual_seed(5)
    perms = [None] + [torch.randperm(1024, device=dev) for _ in range(3)]
    z = z0.clone().requires_grad_(True)
    assert vv.quartet_loss(x, z, perms=perms).item() == outs[0][0]


@pytest.mark.parametrize("B,D,L,big", [(1 << 16, 50, 2, "0"), (1 << 16, 50, 2, "1"), ((1 << 18) + 6, 37, 3, None),
                                       ((1 << 20) + 3, 50, 2, None)])
def test_quartet_large_batch_against_oracle(monkeypatch, B, D, L, big):
    """big=None: the launcher's own choice (eight lanes per quartet at 2^18 rows, 32 quartets per warp at 2^20); odd D
    (no 8-byte loads), L = 3, B mod 4 != 0 and a last round with idle lanes included."""
    if big is None:
        monkeypatch.delenv("VVB_QUARTET_BIG", raising=False)
    else:
        monkeypatch.setenv("VVB_QUARTET_BIG", big)
    rng = np.random.default_rng(3)
    x = clustered(B, D, 10, seed=6)
    z = rng.standard_normal((B, L)).astype(np.float32)
    zt = torch.tensor(z, device="cuda", requires_grad=True)
    loss = vv.quartet_loss(torch.tensor(x, device="cuda"), zt)
    loss.backward()
    ol, og = orc.quartet_loss(x, z)
    np.testing.assert_allclose(loss.item(), ol, rtol=1e-4)
    np.testing.assert_allclose(zt.grad.cpu().numpy(), og, rtol=1e-4, atol=1e-4 * np.abs(og).max())


# ============================================================================ fit
def test_reference_smoke_test_on_gpu():
    """/root/reference/tests/tests.py:5-15, verbatim settings (batch_size=2 -> the quartet term
    is NaN in the log exactly as in the reference, weights stay finite)."""
    np.random.seed(1)
    x = np.random.rand(50, 2).astype(np.float32)
    model = vv.ViVAE(input_dim=x.shape[1], latent_dim=2)
    buf = io.StringIO()
    with redirect_stdout(buf):
        model.fit(x, n_epochs=1, lam_recon=1.0, lam_kldiv=0.1, lam_geom=1.0, lam_egeom=1.0, lam_mds=1.0, batch_size=2)
    emb = model.transform(x)
    assert emb.shape[0] == x.shape[0] and emb.shape[1] == 2
    assert np.all(np.isfinite(emb)) and "mds: nan" in buf.getvalue()


def test_fit_replay_matches_reference_run():
    """Replay the recorded reference run (tests/golden/fit_a.npz: initial weights, loader
    permutations and reparameterisation noise) through the product's training loop on the GPU."""
    g = np.load(os.path.join(GOLDEN, "fit_a.npz"))
    dev = torch.device("cuda")
    model = vv.ViVAE(input_dim=g["X"].shape[1], latent_dim=2, random_state=3)
    sd = {k[len("init__"):]: torch.tensor(g[k], device=dev) for k in g.files if k.startswith("init__")}
    model.net.load_state_dict(sd)
    loader_perms = [g["perms"][0], g["perms"][2]]  # every epoch draws two, the 2nd is unused (see DeviceBatchLoader)
    eps = torch.tensor(g["eps"], device=dev)
    cursor = {"perm": 0, "eps": 0}

    def next_perm():
        p = torch.tensor(loader_perms[cursor["perm"]], device=dev)
        cursor["perm"] += 1
        return p

    real_randn_like = torch.randn_like

    def replay_randn_like(t, *a, **k):
        e = eps[cursor["eps"]:cursor["eps"] + t.shape[0]].reshape(t.shape)
        cursor["eps"] += t.shape[0]
        return e

    # first call builds the loader; install the replay hooks by running zero epochs first
    model.use_cuda_graph = False  # the replay hook below swaps torch.randn_like per call: eager only
    model.fit(g["X"], n_epochs=0, batch_size=int(g["batch_size"]), lam_mds=float(g["lam_mds"]), verbose=False)
    model.data_loader._perm_override = next_perm
    torch.randn_like = replay_randn_like
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            for epoch in (1, 2):
                losses = model.train_epoch(lam_mds=float(g["lam_mds"]))
                print(f"{losses['recon']/epoch:.4f} {losses['kldiv']/epoch:.4f} {losses['mds']/epoch:.4f}")
    finally:
        torch.randn_like = real_randn_like
    assert cursor["eps"] == eps.shape[0]
    ref_lines = [ln.split("\t") for ln in str(g["log"]).strip().splitlines()]
    got = [[float(v) for v in ln.split()] for ln in buf.getvalue().strip().splitlines()]
    for ref, mine in zip(ref_lines, got):
        want = [float(ref[1].split()[1]), float(ref[2].split()[1]), float(ref[5].split()[1])]
        np.testing.assert_allclose(mine, want, rtol=2e-3, atol=2e-4)
    emb = model.transform(g["X"])
    # 10 optimiser steps in FP32 on different hardware: compare at 1e-2 of the embedding's spread
    assert np.abs(emb - g["emb"]).max() < 1e-2 * np.abs(g["emb"]).max()


def test_cuda_graph_step_matches_eager_training():
    """The captured step (forward + quartet kernel + backward + Adam) must train exactly like the
    eager loop: same seed -> same loss log and embedding (warm-up steps are undone before capture)."""
    x = clustered(6_000 + 37, 20, 6, seed=13)  # 37 tail rows: the short last batch runs eagerly
    runs = []
    for use_graph in (False, True):
        m = vv.ViVAE(input_dim=20, random_state=11)
        m.use_cuda_graph = use_graph
        buf = io.StringIO()
        with redirect_stdout(buf):
            m.fit(x, n_epochs=3, batch_size=512, lam_mds=10.0)
        runs.append((buf.getvalue(), m.transform(x), m._graph))
    assert runs[0][2] is None and runs[1][2] is not None and runs[1][2][1] is not None, "graph path not taken"
    a = [[float(t.split()[1]) for t in ln.split("\t")[1:]] for ln in runs[0][0].strip().splitlines()]
    b = [[float(t.split()[1]) for t in ln.split("\t")[1:]] for ln in runs[1][0].strip().splitlines()]
    np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4)
    assert np.abs(runs[0][1] - runs[1][1]).max() < 2e-3 * np.abs(runs[0][1]).max()


def test_geometric_terms_closed_form_on_gpu_and_inside_the_captured_step():
    """the geometric losses with closed-form Jacobians equal the autodiff route on the
    device, and a training run with both terms switched on replays as a CUDA graph like the eager loop."""
    from vivae_b200.losses import EncoderGeometricLoss, GeometricLoss, _batched_jacobian, _logdet_variance

    torch.manual_seed(5)
    m = vv.ViVAE(input_dim=20, random_state=2)
    xb = torch.randn(256, 20, device="cuda")
    zb = torch.randn(256, 2, device="cuda")
    for f, inp, enc, loss in ((m.net.immersion, zb, False, GeometricLoss()), (m.net.submersion, xb, True, EncoderGeometricLoss())):
        a = loss(f, inp)
        b = _logdet_variance(_batched_jacobian(f, inp), encoder_side=enc)
        assert abs(a.item() - b.item()) <= 2e-4 * max(1.0, abs(b.item()))
    x = clustered(4_096, 20, 6, seed=17)
    runs = []
    for use_graph in (False, True):
        mm = vv.ViVAE(input_dim=20, random_state=11)
        mm.use_cuda_graph = use_graph
        buf = io.StringIO()
        with redirect_stdout(buf):
            mm.fit(x, n_epochs=2, batch_size=512, lam_mds=10.0, lam_geom=1.0, lam_egeom=1.0)
        runs.append((buf.getvalue(), mm.transform(x), mm._graph))
    assert runs[1][2] is not None and runs[1][2][1] is not None, "graph path not taken with geometric terms"
    a = [[float(t.split()[1]) for t in ln.split("\t")[1:]] for ln in runs[0][0].strip().splitlines()]
    b = [[float(t.split()[1]) for t in ln.split("\t")[1:]] for ln in runs[1][0].strip().splitlines()]
    assert all(v[2] > 0 and v[3] > 0 for v in a), "geometric terms missing from the log"
    np.testing.assert_allclose(a, b, rtol=2e-3, atol=2e-4)
    assert np.abs(runs[0][1] - runs[1][1]).max() < 5e-3 * np.abs(runs[0][1]).max()


def test_full_pipeline_through_public_api():
    x = clustered(20_000, 38, 12, seed=10)
    knn = vv.make_knn(x, k=50, verbose=False)
    xd = vv.smooth(x, knn, k=50, coef=1.0, n_iter=1)
    model = vv.ViVAE(input_dim=38, random_state=1)
    before = vv._lib.launch_count()
    model.fit(xd, n_epochs=2, batch_size=512, lam_mds=10.0, verbose=False)
    assert vv._lib.launch_count() - before >= 2 * 2 * 40  # fused quartet fwd + bwd kernel per batch
    emb = model.transform(xd)
    assert emb.shape == (20_000, 2) and np.all(np.isfinite(emb))
    assert model.trained and model.decoder_active


# ===================================================== tensor-core screening internals
def _screen_debug(x, k, q0, nq):
    import ctypes

    lib = vv._lib.load()
    cap = lib.vvb_knn_screen_debug_capacity()
    n, d = x.shape
    ci = np.zeros((nq, cap), dtype=np.uint32)
    ck = np.zeros((nq, cap), dtype=np.float32)
    cc = np.zeros(nq, dtype=np.int32)
    ct = np.zeros(nq, dtype=np.float32)
    xn = np.zeros(nq, dtype=np.float32)
    sy = np.zeros(2, dtype=np.float32)
    mu = np.zeros(64, dtype=np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    vv._lib.check(lib.vvb_knn_screen_debug_host(p(x), n, d, k, q0, nq, p(ci), p(ck), p(cc), p(ct), p(xn), p(sy), p(mu)))
    return ci, ck, cc, ct, xn, float(sy[0]), float(sy[1]), mu[:d]


@pytest.mark.parametrize("kind", ["clustered", "offset"])
def test_screening_error_is_far_below_the_certificate_margin(kind):
    """The certificate assumes |key + |x^|^2 - s^2 d2| <= eps_rel(ka) (|x^| + Ymax)^2 with eps_rel(1) = 2^-18 + 6 * 2^-20
    = 2^-16.7 (knn_tensor.cu: tc_eps_rel).  Measure the actual tensor-core error against FP64 on the kept candidates:
    it must stay 8x below (tests/test_gpu_configs.py repeats this at d = 500 and 2000)."""
    n, d, k = 20_000, 50, 50
    x = clustered(n, d, 12, seed=21)
    if kind == "offset":
        x = (x + 300.0).astype(np.float32)
    nq = 1024
    ci, ck, cc, ct, xn, s, ymax, mu = _screen_debug(x, k, 500, nq)
    m = k + 1 + max(16, k // 4)
    if m % 32 <= 4 and (m & ~31) - 4 - (k + 1) >= 8:  # tc_candidates(): stay inside a re-rank group of 32
        m = (m & ~31) - 4
    assert m == 60
    # the kept list holds the m smallest keys and at most 8 more (never past the group of 32), unsorted
    assert np.all(cc >= m) and np.all(cc <= min(m + 8, (m + 31) & ~31)) and s > 0 and 16.0 < ymax <= 64.1
    for t in range(0, nq, 13):
        assert np.all(ck[t, :cc[t]] <= ct[t])
    xc = (x.astype(np.float64) - mu.astype(np.float64)) * s
    worst = 0.0
    for t in range(0, nq, 7):
        g = 500 + t
        j = ci[t, :cc[t]].astype(np.int64)
        d2 = ((xc[g][None, :] - xc[j]) ** 2).sum(1)
        err = np.abs(ck[t, :cc[t]].astype(np.float64) + float(xn[t]) - d2)
        bound = (np.sqrt(float(xn[t])) + ymax) ** 2
        worst = max(worst, float((err / bound).max()))
    print(f"screening error / (|x|+Ymax)^2 = 2^{np.log2(max(worst, 1e-30)):.1f}")
    assert worst < (2.0 ** -18 + 6 * 2.0 ** -20) / 8.0
    # every true neighbour is among the candidates, self included
    oi, _ = orc.knn(x, k, 500, 500 + nq)
    for t in range(0, nq, 11):
        assert set(oi[t].tolist()) <= set(ci[t, :cc[t]].astype(np.int64).tolist())


def test_tensor_path_is_used_and_fallback_is_rare():
    x = clustered(50_000, 50, 30, seed=1)
    st = {}
    idx, dist = vv.make_knn(x, k=50, verbose=False, algo="tensor", stats=st)
    assert st["algo_used"] == 2 and st["candidates"] == 60  # tc_candidates(50)
    assert st["rows_fallback"] <= 0.01 * st["rows"], st
    for q0 in (0, 25_000, 49_000):
        wi, wd = orc.knn(x, 50, q0, q0 + 500)
        assert np.array_equal(idx[q0:q0 + 500], wi) and np.array_equal(dist[q0:q0 + 500], wd)
    with pytest.raises(NotImplementedError):
        vv.make_knn(clustered(300, 4200, 3, seed=2), k=10, verbose=False, algo="tensor")  # d + 3 > 4096: brute only
    with pytest.raises(NotImplementedError):
        vv.make_knn(clustered(2000, 20, 3, seed=2), k=400, verbose=False, algo="tensor")  # k too large: brute only


@pytest.mark.parametrize("n,d,k", [(6000, 100, 50), (4000, 500, 30), (3000, 2000, 50), (2500, 62, 100), (1000, 129, 5)])
def test_knn_wide_rows_use_the_k_loop_tensor_kernel(n, d, k):
    """d + 3 > 64: the K-loop variant (query and database atoms streamed together); PCA-100 and the
    2000-gene dense configuration of BASELINE.json are in this regime."""
    x = clustered(n, d, 7, seed=d)
    st = {}
    got = vv.make_knn(x, k=k, verbose=False, algo="tensor", stats=st)
    assert st["algo_used"] == 2
    assert_knn_equal(got, orc.knn(x, k))
    assert st["rows_fallback"] <= 0.02 * n, st


def test_certificate_rejects_adversarial_rows_and_fallback_repairs_them():
    """Points on a fine grid far from the data's bulk: thousands of near-ties within the screening
    error of each other.  Whatever the certificate decides, the answer must be exact."""
    rng = np.random.default_rng(4)
    bulk = rng.standard_normal((3000, 8)).astype(np.float32) * 50.0
    grid = (rng.integers(0, 3, size=(3000, 8)).astype(np.float32) * 1e-3) + 40.0
    x = np.concatenate([bulk, grid]).astype(np.float32)
    st = {}
    got = vv.make_knn(x, k=20, verbose=False, algo="tensor", stats=st)
    assert_knn_equal(got, orc.knn(x, 20))
    print("adversarial fallback rows:", st["rows_fallback"], "of", st["rows"])


def test_smooth_rejects_out_of_range_indices_without_touching_bad_memory():
    x = clustered(1000, 8, 3, seed=5)
    idx, dist = vv.make_knn(x, k=10, verbose=False)
    bad = idx.copy()
    bad[500, 3] = 10_000_000
    with pytest.raises(ValueError, match="must include self"):
        vv.smooth(x, [bad, dist], k=10, coef=1.0)
    bad[500, 3] = -7
    with pytest.raises(ValueError, match="must include self"):
        vv.smooth(x, [bad, dist], k=10, coef=1.0)
    assert np.array_equal(vv.smooth(x, [idx, dist], k=10, coef=1.0), orc.smooth(x, idx, 10, 1.0, 1))  # still healthy


def test_make_knn_host_chunked_path_matches_oracle_samples():
    """n above one 262,144-row chunk: exercises the chunk loop, the overlapped D2H and the cached buffers."""
    n, k = 300_000, 20
    x = clustered(n, 16, 25, seed=33)
    st = {}
    idx, dist = vv.make_knn(x, k=k, verbose=False, stats=st)
    assert st["rows"] == n and st["algo_used"] == 2
    for q0 in (0, 262_144 - 100, 262_144, n - 300):
        wi, wd = orc.knn(x, k, q0, q0 + 200)
        assert np.array_equal(idx[q0:q0 + 200], wi) and np.array_equal(dist[q0:q0 + 200], wd)
    idx2, dist2 = vv.make_knn(x, k=k, verbose=False)  # second call reuses the cached device buffers
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)
    assert vv._lib.load().vvb_release_cached_memory() == 0


@pytest.mark.parametrize("refine", ["2", "0", "1"])
def test_few_row_fallback_slices_the_database(monkeypatch, refine):
    """A tight knot of 90 nearly coincident points inside 60k cells: their neighbour order is below the
    screening resolution, so exactly those rows fail the certificate.  VVB_TC_REFINE=2: the refine sweep settles them
    on the tensor cores; =0: they take the few-row brute-force fallback (one CTA per row and database slice + merge);
    =1 (default): whichever the cost estimate prefers.  Either way the answer must be the oracle's."""
    monkeypatch.setenv("VVB_TC_REFINE", refine)
    x = clustered(60_000, 50, 10, seed=17)
    rng = np.random.default_rng(1)
    knot = np.arange(1000, 1090)
    x[knot] = x[1000] + (rng.standard_normal((90, 50)) * 2e-6).astype(np.float32)
    st = {}
    idx, dist = vv.make_knn(x, k=50, verbose=False, algo="tensor", stats=st)
    if refine == "2":
        assert 1 <= st["rows_refined"] <= 2048 and st["rows_fallback"] == 0, st
    elif refine == "0":
        assert 1 <= st["rows_fallback"] <= 2048 and st["rows_refined"] == 0, st
    else:
        assert 1 <= st["rows_refined"] + st["rows_fallback"] <= 2048, st
    wi, wd = orc.knn(x, 50, 900, 1200)
    assert np.array_equal(idx[900:1200], wi) and np.array_equal(dist[900:1200], wd)
    wi, wd = orc.knn(x, 50, 30_000, 30_200)
    assert np.array_equal(idx[30_000:30_200], wi) and np.array_equal(dist[30_000:30_200], wd)
    print("refine", refine, "rows refined / brute:", st["rows_refined"], st["rows_fallback"], f"{st['ms_fallback']:.2f} ms")


# ===================================================== cell-ordered sweep (exact pruning)
@pytest.mark.parametrize("cells", [1, 7, 64])
@pytest.mark.parametrize("kind,n,d,k", [("clustered", 20_000, 50, 50), ("iid", 6000, 20, 30), ("ties", 5000, 6, 20),
                                        ("clustered", 5000, 100, 20), ("line", 8000, 3, 10)])
def test_cell_sweep_is_exact_for_any_partition(monkeypatch, cells, kind, n, d, k):
    """The cell bounds only decide which database tiles a query block skips; whatever the partition
    (forced here through VVB_TC_CELLS, including cell counts that are not powers of two and cells far
    smaller than a tile) the result must be the brute-force contract bit for bit."""
    monkeypatch.setenv("VVB_TC_CELLS", str(cells))
    if kind == "clustered":
        x = clustered(n, d, 12, seed=cells)
    elif kind == "iid":
        x = np.random.default_rng(cells).standard_normal((n, d)).astype(np.float32)
    elif kind == "ties":
        x = tie_data(n, d, seed=cells)
    else:  # points along a line, sorted: long thin cells, many empty pivots are impossible but tiny ones are not
        t = np.sort(np.random.default_rng(cells).uniform(0, 1000, n)).astype(np.float32)
        x = np.stack([t, 0.5 * t, np.zeros_like(t)], axis=1).astype(np.float32)
    st = {}
    got = vv.make_knn(x, k=k, verbose=False, algo="tensor", stats=st)
    assert st["cells"] == cells
    assert_knn_equal(got, orc.knn(x, k))
    if kind in ("clustered", "iid"):
        assert st["rows_fallback"] <= 0.02 * n, st


def test_cell_sweep_skips_far_clusters_and_stays_exact(monkeypatch):
    """Well-separated clusters: most database tiles are skipped, the graph is still exact, and a
    query-row shard (rows of an arbitrary range, not a whole cell) sees the same answers."""
    monkeypatch.setenv("VVB_TC_CELLS", "128")
    x = clustered(160_000, 50, 20, seed=5)
    st = {}
    idx, dist = vv.make_knn(x, k=50, verbose=False, algo="tensor", stats=st)
    assert st["tiles_swept"] < 0.5 * st["tiles_dense"], st  # includes the threshold pre-pass (256 tiles per block)
    for q0 in (0, 31_000, 159_000):
        wi, wd = orc.knn(x, 50, q0, q0 + 400)
        assert np.array_equal(idx[q0:q0 + 400], wi) and np.array_equal(dist[q0:q0 + 400], wd)
    print("tiles swept / dense:", st["tiles_swept"], st["tiles_dense"], "fallback rows:", st["rows_fallback"])
    # a query-row shard that starts and ends in the middle of cells (C ABI: q0, nq)
    from vivae_b200.knn import _knn_host

    q0, nq = 77_777, 3_001
    si, sd = _knn_host(x, 50, q0, nq, "tensor", None)
    assert np.array_equal(si, idx[q0:q0 + nq]) and np.array_equal(sd, dist[q0:q0 + nq])


# ============================================================================ database-sharded (ring) build
@pytest.mark.parametrize("kind,n,d,k,world", [("clustered", 30_000, 50, 50, 3), ("ties", 6_000, 6, 20, 4),
                                               ("clustered", 5_000, 130, 100, 2), ("iid", 700, 20, 300, 3)])
def test_ring_steps_with_merge_kernel_equal_whole_matrix_build(kind, n, d, k, world):
    """The ring build's sweeps (builder on [shards ; own rows] in global row order + csrc/knn_merge.cu per contiguous
    range), run here for every 'rank' in one process and for one / two / all foreign shards per sweep, are
    bit-identical to one build over the whole matrix -- including tie-breaks across shard borders and k larger than
    a shard."""
    from vivae_b200.sharded import _DeviceOps, ring_fold, shard_bounds

    x = {"clustered": lambda: clustered(n, d, 8, seed=21), "ties": lambda: tie_data(n, d, seed=22),
         "iid": lambda: np.random.default_rng(23).standard_normal((n, d)).astype(np.float32)}[kind]()
    xd = torch.tensor(x, device="cuda")
    ops = _DeviceOps()
    whole_idx, whole_dist = ops.knn(xd, k)
    before = vv._lib.launch_count()
    for rank in range(world):
        q0, nq = shard_bounds(n, world, rank)
        own = xd[q0:q0 + nq].contiguous()
        foreign = []
        for s in range(1, world):
            s0, sn = shard_bounds(n, world, (rank + s) % world)
            foreign.append((s0, xd[s0:s0 + sn].contiguous()))
        for per_sweep in sorted({1, 2, world - 1}):  # shards per sweep: the schedules make_knn_ring can choose
            run_idx, run_d2 = ops.init(nq, k, q0, xd.device)
            for g0 in range(0, len(foreign), per_sweep):
                ring_fold(ops, own, q0, foreign[g0:g0 + per_sweep], run_idx, run_d2, with_own=(g0 == 0))
            dist = ops.finish(run_d2)
            assert torch.equal(run_idx, whole_idx[q0:q0 + nq]), (rank, per_sweep)
            assert torch.equal(dist, whole_dist[q0:q0 + nq]), (rank, per_sweep)
    assert vv._lib.launch_count() > before

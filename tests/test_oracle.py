"""CPU tests: the oracle (oracle/) against the golden vectors recorded from the
unmodified reference (tests/golden/, made by oracle/make_golden.py) and against
independent restatements."""

import glob
import os

import numpy as np
import pytest

import oracle as orc

from conftest import GOLDEN


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


SMOOTH = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "smooth_*.npz")))
QUARTET = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "quartet_*.npz")))


def test_fixtures_present():
    assert len(SMOOTH) >= 6 and len(QUARTET) >= 7


@pytest.mark.parametrize("name", SMOOTH)
def test_smooth_oracle_bit_exact_vs_reference(name):
    g = _load(name)
    out = orc.smooth(g["x"], g["idx"].astype(np.int64), int(g["k"]), float(g["coef"]), int(g["n_iter"]))
    assert out.dtype == g["out"].dtype
    assert np.array_equal(out, g["out"]), f"max abs diff {np.abs(out - g['out']).max()}"


@pytest.mark.parametrize("name", ["smooth_a.npz", "smooth_b.npz", "smooth_c.npz", "smooth_e.npz"])
def test_smooth_numpy_restatement_bit_exact(name):
    g = _load(name)
    out = orc.smooth_numpy(g["x"], g["idx"].astype(np.int64), int(g["k"]), float(g["coef"]), int(g["n_iter"]))
    assert np.array_equal(out, g["out"])


def test_smooth_jacobi_differs_from_reference():
    # a Jacobi gather-mean is NOT the reference's semantics
    g = _load("smooth_a.npz")
    jac = orc.smooth(g["x"], g["idx"].astype(np.int64), int(g["k"]), float(g["coef"]), int(g["n_iter"]), mode=1)
    assert not np.allclose(jac, g["out"], rtol=1e-5, atol=0)
    jac_np = orc.smooth_numpy(g["x"], g["idx"].astype(np.int64), int(g["k"]), float(g["coef"]), 1, mode=1)
    assert np.array_equal(jac, jac_np)


@pytest.mark.parametrize("name", QUARTET)
def test_quartet_oracle_vs_reference(name):
    g = _load(name)
    ns = int(g["n_sampling"])
    perms = [None] + [g["perms"][i] for i in range(ns - 1)]
    assert len(g["perms"]) == ns - 1
    loss, grad = orc.quartet_loss(g["x"], g["z"], str(g["hd"]), str(g["ld"]), perms)
    # float64 run of the reference: gradients are float64, the loss scalar is
    # accumulated into a float32 tensor (losses.py:321,345) -> 1e-6; the 1/n_sampling
    # factor is therefore a float32 constant in the backward pass -> 1e-7 on gradients
    np.testing.assert_allclose(grad, g["grad_f64"], rtol=2e-7, atol=1e-12)
    np.testing.assert_allclose(loss, float(g["loss_f64"]), rtol=2e-6)
    # float32 run of the reference: north-star tolerance 1e-4 relative
    np.testing.assert_allclose(loss, float(g["loss_f32"]), rtol=1e-4)
    scale = np.abs(g["grad_f32"]).max()
    np.testing.assert_allclose(grad, g["grad_f32"], rtol=1e-4, atol=1e-4 * scale)
    # value-only NumPy restatement
    np.testing.assert_allclose(orc.quartet_loss_numpy(g["x"], g["z"], str(g["hd"]), str(g["ld"]), perms), loss, rtol=1e-12)


def test_quartet_unused_rows_have_zero_grad_and_small_batches_are_nan():
    g = _load("quartet_f.npz")  # B=7 -> one quartet (rows 0..3), rows 4..6 unused
    _, grad = orc.quartet_loss(g["x"], g["z"])
    assert np.all(grad[4:] == 0.0) and np.any(grad[:4] != 0.0)
    loss, _ = orc.quartet_loss(g["x"][:3], g["z"][:3])  # nq = 0 -> nan like the reference
    assert np.isnan(loss)


def test_quartet_grad_matches_finite_differences():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((16, 6))
    z = rng.standard_normal((16, 3))
    for hd, ld in (("euclidean", "euclidean"), ("cosine", "cosine")):
        _, grad = orc.quartet_loss(x, z, hd, ld)
        num = np.zeros_like(z)
        h = 1e-6
        for i in range(z.shape[0]):
            for c in range(z.shape[1]):
                zp, zm = z.copy(), z.copy()
                zp[i, c] += h
                zm[i, c] -= h
                num[i, c] = (orc.quartet_loss(x, zp, hd, ld, want_grad=False)[0]
                             - orc.quartet_loss(x, zm, hd, ld, want_grad=False)[0]) / (2 * h)
        np.testing.assert_allclose(grad, num, rtol=1e-5, atol=1e-8)


# ------------------------------------------------------------------------- kNN
def _tie_data(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 4, size=(n, d)).astype(np.float32)  # many exact ties
    x[n // 2] = x[1]                                          # exact duplicates
    x[n - 1] = x[1]
    return x


@pytest.mark.parametrize("n,d,k,seed", [(40, 3, 5, 0), (33, 2, 32, 1), (25, 7, 1, 2), (48, 5, 10, 3)])
def test_knn_oracle_vs_exact_rational_restatement(n, d, k, seed):
    x = _tie_data(n, d, seed) if seed % 2 == 0 else np.random.default_rng(seed).standard_normal((n, d)).astype(np.float32) * 3
    idx, dist = orc.knn(x, k)
    idx2, dist2 = orc.knn_tiny(x, k)
    assert np.array_equal(idx, idx2)
    assert np.array_equal(dist, dist2)
    assert np.array_equal(idx[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_knn_oracle_query_slices_and_threads():
    x = np.random.default_rng(5).standard_normal((700, 9)).astype(np.float32)
    idx, dist = orc.knn(x, 12)
    i2, d2 = orc.knn(x, 12, 100, 333)
    assert np.array_equal(idx[100:333], i2) and np.array_equal(dist[100:333], d2)
    assert orc.num_threads() >= 1


def test_knn_oracle_vs_kdtree():
    from scipy.spatial import cKDTree

    x = np.random.default_rng(6).standard_normal((3000, 10)).astype(np.float32)
    k = 20
    idx, dist = orc.knn(x, k)
    dd, ii = cKDTree(x.astype(np.float64)).query(x.astype(np.float64), k=k + 1)
    np.testing.assert_allclose(dist, dd, rtol=1e-5, atol=1e-6)
    # neighbour SETS agree except where FP32 vs FP64 rounding reorders near-ties
    same = np.mean([len(set(a) & set(b)) / (k + 1) for a, b in zip(idx, ii)])
    assert same > 0.999


def test_knn_oracle_argument_errors():
    x = np.zeros((5, 2), dtype=np.float32)
    with pytest.raises(ValueError):
        orc.knn(x, 5)  # k must be <= n-1  (knn.py:119-120)
    with pytest.raises(ValueError):
        orc.knn(x, 0)


# ---------------------------------------------------------------------- the `fit` reference arm (oracle/ref_fit.py)
@pytest.mark.parametrize("name", QUARTET)
def test_ref_fit_quartet_restatement_matches_reference_golden(name, monkeypatch):
    """oracle/ref_fit.py restates MDSLoss operation for operation (it is the timed reference arm of `fit` in
    bench.py): value and autograd gradient against the reference-recorded float32 fixtures, recorded
    permutations replayed through torch.randperm."""
    import torch

    import ref_fit

    g = _load(name)
    ns = int(g["n_sampling"])
    recorded = [torch.tensor(g["perms"][i]) for i in range(ns - 1)]
    monkeypatch.setattr(torch, "randperm", lambda n, **kw: recorded.pop(0))
    x = torch.tensor(g["x"])
    z = torch.tensor(g["z"], requires_grad=True)
    loss = ref_fit.ref_mds_loss(x, z, str(g["hd"]), str(g["ld"]), ns)
    loss.backward()
    assert abs(float(loss) - float(g["loss_f32"])) <= 2e-6 * abs(float(g["loss_f32"]))
    np.testing.assert_allclose(z.grad.numpy(), g["grad_f32"], rtol=1e-4, atol=1e-6 * np.abs(g["grad_f32"]).max())


def test_ref_fit_autoencoder_terms_equal_the_pinned_network_on_cpu():
    """RefAutoencoder (reference layer names) loaded with the weights of `vivae_b200.network.Autoencoder` -- which is
    pinned bit-identically against the reference (fit_cpu_nomds.npz) -- gives the same KL and MSE terms."""
    import torch

    import ref_fit
    from vivae_b200.network import Autoencoder

    torch.manual_seed(3)
    ours = Autoencoder(input_dim=9, latent_dim=2, hidden_dims=[32, 64, 128, 32], variational=True)
    ref = ref_fit.RefAutoencoder(9)
    ref.load_state_dict({k: v for k, v in ours.state_dict().items()})
    x = torch.randn(64, 9)
    torch.manual_seed(11)
    a = ours(x, lam_recon=1.0, lam_kldiv=0.5, lam_mds=0.0)
    torch.manual_seed(11)
    b = ref(x, lam_recon=1.0, lam_kldiv=0.5, lam_mds=0.0)
    assert torch.equal(a["kldiv"], b["kldiv"]) and torch.equal(a["recon"], b["recon"])
    t = ref_fit.time_ref_fit(np.random.default_rng(0).standard_normal((600, 9)).astype(np.float32), batch_size=64,
                             warm_batches=2, lam_mds=10.0)
    assert t["s_per_epoch"] > 0 and t["rows"] == 600

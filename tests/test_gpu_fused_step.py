"""GPU tests of the fused training step (csrc/vae_step.cu: forward + three loss terms + backward in one kernel)
against the PyTorch path it replaces (`Autoencoder.forward` + autograd), through the C ABI."""

import numpy as np
import pytest
import torch

import vivae_b200 as vv
from vivae_b200.fused_step import FusedStep, fused_step_supported
from vivae_b200.network import Autoencoder

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,hidden,L,B", [(50, [32, 64, 128, 32], 2, 512), (38, [32, 64, 128, 32], 2, 256),
                                          (9, [16, 8], 3, 64), (100, [64, 32], 2, 36), (12, [128], 1, 4)])
def test_fused_step_gradients_and_terms_equal_autograd(D, hidden, L, B):
    """Same weights, same batch, same noise (seeded default generator): every parameter gradient and the three weighted
    loss terms of the fused kernel against `Autoencoder.forward` + `backward()`.  B = 36 leaves the second CTA with
    one quartet of eight (masking), B = 4 is a single quartet."""
    dev = torch.device("cuda")
    torch.manual_seed(3)
    net = Autoencoder(input_dim=D, latent_dim=L, hidden_dims=hidden, variational=True).to(dev)
    x = torch.randn(B, D, device=dev) * 2.0
    lam = dict(lam_recon=1.0, lam_kldiv=0.7, lam_mds=10.0)
    assert fused_step_supported(net, lam, x)
    torch.manual_seed(11)
    terms = net(x, **lam)
    (terms["recon"] + terms["kldiv"] + terms["mds"]).backward()
    want = [p.grad.detach().clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    fs = FusedStep(net, B)
    sums = torch.zeros(6, device=dev)
    torch.manual_seed(11)
    out = fs.step(x, sums, **lam)
    torch.cuda.synchronize()
    got_terms = out.cpu().numpy()
    np.testing.assert_allclose(got_terms, [terms["recon"].item(), terms["kldiv"].item(), terms["mds"].item()], rtol=2e-5)
    np.testing.assert_allclose(sums.cpu().numpy()[[0, 1, 4]], got_terms, rtol=1e-6)
    for p, w in zip(net.parameters(), want):
        g = p.grad.detach()
        scale = float(w.abs().max()) + 1e-12
        assert float((g - w).abs().max()) <= 2e-4 * scale, (tuple(p.shape), float((g - w).abs().max()), scale)


def test_fused_step_is_deterministic_and_leaves_unsupported_configurations_to_pytorch():
    dev = torch.device("cuda")
    torch.manual_seed(1)
    net = Autoencoder(input_dim=20, latent_dim=2, hidden_dims=[32, 16], variational=True).to(dev)
    x = torch.randn(128, 20, device=dev)
    fs = FusedStep(net, 128)
    runs = []
    for _ in range(2):
        torch.manual_seed(4)
        fs.step(x, None, 1.0, 1.0, 100.0)
        runs.append(fs.grad_flat.clone())
    assert torch.equal(runs[0], runs[1])  # no atomics anywhere: bit-identical run to run
    cfg = dict(lam_recon=1.0, lam_kldiv=1.0, lam_mds=100.0)
    assert fused_step_supported(net, cfg, x)
    assert not fused_step_supported(net, dict(cfg, lam_geom=1.0), x)
    assert not fused_step_supported(net, dict(cfg, mds_distf_hd="cosine"), x)
    assert not fused_step_supported(net, dict(cfg, mds_nsamp=2), x)
    assert not fused_step_supported(net, dict(cfg, lam_recon=0.0), x)
    assert not fused_step_supported(net, cfg, x[:126])
    tanh = Autoencoder(input_dim=20, latent_dim=2, hidden_dims=[32, 16], variational=True, activation=torch.nn.Tanh()).to(dev)
    assert not fused_step_supported(tanh, cfg, x)
    plain = Autoencoder(input_dim=20, latent_dim=2, hidden_dims=[32, 16], variational=False).to(dev)
    assert not fused_step_supported(plain, cfg, x)


def test_fit_with_the_fused_step_tracks_the_pytorch_path(monkeypatch):
    """Two epochs of `ViVAE.fit` (captured CUDA graph in both cases) with the fused step and with the PyTorch step:
    same seed, same batches, same noise -- the per-epoch loss sums stay within 1e-3 relative of each other and the
    embeddings agree closely (rounding differs between the hand-written layers and cuBLAS, nothing else)."""
    rng = np.random.default_rng(0)
    centres = rng.standard_normal((6, 30)) * 3
    x = (centres[rng.integers(0, 6, 20_000)] + rng.standard_normal((20_000, 30))).astype(np.float32)
    res = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("VIVAE_FUSED_STEP", fused)
        m = vv.ViVAE(input_dim=30, random_state=7)
        m.fit(x, n_epochs=1, batch_size=512, lam_mds=10.0, verbose=False)
        before = vv._lib.launch_count()
        losses = m.train_epoch(lam_mds=10.0)
        launches = vv._lib.launch_count() - before
        res[fused] = (losses, m.transform(x[:2000]), launches)
    for name in ("recon", "kldiv", "mds"):
        a, b = res["1"][0][name], res["0"][0][name]
        assert abs(a - b) <= 2e-3 * abs(b), (name, a, b)
    assert np.abs(res["1"][1] - res["0"][1]).max() < 0.05 * np.abs(res["0"][1]).max()
    assert res["1"][2] >= 2 * (20_000 // 512) and res["0"][2] >= 2 * (20_000 // 512)  # this library's kernels ran

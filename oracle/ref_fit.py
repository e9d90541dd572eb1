"""TEST / BENCH INFRASTRUCTURE ONLY -- restatement of the reference's `ViVAE.fit` step in plain PyTorch.

Purpose: the "reference arm" of the `fit` half of BASELINE.json's metric (`ViVAE.fit` s/epoch), timed on the GPU box's
own host cores beside our number (bench.py), and the eager-GPU variant that shows the dispatch-bound baseline.  The
unmodified reference cannot travel to the GPU box (`/root/reference` does not exist there and the package needs
pynndescent / matplotlib at import), so this file restates -- operation for operation, because the cost of the
reference IS its operation count -- what one training epoch executes:

  * data loader            /root/reference/src/vivae/model.py:246-254  (DataLoader(TensorDataset(X)), shuffle=True)
  * epoch loop             model.py:130-177   (zero_grad, forward, three `.item()` host syncs per batch, backward, step)
  * Autoencoder.forward    network.py:219-296 (encode -> KL -> decode -> MSE -> quartet loss)
  * MLP                    network.py:62-99   (Linear + exact-erf GELU; mu / logvar heads; mirrored decoder)
  * reparameterise         network.py:113-130
  * KLDivLoss              losses.py:105-131
  * MDSLoss                losses.py:140-349  (the x[perm][arange] gathers are executed TWICE per pass, as in
                                               losses.py:332-336; six pair distances, two normalisers, six costs)
  * Adam(lr 1e-3, wd 1e-4) model.py:256

Pinned: `tests/test_oracle.py::test_ref_fit_restatement_matches_reference_golden` replays the reference-recorded
fixture `tests/golden/fit_cpu_nomds.npz` / quartet fixtures through these functions.
Nothing under `vivae_b200/` imports this module."""

from __future__ import annotations

import time
from collections import OrderedDict

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader, TensorDataset


def _euclid(a, b, dev):  # losses.py:140-160
    return torch.sqrt(torch.maximum(torch.sum(torch.square(a - b), dim=1, keepdim=False), torch.tensor(1e-9, device=dev)))


def _cosine(a, b, dev):  # losses.py:162-177
    return 1.0 - torch.nn.functional.cosine_similarity(a, b, dim=1)


_PAIRS = ((0, 1), (1, 2), (2, 3), (0, 2), (0, 3), (1, 3))  # losses.py:234, 263


def ref_mds_loss(x, z, distf_hd="euclidean", distf_ld="euclidean", n_sampling=1):
    """losses.py:266-349, same operations in the same order (including the duplicated gathers)."""
    dev = x.device
    fh = {"euclidean": _euclid, "cosine": _cosine}[distf_hd]
    fl = {"euclidean": _euclid, "cosine": _cosine}[distf_ld]
    n = x.shape[0]
    nq = n // 4
    loss = torch.tensor([0.0], device=dev)
    for i in range(n_sampling):
        starts = [0, nq, nq * 2, nq * 3]
        perm = torch.randperm(n, device=dev) if i > 0 else torch.arange(n, device=dev)
        for _ in range(2):  # losses.py:332-336 builds both lists twice
            xq = [x[perm][torch.arange(s, s + nq, device=dev)] for s in starts]
            zq = [z[perm][torch.arange(s, s + nq, device=dev)] for s in starts]
        xd = sum(fh(xq[a], xq[b], dev) for a, b in _PAIRS)  # losses.py:216-235
        zd = sum(fl(zq[a], zq[b], dev) for a, b in _PAIRS)
        cost = None
        for a, b in _PAIRS:  # losses.py:179-214
            term = torch.pow(fh(xq[a], xq[b], dev) / xd - fl(zq[a], zq[b], dev) / zd, torch.tensor(2.0, device=dev))
            cost = term if cost is None else cost + term
        res = torch.multiply(torch.mean(cost), torch.tensor(n, device=dev))
        loss += torch.divide(res, torch.tensor(nq, device=dev))
    return torch.mean(torch.divide(loss, n_sampling))


class RefAutoencoder(nn.Module):
    """network.py:62-164 with the reference's layer names (so a reference state_dict loads)."""

    def __init__(self, input_dim, latent_dim=2, hidden_dims=(32, 64, 128, 32)):
        super().__init__()
        act = nn.GELU()
        enc, prev = OrderedDict(), input_dim
        for i, h in enumerate(hidden_dims):
            enc[f"{i + 1:02d}EncoderPotential"] = nn.Linear(prev, h)
            enc[f"{i + 1:02d}EncoderActivation"] = act
            prev = h
        self.mu, self.logvar = nn.Linear(prev, latent_dim), nn.Linear(prev, latent_dim)
        dec, prev = OrderedDict(), latent_dim
        for i, h in enumerate(reversed(hidden_dims)):
            dec[f"{i + 1:02d}DecoderPotential"] = nn.Linear(prev, h)
            dec[f"{i + 1:02d}DecoderActivation"] = act
            prev = h
        dec[f"{len(hidden_dims) + 1:02d}DecoderPotential"] = nn.Linear(prev, input_dim)
        self.encoder, self.decoder = nn.Sequential(enc), nn.Sequential(dec)
        self.mse = nn.MSELoss()

    def forward(self, x, lam_recon=1.0, lam_kldiv=1.0, lam_mds=100.0, distf_hd="euclidean", distf_ld="euclidean", nsamp=1):
        h = self.encoder(x)
        mu, logvar = self.mu(h), self.logvar(h)
        z = mu + torch.randn_like(torch.exp(0.5 * logvar)) * torch.exp(0.5 * logvar)  # network.py:113-130
        eps_sq = torch.tensor([1e-2], device=x.device) ** 2
        kl = -0.5 * torch.mean(logvar + torch.log(eps_sq) - torch.square(mu) - eps_sq * logvar.exp())  # losses.py:127-130
        kl = lam_kldiv * (kl / mu.shape[0])
        out = {"kldiv": kl, "recon": None, "mds": None}
        if lam_recon > 0.0:
            out["recon"] = lam_recon * self.mse(x, self.decoder(z))
        if lam_mds > 0.0:
            out["mds"] = lam_mds * ref_mds_loss(x, z, distf_hd, distf_ld, nsamp)
        return out


def ref_train_epoch(net, loader, opt, **lam):
    """model.py:130-177: one pass over the loader with the reference's per-batch host syncs."""
    tot = {"recon": 0.0, "kldiv": 0.0, "mds": 0.0}
    for (xb,) in loader:
        opt.zero_grad()
        terms = net(xb, **lam)
        loss = 0.0
        for name in ("recon", "kldiv", "mds"):
            if terms[name] is not None:
                tot[name] += terms[name].item()
                loss = loss + terms[name]
        loss.backward()
        opt.step()
    return tot


def time_ref_fit(X: np.ndarray, batch_size=512, epochs=1, warm_batches=8, device="cpu", lam_mds=100.0, seed=1) -> dict:
    """Seconds per epoch of the restated reference `fit` on `device` ('cpu': all torch intra-op threads; 'cuda': the
    reference's eager GPU path).  One short warm-up (first `warm_batches` batches) is excluded."""
    dev = torch.device(device)
    torch.manual_seed(seed)
    # the reference makes its device the global default (src/vivae/__init__.py:17-27): the DataLoader's own
    # bookkeeping tensors (base seed, sampler permutation) then live there too, next to the generator
    with torch.device(dev):
        net = RefAutoencoder(X.shape[1]).to(dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        loader = DataLoader(TensorDataset(torch.tensor(X, device=dev)), batch_size=batch_size, shuffle=True, generator=g)
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
        net.train()
        it = iter(loader)
        for _ in range(warm_batches):
            (xb,) = next(it)
            opt.zero_grad()
            t = net(xb, lam_mds=lam_mds)
            (t["recon"] + t["kldiv"] + t["mds"]).backward()
            opt.step()
        del it
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(epochs):
            ref_train_epoch(net, loader, opt, lam_mds=lam_mds)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / epochs
    return {"rows": int(X.shape[0]), "batch_size": batch_size, "s_per_epoch": dt, "cells_per_s": X.shape[0] / dt,
            "device": device, "threads": torch.get_num_threads() if dev.type == "cpu" else None, "epochs_timed": epochs}

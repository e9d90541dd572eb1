"""Autoencoder network (encoder/decoder MLPs, loss composition, batched embedding).

Mirrors the interface of /root/reference/src/vivae/network.py (`Autoencoder` 36-334):
same constructor arguments, sub-module and parameter names (so `state_dict`s are
interchangeable with the reference), same `forward` contract (a dict of weighted
loss terms).  The MLP layers stay standard PyTorch (north star); only the quartet
term (`MDSLoss`) is a custom CUDA operator.
"""

from __future__ import annotations

from collections import OrderedDict
from collections.abc import Callable
from typing import Optional, Union

import numpy as np
import torch
import torch.nn as nn

from .losses import EncoderGeometricLoss, GeometricLoss, ImitationLoss, KLDivLoss, MDSLoss

__all__ = ["Autoencoder"]


def _weighted(lam: float, term: torch.Tensor) -> torch.Tensor:
    """lam * term (network.py:262-293); a weight of exactly 1 is the identity bit for bit, so the two kernels
    (forward and backward multiply) are skipped -- the training step is bound by its kernel count."""
    return term if lam == 1.0 else lam * term


def _mlp(prefix: str, dims: list[int], activation, final_linear_to: Optional[int]) -> nn.Sequential:
    """Linear(+activation) stack named like the reference: '01<prefix>Potential', '01<prefix>Activation', ..."""
    layers = OrderedDict()
    for i in range(len(dims) - 1):
        layers[f"{i + 1:02d}{prefix}Potential"] = nn.Linear(dims[i], dims[i + 1])
        if activation is not None:
            layers[f"{i + 1:02d}{prefix}Activation"] = activation
    if final_linear_to is not None:
        layers[f"{len(dims):02d}{prefix}Potential"] = nn.Linear(dims[-1], final_linear_to)
    return nn.Sequential(layers)


class Autoencoder(nn.Module):
    """(Variational) autoencoder; reference: network.py:36-111."""

    def __init__(self, input_dim: int, latent_dim: int = 2, hidden_dims: list[int] = [100, 100, 100],
                 variational: bool = False, activation=nn.GELU()):
        super().__init__()
        self.variational = variational
        self.latent_dim = latent_dim
        self.input_dim = input_dim

        enc_dims = [input_dim] + list(hidden_dims)
        dec_dims = [latent_dim] + list(reversed(hidden_dims))
        # construction order (encoder, heads, decoder) matches the reference so that a
        # seeded initialisation consumes the RNG identically (network.py:71-99)
        if variational:
            self.encoder = _mlp("Encoder", enc_dims, activation, None)
            self.mu = nn.Linear(enc_dims[-1], latent_dim)
            self.logvar = nn.Linear(enc_dims[-1], latent_dim)
            self.kldiv_error = KLDivLoss()
        else:
            self.encoder = _mlp("Encoder", enc_dims, activation, latent_dim)
        self.decoder = _mlp("Decoder", dec_dims, activation, input_dim)
        # registration order of the reference: mu, logvar, encoder, decoder
        if variational:
            for name in ("encoder", "decoder"):
                self._modules[name] = self._modules.pop(name)

        self.reconstruction_error = nn.MSELoss()
        self.geometric_error = GeometricLoss()
        self.encoder_geometric_error = EncoderGeometricLoss()
        self.mds_error = MDSLoss()
        self.imit_error = ImitationLoss()

    # ------------------------------------------------------------------ pieces
    def reparameterise(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        """z = mu + eps * exp(logvar / 2), eps ~ N(0, I)  (network.py:113-130)."""
        std = torch.exp(0.5 * logvar)
        return mu + torch.randn_like(std) * std

    def encode(self, x: torch.Tensor, no_grad: bool = False):
        """Encoder pass; a VAE returns (mu, logvar, z) with z the noisy sample (network.py:132-164)."""
        with torch.set_grad_enabled(not no_grad):
            h = self.encoder(x)
            if not self.variational:
                return h
            mu, logvar = self.mu(h), self.logvar(h)
            return mu, logvar, self.reparameterise(mu, logvar)

    def submersion(self, x: torch.Tensor) -> torch.Tensor:
        """Encoder as a function x -> mu (network.py:166-183)."""
        h = self.encoder(x)
        return self.mu(h) if self.variational else h

    def decode(self, z: torch.Tensor, no_grad: bool = False) -> torch.Tensor:
        with torch.set_grad_enabled(not no_grad):
            return self.decoder(z)

    def immersion(self, z: torch.Tensor) -> torch.Tensor:
        return self.decoder(z)

    # ----------------------------------------------------------------- forward
    def forward(self, x: torch.Tensor, lam_recon: float = 1.0, lam_kldiv: float = 1.0, lam_geom: float = 0.0,
                lam_egeom: float = 0.0, lam_mds: float = 0.0, mds_distf_hd: str = "euclidean",
                mds_distf_ld: str = "euclidean", mds_nsamp: int = 1, lam_imit: float = 0.0,
                ref_model: Optional[Callable] = None) -> dict:
        """Weighted loss terms for one batch (network.py:219-296).  Terms with weight 0 are
        skipped and reported as None; the quartet term is fed the *sampled* z (network.py:265,286)."""
        out = {"recon": None, "kldiv": None, "geom": None, "egeom": None, "mds": None, "imit": None}
        if self.variational:
            mu, logvar, z = self.encode(x)
            out["kldiv"] = _weighted(lam_kldiv, self.kldiv_error(mu, logvar))
        else:
            z = self.encode(x)
        if lam_recon > 0.0:
            out["recon"] = _weighted(lam_recon, self.reconstruction_error(x, self.decode(z)))
        if lam_geom > 0.0:
            out["geom"] = _weighted(lam_geom, self.geometric_error(self.immersion, z))
        if lam_egeom > 0.0:
            out["egeom"] = _weighted(lam_egeom, self.encoder_geometric_error(self.submersion, x))
        if lam_mds > 0.0:
            out["mds"] = _weighted(lam_mds, self.mds_error(x, z, distf_hd=mds_distf_hd, distf_ld=mds_distf_ld,
                                                           n_sampling=mds_nsamp))
        if lam_imit > 0.0 and ref_model is not None:
            out["imit"] = _weighted(lam_imit, self.imit_error(x, z, ref_model))
        return out

    # ------------------------------------------------------------------- embed
    def embed(self, x: Union[np.ndarray, torch.Tensor], batch_size: Optional[int] = None) -> np.ndarray:
        """Latent means of `x`, no grad (network.py:298-334).  Chunks are written into one
        preallocated result instead of the reference's repeated `vstack` (quadratic copying)."""
        dev = next(self.parameters()).device
        if isinstance(x, np.ndarray):
            x = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        x = x.to(device=dev, dtype=torch.float32)
        with torch.no_grad():
            if batch_size is None:
                enc = self.submersion(x)
            else:
                enc = torch.empty((x.shape[0], self.latent_dim), dtype=torch.float32, device=dev)
                for s in range(0, x.shape[0], batch_size):
                    enc[s:s + batch_size] = self.submersion(x[s:s + batch_size])
        return enc.detach().cpu().numpy()

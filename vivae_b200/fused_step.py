"""The training step's forward + losses + backward as one CUDA kernel (`csrc/vae_step.cu`).

What it replaces: for one batch, `Autoencoder.forward` (/root/reference/src/vivae/network.py:219-296) and the
autograd backward of `recon + kldiv + mds` (model.py:146-176) -- about 140 kernel launches of a few microseconds
each for a 32k-parameter network.  What it keeps: the network is the same `nn.Module` with the same parameter
tensors (read in place by the kernel), the noise comes from `torch.randn` on the default CUDA generator exactly
where `reparameterise` draws it (network.py:113-130), and `torch.optim.Adam` applies the update -- it finds the
gradients in `.grad` as always; here those are views of one flat buffer the kernel fills.

Used by `ViVAE.train_epoch` when the configuration is the default one: variational, exact-erf GELU between Linear
layers, `lam_recon > 0`, euclidean quartet distances with `mds_nsamp = 1`, no geometric / imitation terms, batch a
multiple of 4.  Anything else runs the PyTorch path.  `VIVAE_FUSED_STEP=0` switches it off.
"""

from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

__all__ = ["FusedStep", "fused_step_supported"]


def _linear_layers(seq: nn.Sequential):
    """The Linear layers of an `_mlp` stack, or None if it holds anything but Linear + exact GELU."""
    out = []
    for m in seq:
        if isinstance(m, nn.Linear):
            out.append(m)
        elif isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none":
            continue
        else:
            return None
    return out


def fused_step_supported(net, cfg: dict, batch: torch.Tensor) -> bool:
    """True when one batch of this configuration can run through `vvb_vae_step_device`."""
    if os.environ.get("VIVAE_FUSED_STEP", "1") == "0" or not batch.is_cuda or batch.dtype != torch.float32:
        return False
    if not getattr(net, "variational", False) or batch.dim() != 2 or batch.shape[0] < 4 or batch.shape[0] % 4:
        return False
    if cfg.get("lam_geom", 0.0) != 0.0 or cfg.get("lam_egeom", 0.0) != 0.0:
        return False
    if cfg.get("lam_imit", 0.0) != 0.0 and cfg.get("ref_model") is not None:
        return False
    if cfg.get("lam_recon", 1.0) <= 0.0 or cfg.get("lam_mds", 0.0) < 0.0 or cfg.get("mds_nsamp", 1) != 1:
        return False
    if cfg.get("mds_distf_hd", "euclidean") != "euclidean" or cfg.get("mds_distf_ld", "euclidean") != "euclidean":
        return False
    enc, dec = _linear_layers(net.encoder), _linear_layers(net.decoder)
    if not enc or not dec or len(enc) + len(dec) + 2 > 16:
        return False
    # GELU after every encoder layer and after all but the last decoder layer (the `_mlp` builder's shape)
    if len(list(net.encoder)) != 2 * len(enc) or len(list(net.decoder)) != 2 * len(dec) - 1:
        return False
    return all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in net.parameters())


class FusedStep:
    """Buffers and pointer tables of the fused step for one (network, batch size)."""

    def __init__(self, net, batch_size: int):
        lib = _lib.load()
        self.layers = _linear_layers(net.encoder) + [net.mu, net.logvar] + _linear_layers(net.decoder)
        self.n_enc = len(_linear_layers(net.encoder))
        self.n_dec = len(_linear_layers(net.decoder))
        n = len(self.layers)
        self.B, self.D, self.L = int(batch_size), int(net.input_dim), int(net.latent_dim)
        self._w = (ctypes.c_void_p * n)(*[m.weight.data_ptr() for m in self.layers])
        self._b = (ctypes.c_void_p * n)(*[m.bias.data_ptr() for m in self.layers])
        self._in = (ctypes.c_int * n)(*[m.in_features for m in self.layers])
        self._out = (ctypes.c_int * n)(*[m.out_features for m in self.layers])
        smem, n_params, n_cta = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(lib.vvb_vae_step_device(self._w, self._b, self._in, self._out, self.n_enc, self.n_dec, None, None,
                                           self.B, self.D, self.L, 1.0, 1.0, 1.0, None, None, None, None, None,
                                           ctypes.byref(smem), ctypes.byref(n_params), ctypes.byref(n_cta), None))
        dev = self.layers[0].weight.device
        self.n_params, self.n_cta = n_params.value, n_cta.value
        self.grad_flat = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grad_part = _lib.raw_empty(self.n_cta * self.n_params, torch.float32, dev)
        self.loss_part = _lib.raw_empty(self.n_cta * 4, torch.float32, dev)
        self.loss_out = torch.zeros(3, dtype=torch.float32, device=dev)
        self._eps_like = torch.zeros((self.B, self.L), dtype=torch.float32, device=dev)  # shape/dtype of `std`
        # every parameter's gradient is a view of the flat buffer, in the kernel's order: weight, then bias, per layer
        self.views = []
        off = 0
        for m in self.layers:
            for prm in (m.weight, m.bias):
                self.views.append((prm, self.grad_flat[off:off + prm.numel()].view_as(prm)))
                off += prm.numel()
        assert off == self.n_params

    def pointers_current(self) -> bool:
        """The tables hold raw pointers: a parameter tensor that was re-allocated (not updated in place) voids them."""
        return all(m.weight.data_ptr() == self._w[i] and m.bias.data_ptr() == self._b[i]
                   for i, m in enumerate(self.layers))

    def step(self, x: torch.Tensor, sums: Optional[torch.Tensor], lam_recon: float, lam_kldiv: float,
             lam_mds: float) -> torch.Tensor:
        """Gradients of `lam_recon * MSE + lam_kldiv * KL + lam_mds * MDS` for batch `x` into every parameter's
        `.grad`; the three weighted terms are added to `sums` (slots 0, 1, 4) and returned (3-element tensor)."""
        lib = _lib.load()
        x = x.contiguous()
        eps = torch.randn_like(self._eps_like)  # reparameterise's draw (network.py:128), same generator, same count
        with torch.cuda.device(x.device):
            _lib.check(lib.vvb_vae_step_device(
                self._w, self._b, self._in, self._out, self.n_enc, self.n_dec, x.data_ptr(), eps.data_ptr(), self.B,
                self.D, self.L, float(lam_recon), float(lam_kldiv), float(lam_mds), self.grad_part.data_ptr(),
                self.loss_part.data_ptr(), self.grad_flat.data_ptr(), None if sums is None else sums.data_ptr(),
                self.loss_out.data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream))
        for prm, view in self.views:
            prm.grad = view
        return self.loss_out

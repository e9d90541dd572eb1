"""Loss terms of the ViVAE objective.

`MDSLoss` (the stochastic quartet loss, reference: /root/reference/src/vivae/losses.py:134-349)
is the hot operator: forward and backward run as fused sm_100a kernels
(`csrc/quartet.cu`) behind a `torch.autograd.Function`.  The remaining terms are
small and stay ordinary PyTorch, as the reference has them (losses.py:58-131); the two geometric
terms (losses.py:352-415) use closed-form Jacobians of the MLPs (`jacobians.py`) instead of `vmap(jacfwd)`.
"""

from __future__ import annotations

import ctypes
from collections.abc import Callable

import torch

from . import _lib
from .jacobians import analytic_jacobian, logdet_gram

__all__ = ["MDSLoss", "KLDivLoss", "ImitationLoss", "FirstNDimsExtractor", "GeometricLoss", "EncoderGeometricLoss",
           "quartet_loss"]


def _check_cuda_f32(name: str, t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"vivae_b200 quartet loss: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"vivae_b200 quartet loss: `{name}` must be float32, got {t.dtype}")
    return t.contiguous()


class _QuartetLossFn(torch.autograd.Function):
    """loss = MDSLoss(x, z); gradient flows to z only (x is data)."""

    @staticmethod
    def forward(ctx, x, z, metric_hd: int, metric_ld: int, perms):
        lib = _lib.load()
        x = _check_cuda_f32("x", x.detach())
        zc = _check_cuda_f32("z", z.detach())
        if x.dim() != 2 or zc.dim() != 2 or x.shape[0] != zc.shape[0]:
            raise ValueError("quartet loss: x and z must be 2-D with the same number of rows")
        B, D = x.shape
        L = zc.shape[1]
        need_grad = ctx.needs_input_grad[1]
        loss = _lib.raw_empty((), torch.float32, x.device)
        grad_unit = _lib.raw_empty(zc.shape, zc.dtype, zc.device) if need_grad else None
        ws = _lib.raw_empty(int(lib.vvb_quartet_workspace_bytes(B)), torch.uint8, x.device)
        n_sampling = len(perms)
        keep = [None if p is None else p.to(device=x.device, dtype=torch.int64).contiguous() for p in perms]
        arr = (ctypes.c_void_p * n_sampling)(*[None if p is None else p.data_ptr() for p in keep])
        with torch.cuda.device(x.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.vvb_quartet_loss_fwd_device(
                x.data_ptr(), zc.data_ptr(), B, D, L, ctypes.cast(arr, ctypes.c_void_p), n_sampling, metric_hd,
                metric_ld, loss.data_ptr(), None if grad_unit is None else grad_unit.data_ptr(), ws.data_ptr(),
                stream))
        if need_grad:
            ctx.save_for_backward(grad_unit)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad_unit,) = ctx.saved_tensors
        lib = _lib.load()
        go = grad_out.detach().to(dtype=torch.float32).contiguous()
        gz = _lib.raw_empty(grad_unit.shape, grad_unit.dtype, grad_unit.device)
        with torch.cuda.device(grad_unit.device):
            stream = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.vvb_quartet_loss_bwd_device(grad_unit.data_ptr(), go.data_ptr(), grad_unit.numel(),
                                                       gz.data_ptr(), stream))
        return None, gz, None, None, None


def quartet_loss(x, z, distf_hd: str = "euclidean", distf_ld: str = "euclidean", perms=None) -> torch.Tensor:
    """Functional form with explicit sampling permutations (`perms[i] is None` = identity)."""
    if distf_hd not in _lib.DIST_FUNCS or distf_ld not in _lib.DIST_FUNCS:
        raise ValueError("Invalid distance function specification for MDS loss")
    perms = [None] if perms is None else list(perms)
    return _QuartetLossFn.apply(x, z, _lib.DIST_FUNCS[distf_hd], _lib.DIST_FUNCS[distf_ld], perms)


class MDSLoss:
    """Stochastic-MDS (quartet) loss; same call signature as the reference's `MDSLoss`."""

    def __call__(self, x: torch.Tensor, z: torch.Tensor, distf_hd: str = "euclidean", distf_ld: str = "euclidean",
                 n_sampling: int = 1) -> torch.Tensor:
        """Quartet q of a pass is rows {q, nq+q, 2nq+q, 3nq+q} of the (re-)permuted batch,
        nq = B // 4 (losses.py:303-304,326-336).  Pass 0 uses the batch order, later passes a
        fresh `torch.randperm(B)` drawn from the default-device generator (losses.py:328-331),
        so seeded runs consume the RNG exactly as the reference does."""
        if distf_hd not in _lib.DIST_FUNCS or distf_ld not in _lib.DIST_FUNCS:
            raise ValueError("Invalid distance function specification for MDS loss")
        n = x.shape[0]
        perms = [None] + [torch.randperm(n, device=x.device) for _ in range(1, n_sampling)]
        return _QuartetLossFn.apply(x, z, _lib.DIST_FUNCS[distf_hd], _lib.DIST_FUNCS[distf_ld], perms)


class KLDivLoss:
    """KL divergence from the latent prior N(0, eps_std^2 I), eps_std = 1e-2 (losses.py:105-131)."""

    def __init__(self):
        self.eps_sq = 1e-2 ** 2
        self._cache = {}  # (device, dtype) -> (eps^2, log eps^2): no host-to-device copy per call

    def __call__(self, mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        key = (mu.device, mu.dtype)
        if key not in self._cache:
            e = torch.tensor([self.eps_sq], dtype=mu.dtype, device=mu.device)
            self._cache[key] = (e, torch.log(e))  # the logarithm once, not once per batch
        eps_sq, log_eps_sq = self._cache[key]
        inner = logvar + log_eps_sq - mu.square() - eps_sq * logvar.exp()
        return -0.5 * inner.mean() / mu.shape[0]


class ImitationLoss:
    """Mean L2 distance between the encoding and a reference embedding (losses.py:58-88)."""

    def __call__(self, x: torch.Tensor, z: torch.Tensor, ref_model: Callable) -> torch.Tensor:
        diff_sq = (z - ref_model(x)).square().sum(dim=1)
        return diff_sq.clamp_min(1e-9).sqrt().mean()


class FirstNDimsExtractor:
    """Reference model that returns the first `latent_dim` input features (losses.py:91-102)."""

    def __init__(self, latent_dim: int):
        self.latent_dim = latent_dim

    def __call__(self, x):
        return x[:, : self.latent_dim]


def _batched_jacobian(f: Callable, x: torch.Tensor) -> torch.Tensor:
    """Per-sample Jacobians of f (geometry.py:43-58)."""
    try:
        return torch.func.vmap(torch.func.jacfwd(f), in_dims=(0,))(x)
    except NotImplementedError:
        return torch.func.vmap(torch.func.jacrev(f), in_dims=(0,))(x)


def _logdet_variance(jac: torch.Tensor, encoder_side: bool) -> torch.Tensor:
    jac = torch.squeeze(jac)
    jt = jac.transpose(1, 2)
    metric = jac @ jt if encoder_side else jt @ jac  # geometry.py:61-76
    return torch.var(torch.logdet(metric))


class GeometricLoss:
    """Variance of log-det of the decoder pull-back metric (losses.py:352-382).  For the network's own
    decoder the Jacobian is propagated in closed form (`jacobians.py`); any other callable goes through
    autodiff like the reference."""

    def __call__(self, immersion: Callable, z: torch.Tensor) -> torch.Tensor:
        jac = analytic_jacobian(immersion, z) if z.dim() == 2 and z.shape[0] > 1 else None
        if jac is not None:
            return torch.var(logdet_gram(jac.transpose(1, 2)))  # J^T J = Gram matrix of the rows of J^T
        return _logdet_variance(_batched_jacobian(immersion, z), encoder_side=False)


class EncoderGeometricLoss:
    """Variance of log-det of J J^T of the encoder (losses.py:385-415); closed-form Jacobian for the
    network's own encoder."""

    def __call__(self, submersion: Callable, x: torch.Tensor) -> torch.Tensor:
        jac = analytic_jacobian(submersion, x) if x.dim() == 2 and x.shape[0] > 1 else None
        if jac is not None:
            return torch.var(logdet_gram(jac))
        return _logdet_variance(_batched_jacobian(submersion, x), encoder_side=True)

"""Closed-form per-row Jacobians of the encoder / decoder MLPs.

The reference obtains them with `vmap(jacfwd(f))` (/root/reference/src/vivae/geometry.py:43-58) --
one traced forward-mode pass per input coordinate -- inside both geometric losses
(losses.py:352-415), which makes an epoch 6.7x slower [SURVEY].  For a stack of `Linear` and
element-wise activations the Jacobian is a product of small matrices, so it is propagated here as a
handful of batched GEMMs next to the ordinary forward pass:

  decoder  z (B, L) -> D   forward mode: T_0 = I_L,  T_l = (T_{l-1} W_l^T) * act'(pre_l)     -> J^T (B, L, D)
  encoder  x (B, D) -> L   reverse mode: R_last = rows of the head, R_{l-1} = (R_l * act'(pre_l)) W_l -> J (B, L, D)

Either way the batched operand has only L (= 2) rows, i.e. the cost is that of two extra forward
passes.  Everything is written in differentiable torch ops, so autograd supplies the second-order
terms the losses need, and -- unlike `torch.func` transforms and the LU-based `torch.logdet` --
the whole thing can be captured in the CUDA graph of the training step.

Stacks this module does not recognise (a user-supplied activation without a known derivative)
return None and the caller keeps the autodiff route.
"""

from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

__all__ = ["sequential_jacobian_fwd", "sequential_jacobian_rev", "analytic_jacobian", "supports_closed_form", "logdet_gram"]

_INV_SQRT2 = 1.0 / math.sqrt(2.0)
_INV_SQRT2PI = 1.0 / math.sqrt(2.0 * math.pi)


def _act_and_derivative(act: nn.Module, pre: torch.Tensor):
    """(act(pre), act'(pre)) for the activations whose derivative is known in closed form, else None."""
    if isinstance(act, nn.GELU):
        if act.approximate != "none":
            return None
        cdf = 0.5 * (1.0 + torch.erf(pre * _INV_SQRT2))
        pdf = torch.exp(-0.5 * pre * pre) * _INV_SQRT2PI
        return pre * cdf, cdf + pre * pdf
    if isinstance(act, nn.ReLU):
        pos = (pre > 0).to(pre.dtype)
        return pre * pos, pos
    if isinstance(act, nn.Tanh):
        t = torch.tanh(pre)
        return t, 1.0 - t * t
    if isinstance(act, nn.Sigmoid):
        s = torch.sigmoid(pre)
        return s, s * (1.0 - s)
    if isinstance(act, (nn.SiLU,)):
        s = torch.sigmoid(pre)
        return pre * s, s * (1.0 + pre * (1.0 - s))
    if isinstance(act, nn.Softplus) and act.beta == 1.0 and act.threshold >= 20:
        return nn.functional.softplus(pre), torch.sigmoid(pre)
    if isinstance(act, nn.ELU):
        neg = act.alpha * torch.exp(torch.clamp(pre, max=0.0))
        return torch.where(pre > 0, pre, neg - act.alpha), torch.where(pre > 0, torch.ones_like(pre), neg)
    if isinstance(act, nn.LeakyReLU):
        slope = torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, act.negative_slope))
        return pre * slope, slope
    if isinstance(act, nn.Identity):
        return pre, torch.ones_like(pre)
    return None


def _layers(seq: nn.Sequential) -> Optional[list]:
    """[(Linear, activation or None), ...] if `seq` alternates Linear / element-wise activation."""
    out, mods, i = [], list(seq), 0  # not .children(): that drops repeats of a shared activation module
    while i < len(mods):
        lin = mods[i]
        if not isinstance(lin, nn.Linear):
            return None
        act = None
        if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear):
            act = mods[i + 1]
            i += 1
        out.append((lin, act))
        i += 1
    return out


def sequential_jacobian_fwd(seq: nn.Sequential, z: torch.Tensor) -> Optional[torch.Tensor]:
    """Jacobians (B, out, in) of `seq` at the rows of `z`, forward mode (cheap when `in` is small)."""
    layers = _layers(seq)
    if layers is None or z.dim() != 2:
        return None
    h = z
    tang = None  # (B, in, width)
    for lin, act in layers:
        pre = nn.functional.linear(h, lin.weight, lin.bias)
        # the first tangent is the identity: I W^T = W^T for every row
        tang = lin.weight.t().unsqueeze(0).expand(z.shape[0], -1, -1) if tang is None else tang @ lin.weight.t()
        if act is None:
            h = pre
        else:
            ad = _act_and_derivative(act, pre)
            if ad is None:
                return None
            h, der = ad
            tang = tang * der.unsqueeze(1)
    return tang.transpose(1, 2)


def sequential_jacobian_rev(seq: nn.Sequential, head: Optional[nn.Linear], x: torch.Tensor) -> Optional[torch.Tensor]:
    """Jacobians (B, out, in) of `head(seq(x))` (or `seq(x)`), reverse mode (cheap when `out` is small)."""
    layers = _layers(seq)
    if layers is None or x.dim() != 2:
        return None
    h, ders = x, []
    for lin, act in layers:
        pre = nn.functional.linear(h, lin.weight, lin.bias)
        if act is None:
            h = pre
            ders.append(None)
        else:
            ad = _act_and_derivative(act, pre)
            if ad is None:
                return None
            h, der = ad
            ders.append(der)
    cot = None  # (B, out, width)
    if head is not None:
        cot = head.weight.unsqueeze(0).expand(x.shape[0], -1, -1)
    for (lin, _), der in zip(reversed(layers), reversed(ders)):
        if cot is None:  # no head: the last layer's output is the function value
            cot = lin.weight.unsqueeze(0).expand(x.shape[0], -1, -1) if der is None else der.unsqueeze(2) * lin.weight
            continue
        if der is not None:
            cot = cot * der.unsqueeze(1)
        cot = cot @ lin.weight
    return cot


def analytic_jacobian(f, x: torch.Tensor) -> Optional[torch.Tensor]:
    """Closed-form Jacobians (B, out, in) if `f` is a network's `submersion` / `immersion` method or a plain
    `nn.Sequential` of Linear + known activations; None otherwise (the caller then differentiates `f`)."""
    if x.dim() != 2 or x.shape[0] == 0:
        return None
    owner, name = getattr(f, "__self__", None), getattr(f, "__name__", "")
    if isinstance(owner, nn.Module) and name == "submersion" and isinstance(getattr(owner, "encoder", None), nn.Sequential):
        return sequential_jacobian_rev(owner.encoder, owner.mu if getattr(owner, "variational", False) else None, x)
    if isinstance(owner, nn.Module) and name == "immersion" and isinstance(getattr(owner, "decoder", None), nn.Sequential):
        return sequential_jacobian_fwd(owner.decoder, x)
    if isinstance(f, nn.Sequential):
        layers = _layers(f)
        if not layers:
            return None
        if layers[0][0].in_features <= layers[-1][0].out_features:
            return sequential_jacobian_fwd(f, x)
        return sequential_jacobian_rev(f, None, x)
    return None


def supports_closed_form(net: nn.Module) -> bool:
    """True if both MLPs of `net` are stacks this module can differentiate in closed form."""
    probe = torch.zeros(1)
    for name in ("encoder", "decoder"):
        seq = getattr(net, name, None)
        layers = _layers(seq) if isinstance(seq, nn.Sequential) else None
        if not layers or any(a is not None and _act_and_derivative(a, probe) is None for _, a in layers):
            return False
    return True


def logdet_gram(mat: torch.Tensor) -> torch.Tensor:
    """log det of the rows-Gram matrices `mat mat^T` of a (B, r, c) batch; closed form for r <= 2
    (no LU factorisation: graph-capturable, and the same NaN / -inf conventions as torch.logdet)."""
    gram = mat @ mat.transpose(1, 2)
    r = gram.shape[1]
    if r == 1:
        return torch.log(gram[:, 0, 0])
    if r == 2:
        return torch.log(gram[:, 0, 0] * gram[:, 1, 1] - gram[:, 0, 1] * gram[:, 1, 0])
    return torch.logdet(gram)

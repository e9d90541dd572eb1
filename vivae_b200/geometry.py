"""Local-distortion diagnostics of a trained model: encoder / decoder indicatrices and decoder
Jacobian determinants.

Mirror of the reference's /root/reference/src/vivae/geometry.py (`jacobian` 43-58,
`metric_tensor` 61-76, `EncoderIndicatome` 79-314, `DecoderIndicatome` 317-447) and
diagnostics.py (`decoder_jacobian_determinants` 41-112, `encoder_indicatrices` 115-188,
`decoder_indicatrices` 191-213): same class and method names, arguments, defaults and
attributes (`act`, `zr`, `zp`, `r`, `stepsize`, `coords`, `patches`).

Re-plumbed for the GPU: the embedding is produced by one chunked
no-grad pass instead of a quadratic `vstack` loop, the grid pick (geometry.py:99-151: two nested
Python loops over grid tiles, each scanning all points) is ONE bucketed arg-min over all points, and
the circles of all selected points are built and submerged as one batch.  Everything stays on the
model's device; the arithmetic (autodiff Jacobians, SVD, metric tensors) is the reference's.
"""

from __future__ import annotations

import math
from collections.abc import Callable
from typing import Optional, Union

import numpy as np
import torch
import torch.func

from .jacobians import analytic_jacobian


def jacobian(f: Callable, x: torch.Tensor) -> torch.Tensor:
    """Per-row Jacobian matrices of `f` at the rows of `x` (geometry.py:43-58): closed form for the
    network's own MLPs (`jacobians.py`), autodiff for anything else."""
    jac = analytic_jacobian(f, x)
    if jac is not None:
        return jac
    try:
        return torch.func.vmap(torch.func.jacfwd(f), in_dims=(0,))(x)
    except NotImplementedError:
        return torch.func.vmap(torch.func.jacrev(f), in_dims=(0,))(x)


def metric_tensor(jac: torch.Tensor, rev: bool = False) -> torch.Tensor:
    """J^T J (immersion / decoder) or J J^T with `rev` (submersion / encoder), geometry.py:61-76."""
    jt = torch.transpose(jac, 1, 2)
    return torch.matmul(jac, jt) if rev else torch.matmul(jt, jac)


def _model_device(model) -> torch.device:
    return next(model.parameters()).device


def _latent_means(model, x: torch.Tensor, batch_size: int) -> torch.Tensor:
    """Encoder means of all rows: chunked, no autograd graph, one concatenation."""
    out = []
    with torch.no_grad():
        for s in range(0, x.shape[0], max(1, int(batch_size))):
            enc = model.encode(x[s:s + batch_size])
            out.append(enc[0] if model.variational else enc)
    return torch.cat(out, dim=0) if out else x.new_zeros((0, model.latent_dim))


def _grid_steps(act: torch.Tensor, n_steps: int):
    """Grid over the embedding's bounding box (geometry.py:258-268 and 368-379)."""
    xmin, xmax = torch.min(act[:, 0]).item(), torch.max(act[:, 0]).item()
    ymin, ymax = torch.min(act[:, 1]).item(), torch.max(act[:, 1]).item()
    nsteps_x = n_steps
    nsteps_y = math.ceil((ymax - ymin) / (xmax - xmin) * nsteps_x)
    stepsize = min((xmax - xmin) / nsteps_x, (ymax - ymin) / nsteps_y)
    xs = torch.linspace(xmin, xmax, steps=nsteps_x, device=act.device)
    ys = torch.linspace(ymin, ymax, steps=nsteps_y, device=act.device)
    return xs, ys, stepsize


class EncoderIndicatome:
    """Set of encoder indicatrices (geometry.py:79-314)."""

    def __init__(self, model, x: Union[np.ndarray, torch.Tensor]):
        self.model = model
        if isinstance(x, np.ndarray):
            x = torch.tensor(x)
        self.x = x.to(_model_device(model))
        self.act = None
        self.stepsize = None
        self.zr = None
        self.zp = None
        self.r = None

    @staticmethod
    def gridpoint_idcs(ref: torch.Tensor, xgrid: torch.Tensor, ygrid: torch.Tensor, padding: float = 0.1,
                       minpts: Optional[int] = None) -> torch.Tensor:
        """For every grid tile, the reference point nearest to the tile centre among the points inside the
        tile's inner window; tiles without points (or with fewer than `minpts`) are skipped.  Result in the
        reference's order: x index outer, y index inner (geometry.py:99-151)."""
        nx, ny = int(xgrid.shape[0]), int(ygrid.shape[0])
        xstep = xgrid[1] - xgrid[0]
        ystep = ygrid[1] - ygrid[0]
        hx = xstep / 2 * (1 - padding)
        hy = ystep / 2 * (1 - padding)
        px, py = ref[:, 0], ref[:, 1]
        # windows of neighbouring tiles do not overlap, so a point can only belong to its nearest tile
        ti = torch.clamp(torch.round((px - xgrid[0]) / xstep).to(torch.int64), 0, nx - 1)
        tj = torch.clamp(torch.round((py - ygrid[0]) / ystep).to(torch.int64), 0, ny - 1)
        x0, x1 = xgrid[ti] - hx, xgrid[ti] + hx
        y0, y1 = ygrid[tj] - hy, ygrid[tj] + hy
        inside = (px > x0) & (px <= x1) & (py > y0) & (py <= y1)
        cx, cy = (x0 + x1) / 2, (y0 + y1) / 2
        d2 = (cx - px) ** 2 + (cy - py) ** 2
        tile = ti * ny + tj
        n_tiles = nx * ny
        big = torch.full((n_tiles,), float("inf"), dtype=d2.dtype, device=ref.device)
        pts = torch.nonzero(inside).squeeze(1)
        best = big.scatter_reduce(0, tile[pts], d2[pts], reduce="amin", include_self=True)
        # first point (lowest index) that attains the tile minimum, like torch.argmin over the pool
        is_best = d2[pts] == best[tile[pts]]
        cand = pts[is_best]
        first = torch.full((n_tiles,), ref.shape[0], dtype=torch.int64, device=ref.device)
        first = first.scatter_reduce(0, tile[cand], cand, reduce="amin", include_self=True)
        ok = first < ref.shape[0]
        if minpts is not None:
            counts = torch.zeros(n_tiles, dtype=torch.int64, device=ref.device).scatter_add_(
                0, tile[pts], torch.ones_like(pts))
            ok &= counts >= minpts
        return first[ok]  # tiles are numbered x-major, so this is the reference's append order

    @staticmethod
    def circle_on_2manifold(origin: torch.Tensor, t1: torch.Tensor = torch.tensor([1.0, 0.0]),
                            t2: torch.Tensor = torch.tensor([0.0, 1.0]), r: float = 0.1, n: int = 50) -> torch.Tensor:
        """`n` points of the circle of radius `r` around `origin` in the plane spanned by the orthogonal
        tangent vectors `t1`, `t2` (geometry.py:153-185)."""
        t1n = t1 / torch.linalg.norm(t1)
        t2n = t2 / torch.linalg.norm(t2)
        if torch.dot(t1n, t2n) > 1e-6:
            raise ValueError("Vectors `t1` and `t2` must be orthogonal.")
        ang = torch.arange(0, 1, step=1 / n, device=origin.device) * 2 * torch.pi
        return origin + r * torch.cos(ang)[:, None] * t1n[None, :] + r * torch.sin(ang)[:, None] * t2n[None, :]

    @staticmethod
    def scale_ellipse(circum: torch.Tensor, origin: torch.Tensor, factor: float = 10.0, log: bool = False) -> torch.Tensor:
        """Scale an ellipse given by circumference points around `origin` (geometry.py:187-211)."""
        vecs = circum - origin
        if log:
            w = torch.log10(torch.norm(vecs, p=2, dim=0))
            vecs = vecs * (w * factor)
        else:
            vecs = vecs * factor
        return origin + vecs

    def compute_indicatrices(self, r: float = 1e-3, batch_size: int = 256, n_steps: int = 20, all_points: bool = False,
                             n_polygon: int = 100):
        """geometry.py:213-286."""
        self.r = r
        self.act = _latent_means(self.model, self.x, batch_size)
        if not all_points:
            xs, ys, self.stepsize = _grid_steps(self.act, n_steps)
            idcs = EncoderIndicatome.gridpoint_idcs(ref=self.act, xgrid=xs, ygrid=ys)
        else:
            idcs = torch.arange(self.act.shape[0], device=self.act.device)
        xr = self.x[idcs]
        self.zr = self.model.submersion(xr)
        jac = jacobian(self.model.submersion, xr)
        _, _, v = torch.svd(jac, some=False)
        # The reference takes `v[:, :2]` (geometry.py:274-275), i.e. the first two ROWS of V, as the tangent
        # pair of every point; kept as is (the rows of an orthogonal matrix are orthonormal too, so its own
        # orthogonality check passes).  They depend on the arbitrary basis the SVD picks for the null space of
        # the 2 x D Jacobian, so `zp` is only reproducible up to that choice -- on any backend.
        t1 = v[:, 0, :] / torch.linalg.norm(v[:, 0, :], dim=1, keepdim=True)
        t2 = v[:, 1, :] / torch.linalg.norm(v[:, 1, :], dim=1, keepdim=True)
        ang = torch.arange(0, 1, step=1 / n_polygon, device=xr.device) * 2 * torch.pi
        circles = (xr[:, None, :] + r * torch.cos(ang)[None, :, None] * t1[:, None, :]
                   + r * torch.sin(ang)[None, :, None] * t2[:, None, :])          # (m, n_polygon, D)
        m, npts, dim = circles.shape
        zp = self.model.submersion(circles.reshape(m * npts, dim)).reshape(m, npts, -1)
        self.zp = [zp[i] for i in range(m)]

    def get_embedding(self):
        return self.act

    def get_polygons(self, scale_factor: float = 1e-2, log: bool = False):
        """matplotlib PatchCollection of the (re-scaled) indicatrices (geometry.py:297-314)."""
        from matplotlib.collections import PatchCollection
        from matplotlib.patches import Polygon

        polys = []
        for i in range(self.zr.shape[0]):
            res = EncoderIndicatome.scale_ellipse(self.zp[i].detach(), self.zr[i].detach(), scale_factor, log)
            polys.append(Polygon(tuple(res.cpu().tolist()), closed=True))
        return PatchCollection(polys)


class DecoderIndicatome:
    """Set of decoder indicatrices (geometry.py:317-447)."""

    def __init__(self, model, x: Union[np.ndarray, torch.Tensor]):
        self.model = model
        if isinstance(x, np.ndarray):
            x = torch.tensor(x, dtype=torch.float32)
        self.x = x.to(_model_device(model))
        self.act = None
        self.patches = None
        self.coords = None
        self.stepsize = None

    def compute_indicatrices(self, batch_size: int = 256, n_steps: int = 20, n_polygon: int = 100):
        """geometry.py:339-397: unit circles of the decoder's pull-back metric on a grid inside the
        embedding's Delaunay hull."""
        from scipy.spatial import Delaunay

        dev = _model_device(self.model)
        self.act = _latent_means(self.model, self.x, batch_size).detach().cpu()
        xs, ys, self.stepsize = _grid_steps(self.act, n_steps)
        mg = torch.meshgrid([xs, ys], indexing="ij")
        coords = torch.vstack([torch.flatten(mg[0]), torch.flatten(mg[1])]).T
        hull = Delaunay(self.act.numpy())
        coords = coords[torch.as_tensor(hull.find_simplex(coords.numpy()) >= 0)]
        self.coords = coords

        phi = torch.linspace(0.0, 2 * np.pi, n_polygon, device=dev)
        raw = torch.stack((torch.sin(phi), torch.cos(phi)))                      # (2, n_polygon)
        metric = metric_tensor(jacobian(self.model.decoder, coords.to(dev)))     # (g, 2, 2)
        # length of every direction in the pull-back metric: sqrt(v^T G v)
        norm = torch.sqrt(torch.einsum("mn,imn->in", raw, torch.einsum("ijk,kl->ijl", metric, raw)))  # (g, n_polygon)
        vecs = raw.T[None, :, :].expand(coords.shape[0], -1, -1)                 # (g, n_polygon, 2)
        norm3 = norm[:, :, None].expand(-1, -1, 2)
        self.patches = torch.where(norm3 != 0, vecs / norm3, torch.zeros_like(vecs)).detach().cpu()

    def get_embedding(self):
        return self.act

    def get_polygons(self, scale_factor: float = 1e-2):
        """geometry.py:407-434."""
        from matplotlib.collections import PatchCollection
        from matplotlib.patches import Polygon

        if self.patches is None:
            raise ValueError("Compute indicatrices first")
        norms = torch.linalg.norm(self.patches.reshape(-1, 2), dim=1)
        min_norm = torch.min(norms[torch.nonzero(norms)])
        normed = self.patches / min_norm * self.stepsize * scale_factor
        anchored = self.coords.unsqueeze(1).expand(*normed.shape) + normed
        p = PatchCollection([Polygon(tuple(v.tolist()), closed=True) for v in anchored])
        p.set_color([0 / 255, 0 / 255, 0 / 255, 0.3])
        return p


# --------------------------------------------------------------------------- diagnostics.py
def decoder_jacobian_determinants(model, x: Union[np.ndarray, torch.Tensor], batch_size: int = 256) -> np.ndarray:
    """log10 determinant of the decoder's pull-back metric per row, divided by |mean| and shifted by -1
    (diagnostics.py:41-112).  The reference indexes every batch with `[0]` (diagnostics.py:94), which for
    an ndarray / tensor input leaves ONE row and makes `jacobian` raise ("both arguments to matmul need to
    be at least 1D"); the documented per-row result is computed here instead."""
    if isinstance(x, np.ndarray):
        x = torch.tensor(x)
    x = x.to(_model_device(model))
    dets = []
    for s in range(0, x.shape[0], max(1, int(batch_size))):
        with torch.no_grad():
            enc = model.encode(x[s:s + batch_size])
            act = enc[0] if model.variational else enc
        dets.append(torch.linalg.det(metric_tensor(jacobian(model.decoder, act))).detach())
    dets = torch.log10(torch.cat(dets))
    rel = dets / torch.abs(torch.mean(dets)) - 1
    return rel.cpu().numpy()


def encoder_indicatrices(model, x, batch_size: int = 256, radius: float = 1e-3, n_steps: int = 20,
                         all_points: bool = False, n_polygon: int = 100) -> EncoderIndicatome:
    """diagnostics.py:115-188."""
    ei = EncoderIndicatome(model=model, x=x)
    ei.compute_indicatrices(r=radius, batch_size=batch_size, n_steps=n_steps, all_points=all_points,
                            n_polygon=n_polygon)
    return ei


def decoder_indicatrices(model, x, batch_size: int = 256, n_steps: int = 20, n_polygon: int = 50) -> DecoderIndicatome:
    """diagnostics.py:191-213."""
    di = DecoderIndicatome(model=model, x=x)
    di.compute_indicatrices(batch_size=batch_size, n_steps=n_steps, n_polygon=n_polygon)
    return di

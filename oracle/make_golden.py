"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

It imports the reference through oracle/refshim.py (CPU backend) and records
input/output vectors of

  * `vivae.smooth`                       (/root/reference/src/vivae/knn.py:147-185)
  * `vivae.losses.MDSLoss()` + autograd  (/root/reference/src/vivae/losses.py:266-349)
  * `vivae.knn.ensure_valid_knn`         (/root/reference/src/vivae/knn.py:24-63)
  * a short `ViVAE.fit` run              (/root/reference/src/vivae/model.py:188-279)
  * the diagnostics of a trained model   (/root/reference/src/vivae/geometry.py, diagnostics.py)

The fixtures are committed; the tests (CPU and GPU) only read the .npz files,
so nothing at test time needs /root/reference.
"""

from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
GOLD = os.path.join(ROOT, "tests", "golden")

import oracle as orc  # noqa: E402
import refshim  # noqa: E402


def mixture(n, d, n_clusters, seed, dtype=np.float32):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.0, 6.0, size=(n_clusters, d))
    lab = rng.integers(0, n_clusters, size=n)
    x = centres[lab] + 0.5 * rng.standard_normal((n, d))
    return np.maximum(x, 0.0).astype(dtype)


def gen_smooth(vv):
    cases = [
        # name, n, d, kgraph, k, coef, n_iter, dtype
        ("a", 400, 7, 20, 10, 1.0, 1, np.float32),
        ("b", 400, 7, 20, 20, 0.1, 3, np.float32),
        ("c", 300, 5, 15, 15, 0.35, 2, np.float64),
        ("d", 2500, 39, 50, 50, 1.0, 1, np.float32),
        ("e", 64, 3, 63, 63, 1.0, 1, np.float32),      # k = n-1
        ("f", 500, 50, 12, 1, 0.5, 1, np.float32),      # k = 1 (self only)
    ]
    for name, n, d, kg, k, coef, n_iter, dt in cases:
        x = mixture(n, d, 6, seed=100 + ord(name), dtype=dt)
        if name == "b":  # exact duplicate rows
            x[10] = x[3]
            x[200] = x[3]
        idx, dist = orc.knn(x.astype(np.float32), kg)
        out = vv.smooth(x, [idx, dist], k=k, coef=coef, n_iter=n_iter)
        assert out.dtype == dt
        np.savez_compressed(
            os.path.join(GOLD, f"smooth_{name}.npz"),
            x=x, idx=idx.astype(np.int32), dist=dist, k=k, coef=coef, n_iter=n_iter, out=out)
        print(f"smooth_{name}: n={n} d={d} k={k} coef={coef} n_iter={n_iter} {dt.__name__}")


def gen_quartet(vv):
    import torch
    from vivae.losses import MDSLoss

    cases = [
        # name, B, D, L, hd, ld, n_sampling
        ("a", 256, 50, 2, "euclidean", "euclidean", 1),
        ("b", 258, 50, 2, "euclidean", "euclidean", 3),
        ("c", 64, 5, 3, "cosine", "euclidean", 2),
        ("d", 128, 38, 2, "euclidean", "cosine", 1),
        ("e", 512, 100, 2, "cosine", "cosine", 2),
        ("f", 7, 4, 2, "euclidean", "euclidean", 1),    # a single quartet, 3 unused rows
        ("g", 2048, 50, 2, "euclidean", "euclidean", 1),
    ]
    for name, B, D, L, hd, ld, ns in cases:
        rng = np.random.default_rng(200 + ord(name))
        x = mixture(B, D, 5, seed=300 + ord(name))
        z = rng.standard_normal((B, L)).astype(np.float32) * 2.0
        if name == "a":  # identical latent points inside one quartet -> clamped sqrt branch
            z[64] = z[0]
        perms = []
        real_randperm = torch.randperm

        def rec_randperm(n, *a, **k):
            p = real_randperm(n, *a, **k)
            perms.append(p.detach().cpu().numpy().astype(np.int64))
            return p

        out = {}
        for tag, tdt in (("f32", torch.float32), ("f64", torch.float64)):
            perms.clear()
            torch.manual_seed(1234)
            torch.randperm = rec_randperm
            try:
                xt = torch.tensor(x, dtype=tdt)
                zt = torch.tensor(z, dtype=tdt, requires_grad=True)
                loss = MDSLoss()(xt, zt, distf_hd=hd, distf_ld=ld, n_sampling=ns)
                loss.backward()
            finally:
                torch.randperm = real_randperm
            out[f"loss_{tag}"] = loss.detach().numpy()
            out[f"grad_{tag}"] = zt.grad.detach().numpy()
            out[f"perms_{tag}"] = np.stack(perms) if perms else np.zeros((0, B), dtype=np.int64)
        assert np.array_equal(out["perms_f32"], out["perms_f64"])
        np.savez_compressed(
            os.path.join(GOLD, f"quartet_{name}.npz"),
            x=x, z=z, hd=hd, ld=ld, n_sampling=ns, perms=out["perms_f32"],
            loss_f32=out["loss_f32"], grad_f32=out["grad_f32"],
            loss_f64=out["loss_f64"], grad_f64=out["grad_f64"])
        print(f"quartet_{name}: B={B} D={D} L={L} {hd}/{ld} ns={ns} loss={float(out['loss_f32']):.6f}")


def gen_validators(vv):
    """ensure_valid_knn / correct_knn_for_duplicates behaviour (knn.py:24-84)."""
    from vivae.knn import ensure_valid_knn

    x = mixture(50, 4, 3, seed=7)
    x[5] = x[2]
    idx, dist = orc.knn(x, 6)
    # put a duplicate's index first in row 5 (what an approximate builder may return)
    swapped = idx.copy()
    pos = int(np.where(swapped[5] == 2)[0][0])
    swapped[5, 0], swapped[5, pos] = swapped[5, pos], swapped[5, 0]
    fixed = ensure_valid_knn(50, [swapped.copy(), dist.copy()])
    as_float = ensure_valid_knn(50, [idx.astype(np.float64), dist.astype(np.float64)])
    np.savez_compressed(
        os.path.join(GOLD, "validators.npz"),
        x=x, idx=idx, dist=dist, swapped=swapped, fixed_idx=fixed[0], fixed_dist=fixed[1],
        float_idx=as_float[0], float_idx_dtype=str(as_float[0].dtype))
    print("validators: ok")


def gen_fit(vv):
    """A short seeded `ViVAE.fit` on CPU with every random draw recorded, so the
    product's training loop can be replayed on identical batches / noise."""
    import torch

    n, d = 1030, 12  # 1030 = 4*256 + 6 -> last batch of 6 rows (one quartet + 2 unused rows)
    X = mixture(n, d, 5, seed=11)
    model = vv.ViVAE(input_dim=d, latent_dim=2, random_state=3)
    init = {k: v.detach().cpu().numpy().copy() for k, v in model.net.state_dict().items()}

    perms, eps = [], []
    real_randperm, real_randn_like = torch.randperm, torch.randn_like

    def rec_randperm(n_, *a, **k):
        p = real_randperm(n_, *a, **k)
        perms.append(p.detach().cpu().numpy().astype(np.int64))
        return p

    def rec_randn_like(t, *a, **k):
        e = real_randn_like(t, *a, **k)
        eps.append(e.detach().cpu().numpy().copy())
        return e

    import io
    from contextlib import redirect_stdout

    torch.randperm, torch.randn_like = rec_randperm, rec_randn_like
    buf = io.StringIO()
    try:
        with redirect_stdout(buf):
            model.fit(X, n_epochs=2, batch_size=256, lam_mds=10.0, verbose=True)
    finally:
        torch.randperm, torch.randn_like = real_randperm, real_randn_like
    emb = model.transform(X)
    final = {k: v.detach().cpu().numpy().copy() for k, v in model.net.state_dict().items()}
    save = {"X": X, "emb": emb, "log": buf.getvalue(), "n_epochs": 2, "batch_size": 256, "lam_mds": 10.0}
    for k_, v in init.items():
        save["init__" + k_] = v
    for k_, v in final.items():
        save["final__" + k_] = v
    save["perms"] = np.stack(perms)                       # one loader permutation per epoch
    save["eps"] = np.concatenate([e.reshape(-1, 2) for e in eps])  # concatenated in draw order
    save["eps_rows"] = np.array([e.shape[0] for e in eps], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "fit_a.npz"), **save)
    print("fit_a:", buf.getvalue().strip().splitlines()[-1])


def gen_fit_cpu(vv):
    """Seeded CPU `fit` WITHOUT the quartet term: exercises loader / optimiser / loss
    bookkeeping only, so the product's host-side training loop can be compared bit for
    bit on a CPU-only box (no custom kernel involved)."""
    import io
    from contextlib import redirect_stdout

    X = mixture(300, 9, 4, seed=21)
    model = vv.ViVAE(input_dim=9, latent_dim=2, random_state=5)
    buf = io.StringIO()
    with redirect_stdout(buf):
        model.fit(X, n_epochs=3, batch_size=64, lam_mds=0.0, lam_kldiv=0.5, verbose=True)
    emb = model.transform(X)
    emb_b = model.transform(X, batch_size=100)
    assert np.array_equal(emb, emb_b) or np.allclose(emb, emb_b, atol=1e-6)
    np.savez_compressed(os.path.join(GOLD, "fit_cpu_nomds.npz"), X=X, emb=emb, log=buf.getvalue(),
                        keys=np.array(list(model.net.state_dict().keys())))
    print("fit_cpu_nomds:", buf.getvalue().strip().splitlines()[-1])



def gen_geometry(vv):
    """Diagnostics of a small trained reference model (geometry.py / diagnostics.py): embedding, grid pick,
    indicatrix centres, decoder metric patches.  The encoder polygons themselves (`zp`) depend on the basis the
    SVD picks for a null space (see vivae_b200/geometry.py) and are not recorded."""
    import torch
    from vivae.geometry import EncoderIndicatome, jacobian, metric_tensor

    rng = np.random.default_rng(5)
    centres = rng.uniform(0.0, 6.0, size=(5, 10))
    X = (centres[rng.integers(0, 5, 900)] + 0.4 * rng.standard_normal((900, 10))).astype(np.float32)
    model = vv.ViVAE(input_dim=10, latent_dim=2, random_state=4)
    model.fit(X, n_epochs=3, batch_size=128, lam_mds=10.0, verbose=False)
    sd = {k: v.detach().cpu().numpy() for k, v in model.net.state_dict().items()}
    ei = model.encoder_indicatrices(X, batch_size=200, radius=1e-3, n_steps=7, n_polygon=16)
    act = ei.act.detach()
    xs = torch.linspace(act[:, 0].min().item(), act[:, 0].max().item(), steps=7)
    ny = int(np.ceil((act[:, 1].max().item() - act[:, 1].min().item()) / (act[:, 0].max().item() - act[:, 0].min().item()) * 7))
    ys = torch.linspace(act[:, 1].min().item(), act[:, 1].max().item(), steps=ny)
    idcs = EncoderIndicatome.gridpoint_idcs(ref=act, xgrid=xs, ygrid=ys)
    di = model.decoder_indicatrices(X, batch_size=200, n_steps=7, n_polygon=12)
    z = act[:50]
    dets = torch.linalg.det(metric_tensor(jacobian(model.net.decoder, z))).detach()
    out = {f"sd__{k}": v for k, v in sd.items()}
    out.update(X=X, act=act.numpy(), idcs=idcs.numpy(), zr=ei.zr.detach().numpy(), stepsize=np.float64(ei.stepsize),
               n_zp=np.int64(len(ei.zp)), zp_shape=np.array(ei.zp[0].shape), dcoords=di.coords.numpy(),
               dpatches=di.patches.detach().numpy(), dstepsize=np.float64(di.stepsize), dets50=dets.numpy())
    np.savez_compressed(os.path.join(GOLD, "geometry_a.npz"), **out)
    print("geometry_a:", idcs.shape[0], "grid points,", di.coords.shape[0], "decoder grid points")

def main():
    if len(sys.argv) > 1 and sys.argv[1] == "geometry":  # regenerate one fixture only
        gen_geometry(refshim.load_reference())
        return
    os.makedirs(GOLD, exist_ok=True)
    orc.build()
    vv = refshim.load_reference()
    gen_smooth(vv)
    gen_quartet(vv)
    gen_validators(vv)
    gen_fit(vv)
    gen_fit_cpu(vv)
    gen_geometry(vv)


if __name__ == "__main__":
    main()

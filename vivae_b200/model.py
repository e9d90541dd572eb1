"""ViVAE model driver (sklearn-like `fit` / `transform`).

Mirrors /root/reference/src/vivae/model.py (`ViVAE` 37-428): same constructor and method
signatures, defaults, attributes (`net`, `data_loader`, `optimizer`, `trained`,
`decoder_active`, `device_name`), progress line and `ValueError`s.

What is re-plumbed for the GPU:
  * batches come from ONE on-device `randperm` per epoch and one `index_select` per batch,
    instead of the DataLoader's per-sample indexing + stack (model.py:249-254);
  * per-term loss sums are accumulated on the device and read back once per epoch,
    instead of three `.item()` host syncs per batch (model.py:151,157,169);
  * the quartet term runs as one fused kernel (see `losses.MDSLoss`);
  * in the default configuration the forward pass, the three loss terms and the backward pass of a batch are ONE
    kernel (`fused_step.FusedStep`, csrc/vae_step.cu) instead of ~140; `VIVAE_FUSED_STEP=0` keeps the PyTorch path;
  * full-size batches replay ONE captured CUDA graph of the whole step (forward, backward, Adam,
    loss bookkeeping).  Disable with `VIVAE_CUDA_GRAPH=0` or `model.use_cuda_graph = False`.
RNG call order is the reference's: the loader permutation comes from a dedicated
generator seeded with `random_state` (model.py:246-248), `randn_like` and the quartet
re-sampling permutations from the global generator.
"""

from __future__ import annotations

import copy
import os
import random
from collections.abc import Callable
from typing import Optional

import numpy as np
import torch

from . import _lib
from .jacobians import supports_closed_form
from .network import Autoencoder

__all__ = ["ViVAE", "DeviceBatchLoader"]


def _device() -> torch.device:
    from . import DEVICE

    return DEVICE


class DeviceBatchLoader:
    """Shuffled mini-batches of a device-resident matrix.

    Iterating yields 1-tuples `(batch,)` like `DataLoader(TensorDataset(X), shuffle=True)`
    does (model.py:249-254), so `train_epoch` reads the same.  `drop_last` is False."""

    def __init__(self, data: torch.Tensor, batch_size: int, generator: Optional[torch.Generator] = None,
                 shuffle: bool = True):
        self.data = data
        self.batch_size = int(batch_size)
        self.generator = generator
        self.shuffle = shuffle
        self._perm_override = None  # test hook: replay a recorded permutation

    def __len__(self) -> int:
        return (self.data.shape[0] + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for idx in self.index_batches():
            yield (self.data.index_select(0, idx),)

    def index_batches(self):
        """The row indices of every batch of one epoch (1-D int64 device tensors), drawing from the generator exactly
        what `__iter__` always drew.  `train_epoch` gathers full-size batches straight into the captured step's input
        buffer with these (one kernel instead of index_select + copy)."""
        n = self.data.shape[0]
        dev = self.data.device
        if self._perm_override is not None:
            perm = self._perm_override()
        elif self.shuffle:
            # keep the generator in lockstep with the reference's DataLoader: every epoch it draws
            # a base seed (torch/utils/data/dataloader.py, `_base_seed`), the epoch's permutation,
            # and -- when the sampler is exhausted -- one more, unused, permutation
            # (RandomSampler.__iter__ ends with `randperm(n)[: num_samples % n]`).
            torch.empty((), dtype=torch.int64, device=dev).random_(generator=self.generator)
            perm = torch.randperm(n, generator=self.generator, device=dev)
        else:
            perm = torch.arange(n, device=dev)
        for s in range(0, n, self.batch_size):
            yield perm[s:s + self.batch_size]
        if self._perm_override is None and self.shuffle:
            torch.randperm(n, generator=self.generator, device=dev)


_TERMS = ("recon", "kldiv", "geom", "egeom", "mds", "imit")


class ViVAE:
    """ViVAE dimension-reduction model (reference: model.py:37-78)."""

    def __init__(self, input_dim: int, latent_dim: int = 2, hidden_dims: list[int] = [32, 64, 128, 32],
                 variational: bool = True, activation=torch.nn.GELU(), random_state: Optional[int] = None):
        from . import DEVICE_NAME

        self.random_state = random_state
        if self.random_state is not None:
            torch.cuda.manual_seed(self.random_state)
            torch.manual_seed(self.random_state)
        dev = _device()
        # like the reference (global default device), parameters are initialised on the device RNG
        with torch.device(dev):
            net = Autoencoder(input_dim=input_dim, latent_dim=latent_dim, hidden_dims=hidden_dims,
                              variational=variational, activation=activation)
        self.net = net.to(dev)
        self.device_name = DEVICE_NAME
        self.data_loader = None
        self.optimizer = None
        self.trained = False
        self.decoder_active = False
        self.use_cuda_graph = os.environ.get("VIVAE_CUDA_GRAPH", "1") != "0"
        self._graph = None  # (key, CUDAGraph, static batch, static loss sums)

    def __repr__(self):
        return f"ViVAE(input_dim={self.net.input_dim}, latent_dim={self.net.latent_dim}, device={self.device_name})"

    def __str__(self):
        return f"ViVAE model {self.net.input_dim}->{self.net.latent_dim}, {'trained' if self.trained else 'not trained'}"

    # ------------------------------------------------------------------ training
    def train_epoch(self, lam_recon: float = 1.0, lam_kldiv: float = 1.0, lam_geom: float = 0.0,
                    lam_egeom: float = 0.0, lam_mds: float = 100.0, mds_distf_hd: str = "euclidean",
                    mds_distf_ld: str = "euclidean", mds_nsamp: int = 1, lam_imit: float = 0.0,
                    ref_model: Optional[Callable] = None) -> dict:
        """One pass over the data (reference: model.py:89-186); returns the per-term loss sums."""
        if self.data_loader is None:
            raise ValueError("No data to train on: use the `fit` or `fit_transform` method")
        dev = next(self.net.parameters()).device
        cfg = dict(lam_recon=lam_recon, lam_kldiv=lam_kldiv, lam_geom=lam_geom, lam_egeom=lam_egeom, lam_mds=lam_mds,
                   mds_distf_hd=mds_distf_hd, mds_distf_ld=mds_distf_ld, mds_nsamp=mds_nsamp, lam_imit=lam_imit,
                   ref_model=ref_model)
        sums = torch.zeros(len(_TERMS), dtype=torch.float32, device=dev)
        # the geometric terms are graph-safe only with the closed-form Jacobians (no torch.func, no LU logdet)
        geom_ok = (lam_geom == 0.0 and lam_egeom == 0.0) or (supports_closed_form(self.net) and self.net.latent_dim <= 2)
        graph_ok = (self.use_cuda_graph and dev.type == "cuda" and geom_ok
                    and mds_nsamp == 1 and (lam_imit == 0.0 or ref_model is None)
                    and isinstance(self.data_loader, DeviceBatchLoader))
        full = getattr(self.data_loader, "batch_size", None)
        with _lib.no_uninitialised_fill():  # also in force while the step is captured
            if graph_ok:
                data = self.data_loader.data
                for idx in self.data_loader.index_batches():
                    if idx.shape[0] == full and full >= 4:
                        g = self._graph if self._graph is not None and self._graph[1] is not None else None
                        if g is None or g[2].shape[0] != full:
                            g = self._graphed_step(data.index_select(0, idx), cfg)
                        else:
                            g = self._graphed_step(g[2], cfg)  # same key: returns the captured graph (or re-captures)
                        if g is not None:
                            _, graph, static_x, static_sums, n_kernels = g
                            torch.index_select(data, 0, idx, out=static_x)  # the batch, straight into the graph's input
                            graph.replay()  # adds the batch's terms into static_sums
                            _lib.note_graph_replay(n_kernels)
                            continue
                    self._eager_step(data.index_select(0, idx), cfg, sums)
                if self._graph is not None and self._graph[1] is not None:
                    sums += self._graph[3]  # the captured steps' running sums: read once per epoch
                    self._graph[3].zero_()
            else:
                for batch in self.data_loader:
                    self._eager_step(batch[0], cfg, sums)
        host = sums.tolist()  # the only host sync of the epoch
        return dict(zip(_TERMS, host))

    # one optimisation step on batch x; adds the batch's loss terms into `sums` (device tensor)
    def _eager_step(self, x: torch.Tensor, cfg: dict, sums: torch.Tensor) -> None:
        fused = self._fused_for(x, cfg)
        if fused is not None:
            # forward, the three loss terms and backward as ONE kernel (csrc/vae_step.cu); the gradients land in the
            # parameters' .grad (views of one flat buffer), the weighted terms in `sums`
            fused.step(x, sums, cfg["lam_recon"], cfg["lam_kldiv"], cfg["lam_mds"])
            self.optimizer.step()
            return
        self.optimizer.zero_grad()
        terms = self.net(x, **cfg)
        loss = None
        vals = []
        for name in _TERMS:  # same summation order as model.py:149-174
            t = terms[name]
            if name == "kldiv" and not self.net.variational:
                t = None
            if t is None:
                vals.append(None)
                continue
            vals.append(t.detach().reshape(()))
            loss = t if loss is None else loss + t
        with torch.no_grad():  # graph-capture safe: no host-side index tensors
            zero = getattr(self, "_zero", None)
            if zero is None or zero.device != sums.device:
                zero = self._zero = torch.zeros((), dtype=torch.float32, device=sums.device)  # once, not per step
            sums += torch.stack([zero if v is None else v.to(torch.float32) for v in vals])
        loss.backward()
        self.optimizer.step()

    def _fused_for(self, x: torch.Tensor, cfg: dict):
        """The fused-step object for this batch, or None when the configuration needs the PyTorch path
        (`fused_step.fused_step_supported`); built once per (batch size, optimiser)."""
        from .fused_step import FusedStep, fused_step_supported

        if not fused_step_supported(self.net, cfg, x):
            return None
        key = (x.shape[0], id(self.optimizer))
        cache = getattr(self, "_fused", None)
        if cache is None:
            cache = self._fused = {}
        fs = cache.get(key)
        if fs is None or not fs.pointers_current():
            try:
                fs = FusedStep(self.net, x.shape[0])
            except NotImplementedError:  # shapes outside the kernel's envelope: PyTorch path
                fs = False
            if len(cache) > 8:
                cache.clear()
            cache[key] = fs
        return fs or None

    def _graphed_step(self, x: torch.Tensor, cfg: dict):
        """Capture (once per configuration) the whole step as a CUDA graph.  Warm-up steps run on a
        side stream first (lazy cuBLAS / optimiser state initialisation); parameters, optimiser
        state and the RNG are restored afterwards so training is unaffected by the warm-up."""
        key = (tuple(x.shape), id(self.optimizer), tuple(sorted((k, str(v)) for k, v in cfg.items())))
        if self._graph is not None and self._graph[0] == key:
            return self._graph
        if self._graph is not None and self._graph[0] == ("failed",) + key:
            return None
        dev = x.device
        try:
            static_x = x.clone()
            static_sums = torch.zeros(len(_TERMS), dtype=torch.float32, device=dev)
            params = [p.detach().clone() for p in self.net.parameters()]
            opt_state = copy.deepcopy(self.optimizer.state_dict())
            rng = torch.cuda.get_rng_state(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager_step(static_x, cfg, static_sums)
            torch.cuda.current_stream(dev).wait_stream(side)
            # undo the warm-up: same tensors, original contents
            with torch.no_grad():
                for p, saved in zip(self.net.parameters(), params):
                    p.copy_(saved)
                for st in self.optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            for grp, saved in zip(self.optimizer.param_groups, opt_state["param_groups"]):
                for k_, v_ in saved.items():
                    if k_ != "params":
                        grp[k_] = v_
            if opt_state["state"]:
                self.optimizer.load_state_dict(opt_state)
            torch.cuda.set_rng_state(rng, dev)
            self.optimizer.zero_grad(set_to_none=True)
            static_sums.zero_()
            graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(graph):
                self._eager_step(static_x, cfg, static_sums)
            n_kernels = _lib.launch_count() - before  # this library's kernel nodes in the graph
            _lib.note_graph_replay(-n_kernels)         # capture records, it does not execute
            self._graph = (key, graph, static_x, static_sums, n_kernels)
        except Exception as exc:  # capture unsupported for this configuration: stay eager, loudly
            import warnings

            warnings.warn(f"vivae_b200: CUDA-graph capture of the training step failed ({exc!r}); running eagerly")
            torch.cuda.synchronize(dev)
            self._graph = (("failed",) + key, None, None, None, 0)
            return None
        return self._graph

    def fit(self, X: np.ndarray, n_epochs: int = 50, batch_size: int = 256, learning_rate: float = 1e-3,
            weight_decay: float = 1e-4, lam_recon: float = 1.0, lam_kldiv: float = 1.0, lam_geom: float = 0.0,
            lam_egeom: float = 0.0, lam_mds: float = 100.0, mds_distf_hd: str = "euclidean",
            mds_distf_ld: str = "euclidean", mds_nsamp: int = 1, lam_imit: float = 0.0,
            ref_model: Optional[Callable] = None, verbose: bool = True):
        """Fit the model (reference: model.py:188-279)."""
        if X.shape[1] != self.net.input_dim:
            raise ValueError("Mismatch between data dimensionality and network shape")
        if lam_geom > 0.0 and lam_recon == 0.0:
            raise ValueError("Decoder geometric loss can only be used if reconstruction loss is used")

        if self.random_state is not None:
            np.random.seed(self.random_state)
            torch.manual_seed(self.random_state)
            random.seed(self.random_state)

        dev = next(self.net.parameters()).device
        g = torch.Generator(device=dev)
        if self.random_state is not None:
            g.manual_seed(self.random_state)
        # whole matrix to the device once (model.py:250); float64 input is cast (the reference
        # would raise a dtype mismatch inside nn.Linear)
        data = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).to(dev)
        self.data_loader = DeviceBatchLoader(data, batch_size=batch_size, generator=g, shuffle=True)
        graphs = self.use_cuda_graph and dev.type == "cuda"
        # one fused multi-tensor kernel per step instead of ~15 foreach kernels (same update rule as the
        # reference's torch.optim.Adam, model.py:256; L2 weight decay folded into the gradient)
        fused = graphs and os.environ.get("VVB_ADAM_FUSED", "1") != "0"
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=learning_rate, weight_decay=weight_decay,
                                          capturable=graphs, fused=True if fused else None)
        self._graph = None
        self.net.train()

        for epoch in range(1, n_epochs + 1):
            losses = self.train_epoch(lam_recon=lam_recon, lam_kldiv=lam_kldiv, lam_geom=lam_geom,
                                      lam_egeom=lam_egeom, lam_mds=lam_mds, mds_distf_hd=mds_distf_hd,
                                      mds_distf_ld=mds_distf_ld, mds_nsamp=mds_nsamp, lam_imit=lam_imit,
                                      ref_model=ref_model)
            if verbose:
                # the reference divides the epoch sums by the epoch INDEX (model.py:272-275); kept as is
                print(
                    f"Epoch {epoch}/{n_epochs}\trecon: {losses['recon']/epoch:.4f}\tkldiv: {losses['kldiv']/epoch:.4f}\tgeom: {losses['geom']/epoch:.4f}\tegeom: {losses['egeom']/epoch:.4f}\tmds: {losses['mds']/epoch:.4f}\timit: {losses['imit']/epoch:.4f}"
                )
        self.trained = True
        if lam_recon > 0.0:
            self.decoder_active = True

    def transform(self, X: np.ndarray, batch_size: Optional[int] = None) -> np.ndarray:
        """Latent means of X (reference: model.py:281-292)."""
        return self.net.embed(X, batch_size=batch_size)

    def fit_transform(self, X: np.ndarray, n_epochs: int = 50, batch_size: int = 256, learning_rate: float = 1e-3,
                      weight_decay: float = 1e-4, lam_recon: float = 1.0, lam_kldiv: float = 1.0,
                      lam_geom: float = 0.0, lam_egeom: float = 0.0, lam_mds: float = 100.0,
                      mds_distf_hd: str = "euclidean", mds_distf_ld: str = "euclidean", mds_nsamp: int = 1,
                      lam_imit: float = 0.0, ref_model: Optional[Callable] = None, verbose: bool = True):
        """`fit` then `transform` with the same batch size (reference: model.py:294-351)."""
        self.fit(X, n_epochs=n_epochs, batch_size=batch_size, learning_rate=learning_rate, weight_decay=weight_decay,
                 lam_recon=lam_recon, lam_kldiv=lam_kldiv, lam_geom=lam_geom, lam_egeom=lam_egeom, lam_mds=lam_mds,
                 mds_distf_hd=mds_distf_hd, mds_distf_ld=mds_distf_ld, mds_nsamp=mds_nsamp, lam_imit=lam_imit,
                 ref_model=ref_model, verbose=verbose)
        return self.transform(X, batch_size=batch_size)

    # ---- diagnostics (reference: model.py:352-431; implementation in vivae_b200/geometry.py)
    def decoder_jacobian_determinants(self, X: np.ndarray, batch_size: int = 256) -> np.ndarray:
        """Scaled Jacobian determinants of the decoder per row of `X` (model.py:352-364)."""
        from .geometry import decoder_jacobian_determinants

        return decoder_jacobian_determinants(model=self.net, x=np.ascontiguousarray(X, dtype=np.float32),
                                             batch_size=batch_size)

    def encoder_indicatrices(self, X: np.ndarray, batch_size: int = 256, radius: float = 1e-4, n_steps: int = 50,
                             all_points: bool = False, n_polygon: int = 100):
        """Encoder indicatrices on a grid over the embedding (model.py:366-398)."""
        from .geometry import encoder_indicatrices

        return encoder_indicatrices(model=self.net, x=np.ascontiguousarray(X, dtype=np.float32), batch_size=batch_size,
                                    radius=radius, n_steps=n_steps, all_points=all_points, n_polygon=n_polygon)

    def decoder_indicatrices(self, X: np.ndarray, batch_size: int = 256, n_steps: int = 50, n_polygon: int = 100):
        """Decoder indicatrices on a grid inside the embedding's hull (model.py:400-431)."""
        if not self.decoder_active:
            raise RuntimeError("Cannot compute decoder indicatrices if decoder was not active in training")
        from .geometry import decoder_indicatrices

        return decoder_indicatrices(model=self.net, x=np.ascontiguousarray(X, dtype=np.float32), batch_size=batch_size,
                                    n_steps=n_steps, n_polygon=n_polygon)

"""vipsy's estimator classes (vi.py:521-981) on the fused B200 path.

Same class names, constructor arguments and ``fit(optim, loss, max_iter,
random_instance)`` API as the reference; the per-iteration ELBO + gradient
evaluation -- what ``SVI.step`` (vi.py:503-516) spends its time in -- is one call
into the sm_100a kernels through the C ABI (``vipsy_b200.ops``).  The encoder MLPs
stay in PyTorch and receive the gradient of their outputs from the kernel.

There is no CPU fallback: the classes need a CUDA device and the built library.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .infer import (Adam, PyroLRScheduler, SVI, Trace_ELBO, TraceEnum_ELBO, constraints, get_param_store,
                    LOWER_CHOLESKY, POSITIVE, REAL, UNIT_INTERVAL)

try:                                    # progress bar like the reference; optional
    from tqdm import trange
except Exception:                       # pragma: no cover
    trange = None

Site = Tuple[str, torch.Tensor, Optional[torch.Tensor]]


def _require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("vipsy_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _accum(leaf: torch.Tensor, g: torch.Tensor, idx: Optional[torch.Tensor] = None):
    g = g.reshape((-1,) + tuple(leaf.shape[1:])) if idx is not None else g.reshape(leaf.shape)
    if leaf.grad is None:
        if idx is None:
            leaf.grad = g.clone() if g.data_ptr() == leaf.data_ptr() else g.contiguous()
            return
        leaf.grad = torch.zeros_like(leaf)
    if idx is None:
        leaf.grad.add_(g)
    else:
        leaf.grad.index_add_(0, idx, g)


# =========================================================================== encoders
class NormEncoder(nn.Module):           # vi.py:417-435
    def __init__(self, item_size, x_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(item_size, hidden_dim)
        self.fc21 = nn.Linear(hidden_dim, x_dim)
        self.fc22 = nn.Linear(hidden_dim, x_dim)
        self.softplus = nn.Softplus()

    def forward(self, x):
        hidden = self.softplus(self.fc1(x))
        return self.fc21(hidden), torch.exp(self.fc22(hidden))


class MvnEncoder(nn.Module):            # vi.py:438-455
    """``forward`` returns (loc, packed raw tril) -- the K(K+1)/2 vector fc22 emits, in
    ``torch.tril_indices`` order.  The reference scatters it into a dense K x K and applies
    LowerCholeskyTransform (vi.py:452-454, 693); the kernel consumes the packed vector
    directly and fuses the transform, which is the same function of the same numbers."""

    def __init__(self, item_size, x_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(item_size, hidden_dim)
        self.fc21 = nn.Linear(hidden_dim, x_dim)
        self.fc22 = nn.Linear(hidden_dim, int(x_dim * (x_dim + 1) / 2))
        self.x_dim = x_dim
        self.softplus = nn.Softplus()

    def forward(self, x):
        hidden = self.softplus(self.fc1(x))
        return self.fc21(hidden), self.fc22(hidden)


class BinEncoder(nn.Module):            # vi.py:458-470
    def __init__(self, item_size, x_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(item_size, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, x_dim)
        self.softplus = nn.Softplus()

    def forward(self, x):
        return torch.sigmoid(self.fc2(self.softplus(self.fc1(x))))


class SoftmaxEncoder(nn.Module):        # vi.py:473-485 (softmax over dim 0, sic)
    def __init__(self, item_size, x_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(item_size, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, x_dim)

    def forward(self, x):
        return torch.softmax(self.fc2(torch.relu(self.fc1(x))), dim=0)


# =========================================================================== base
class BasePsy(nn.Module):               # vi.py:521-533
    def __init__(self, data, subsample_size=None, device=None, seed: int = 0, **kwargs):
        super().__init__()
        self.device = _require_cuda(device)
        self.data = data
        self.sample_size = data.size(0)
        self.item_size = data.size(1)
        self.subsample_size = subsample_size if subsample_size is not None else self.sample_size
        self.kwargs = kwargs
        self._data_dev = data.to(self.device, torch.float32)
        self._codes = ops.pack_responses(self._data_dev)        # uint8 0/1/255, stays resident
        # mostly-missing matrices take the compaction path of the K = 1 kernel (same results)
        self._n_missing = int((self._codes == 255).sum().item()) if self._codes.numel() else 0
        self._sparse = self._n_missing > 0.5 * self._codes.numel()
        self._obs_csr = None          # CSR of the observed cells, built on first use (K = 1, mostly-missing data)
        self._enc_x_full = None       # encoder input of the full batch (NaN -> -1), built on first use
        self._x_tiles, self._x_tiles_tried = None, False      # resident operand tiles of the encoder kernels
        self._seed, self._calls = int(seed), 0
        # device-resident RNG state (SVI(cuda_graph=True)): the noise call counter is a device scalar and the
        # minibatch is drawn with graph-safe torch ops, so a captured iteration differs on every replay
        self._device_rng = False
        self._calls_dev = None
        self._plate_calls_dev = None  # minibatch-draw counter (ops.sample_indices)
        # optional person sharding (vipsy_b200.dist.attach): this rank holds rows
        # [person_offset, person_offset + sample_size) of a global_size-person job
        self.global_size = self.sample_size
        self.person_offset = 0
        self.reducer = None

    def _observed_csr(self):
        """(row_ptr, entries, nnz) of the observed cells; the matrix is constant, so this is built once."""
        if self._obs_csr is None and self.item_size <= 32768:
            self._obs_csr = ops.observed_csr(self._codes)
        return self._obs_csr

    # ---- plate('data', N, subsample_size=B): idx + scale -------------
    def _plate(self):
        """idx (None = all rows in order), batch size, N/B of this rank's shard.  With person
        sharding every rank subsamples its own shard proportionally, so N/B is the global ratio."""
        N, B = self.sample_size, min(self.subsample_size, self.sample_size)
        if B >= N:
            return None, N, 1.0
        if B <= ops.SAMPLE_MAX_BATCH:             # one launch, counter in device memory: the same index stream in
            if self._plate_calls_dev is None:     # eager and captured iterations (graph-safe, no host sync)
                self._plate_calls_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
            return ops.sample_indices(N, B, self._seed ^ (self.person_offset * 0x9E3779B1), self._plate_calls_dev), B, N / B
        if self._device_rng:          # uniform random B-subset in random order, no host sync (graph-safe)
            return torch.rand(N, device=self.device).topk(B).indices, B, N / B
        return torch.randperm(N, device=self.device)[:B], B, N / B

    def _noise_args(self, kind=None, S=1, B=0, K=1):
        """Noise arguments of one kernel call.  Default: (seed, offset, person_offset) for the in-kernel Philox
        generator.  With device-resident RNG state the same generator is run by ``ops.fill_*`` with its call
        counter read from device memory and the draws are injected (``eps`` / ``u``): identical streams."""
        self._calls += 1
        if not self._device_rng or kind is None:
            if self._calls_dev is not None:       # both counters advance on every call: one stream across modes
                self._calls_dev.add_(1)
            return dict(seed=self._seed, offset=self._calls, person_offset=self.person_offset)
        if self._calls_dev is None:   # starts at the host count: eager evaluations before a captured fit are not replayed
            self._calls_dev = torch.full((1,), self._calls - 1, dtype=torch.int64, device=self.device)
        self._calls_dev.add_(1)
        fill = ops.fill_normal if kind == "eps" else ops.fill_uniform
        # the host counter and the device counter advance together (SVI adds the replayed calls to the host
        # count), so switching between eager, captured and read-out calls never reuses a (seed, offset) stream
        return {kind: fill(S, B, K, self._seed, self._calls_dev, self.person_offset,
                           device=self.device)}

    def _encoder_input(self, idx):
        if idx is None:                   # full batch: the matrix is constant, substitute once
            if self._enc_x_full is None:
                x = self._data_dev
                self._enc_x_full = torch.where(torch.isnan(x), torch.full_like(x, -1.0), x)
            return self._enc_x_full
        x = self._data_dev[idx]
        return torch.where(torch.isnan(x), torch.full_like(x, -1.0), x)      # vi.py:679-682

    def _norm_encoder_kernels(self, idx):
        """``NormEncoder`` (vi.py:417-435) on this batch through the CUDA kernels (csrc/enc_norm*.cu): no fp32 input
        matrix, no stored hidden activations, deterministic weight gradients.  Returns mu [B, 1], sigma [B, 1] and
        ``backward(d_mu, d_sigma, weight)`` that accumulates the gradients of the six encoder tensors."""
        enc = self.encoder
        if idx is None and self._x_tiles is None and not self._x_tiles_tried:
            self._x_tiles_tried = True                  # full batch: resident operand tiles, built once
            self._x_tiles = ops.norm_encoder_pack_tiles(self._codes)
        xt = self._x_tiles if idx is None else None
        mu, sigma, pre = ops.norm_encoder_forward(self._codes, enc, idx, x_tiles=xt)

        def backward(d_mu, d_sigma, w):
            g = ops.norm_encoder_backward(self._codes, enc, sigma, pre, d_mu, d_sigma, idx, grad_scale=w, x_tiles=xt)
            for n, p in enc.named_parameters():
                _accum(p, g[n])
        return mu.unsqueeze(1), sigma.unsqueeze(1), backward

    def _norm_encoder_fusable(self) -> bool:
        return (getattr(self, "fused_encoder", False) and isinstance(self.encoder, NormEncoder)
                and self.encoder.fc21.out_features == 1
                and self.encoder.fc1.out_features <= ops.NORM_ENCODER_MAX_HIDDEN)

    def _mvn_encoder_fusable(self) -> bool:
        return (getattr(self, "fused_encoder", False) and isinstance(self.encoder, MvnEncoder)
                and self.encoder.fc1.out_features <= ops.MVN_ENCODER_MAX_HIDDEN and 2 <= self.x_feature <= 128)

    def _finish(self, loss: torch.Tensor, sites: List[Site]):
        """Person-sharded jobs: all-reduce the loss and every replicated parameter's gradient."""
        if self.reducer is not None:
            loss = self.reducer(loss, [p for n, p, _ in sites if not self._is_local(n)])
        return loss, sites

    def _is_local(self, name: str) -> bool:
        return name in ("x_local", "x_scale", "attr_p", "theta_local", "theta_scale") and not (
            name == "x_scale" and getattr(self, "share_cov", False) and getattr(self, "x_feature", 1) > 1)

    def _use_graph(self, cuda_graph, optim, S) -> bool:
        """cuda_graph=None: capture the iteration whenever it can be (Adam-family optimiser, single process).  Small
        iterations are launch- and interpreter-bound (c1: 570 -> 29 us); large ones still shed the ~0.5 ms of host
        glue per step.  Pass cuda_graph=False for the eager form (torch.optim objects between launches)."""
        if cuda_graph is None:
            return None if self.reducer is None else False     # None: SVI captures when the optimiser allows it
        return bool(cuda_graph)

    def _fit_loop(self, svi, optim, max_iter, postfix_fn, evaluate=False):
        it = trange(max_iter) if trange is not None else range(max_iter)
        loss = float("nan")
        for i in it:
            show = trange is not None and (i % 50 == 0 or i == max_iter - 1)
            last = i == max_iter - 1
            # the loss is read back (a device sync) only when it is shown or returned
            loss_t = svi.step_async(self.data)
            if evaluate:
                loss = svi.evaluate_loss(self.data)
            elif show or last:
                loss = float(loss_t)
            if isinstance(optim, PyroLRScheduler):
                optim.step()
            if show:
                with torch.no_grad():
                    it.set_postfix(loss="{0:1.2f}".format(loss), **postfix_fn())
        return loss


# =========================================================================== IRT
class BaseIRT(BasePsy):                 # vi.py:536-656
    IRT_FUN = ("irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl")

    def __init__(self, model="irt_2pl", x_feature=1, share_cov=False, D=1, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if model not in self.IRT_FUN:
            raise KeyError(model)
        self._model, self.x_feature, self.share_cov, self.D = model, x_feature, share_cov, D
        self.a_free = kwargs.get("a_free")
        self.a0 = kwargs.get("a0", torch.ones((self.x_feature, self.item_size)))
        if x_feature > 1 and self.a_free is None:           # staircase identification, vi.py:568-572
            self.a_free = torch.ones(x_feature, self.item_size, dtype=torch.bool)
            self.a0 = self.a0.clone()
            for i in range(x_feature):
                self.a_free[i, self.item_size - i:] = False
                self.a0[i, self.item_size - i:] = 0
        self._a_free_dev = None if self.a_free is None else self.a_free.to(self.device)

    # ---- item parameter sites (BaseIRT.model, vi.py:577-587) -----------------------
    def _item_params(self):
        st, J, dev = get_param_store(), self.item_size, self.device
        b_u = st.setdefault("b", self.kwargs.get("b0", torch.zeros((1, J))), REAL, dev)
        out = {"b": (b_u, b_u.detach())}
        if self._model != "irt_1pl":
            a_u = st.setdefault("a", self.a0, REAL, dev)
            out["a"] = (a_u, a_u.detach())
        if self._model in ("irt_3pl", "irt_4pl"):
            c_u = st.setdefault("c", torch.zeros((1, J)) + 0.1, UNIT_INTERVAL, dev)
            out["c"] = (c_u, torch.sigmoid(c_u.detach()))
        if self._model == "irt_4pl":
            d_u = st.setdefault("d", torch.ones((1, J)) - 0.1, UNIT_INTERVAL, dev)
            out["d"] = (d_u, torch.sigmoid(d_u.detach()))
        return out

    def _push_item_grads(self, items, r, w):
        for n in items:
            leaf, val = items[n]
            g = r["d_" + n] * w
            if n in ("c", "d"):
                g = g * val * (1 - val)                       # sigmoid chain rule
            _accum(leaf, g)

    def _item_sites(self, items) -> List[Site]:
        return [(n, items[n][0], self._a_free_dev if n == "a" else None) for n in items]

    def _guide(self, idx, B):
        """Returns (kernel kwargs, backward closure(r, weight), sites)."""
        raise NotImplementedError

    def loss_and_grads(self, cfg: Trace_ELBO, data=None, want_grad=True):
        N, B = self.sample_size, min(self.subsample_size, self.sample_size)
        S = cfg.num_particles
        groups = [(None, S)] if B >= N else [(None, 1)] * S       # fresh minibatch per particle (A.1)
        items = self._item_params()
        total = None
        sites: List[Site] = self._item_sites(items)
        seen = {n for n, _, _ in sites}
        for _, s in groups:
            idx, Bg, scale = self._plate()
            gkw, backward, gsites = self._guide(idx, Bg)
            noise = self._noise_args("eps", s, Bg, self.x_feature)
            r = ops.irt_elbo_grad(self._codes, model=self._model, D=float(self.D), scale=scale, n_particles=s,
                                  a=items["a"][1] if "a" in items else None, b=items["b"][1],
                                  c=items["c"][1] if "c" in items else None, d=items["d"][1] if "d" in items else None,
                                  idx=idx, a_free=self._a_free_dev, estimator=cfg.estimator, baseline=cfg.baseline,
                                  want_grad=want_grad, sparse=self._sparse and self.x_feature == 1,
                                  n_masked=self._n_missing if idx is None else None,
                                  obs=self._observed_csr() if (self._sparse and self.x_feature == 1) else None,
                                  **gkw, **noise)
            w = 1.0 / len(groups)
            total = r["loss"] * w if total is None else total + r["loss"] * w
            if want_grad:
                self._push_item_grads(items, r, w)
                backward(r, w)
            for sname, leaf, free in gsites:
                if sname not in seen:
                    sites.append((sname, leaf, free))
                    seen.add(sname)
        return self._finish(total, sites)

    def loss(self, cfg, data=None):
        return self.loss_and_grads(cfg, data, want_grad=False)[0]

    def model(self, data):       # bound-method handles for SVI(model, guide, ...) like the reference
        raise RuntimeError("the model is evaluated by the fused kernel; call fit() or SVI.step()")

    def guide(self, data):
        raise RuntimeError("the guide is evaluated by the fused kernel; call fit() or SVI.step()")

    def fit(self, optim=None, loss=None, max_iter=5000, random_instance=None, cuda_graph=None):     # vi.py:627-656
        optim = Adam({"lr": 5e-2}) if optim is None else optim
        loss = Trace_ELBO(num_particles=1) if loss is None else loss
        svi = SVI(self.model, self.guide, optim=optim, loss=loss,
                  cuda_graph=self._use_graph(cuda_graph, optim, loss.num_particles))
        st = get_param_store()

        def postfix():
            out = {}
            if random_instance is not None:
                ri = random_instance
                out["threshold_error"] = "{0}".format(float((st.get("b").cpu() - ri.b.cpu()).abs().mean()))
                if self._model != "irt_1pl":
                    K, J = self.x_feature, self.item_size
                    out["slop_error"] = "{0}".format(float((st.get("a").cpu() - ri.a.cpu()).abs().sum() / (K * J - K * (K - 1) / 2)))
                if self._model in ("irt_3pl", "irt_4pl"):
                    out["guess_error"] = "{0}".format(float((st.get("c").cpu() - ri.c.cpu()).abs().mean()))
                if self._model == "irt_4pl":
                    out["slip_error"] = "{0}".format(float((st.get("d").cpu() - ri.d.cpu()).abs().mean()))
            return out
        return self._fit_loop(svi, optim, max_iter, postfix)


class VIRT(BaseIRT):                    # vi.py:696-723
    def _guide(self, idx, B):
        st, N, K, dev = get_param_store(), self.sample_size, self.x_feature, self.device
        loc_u = st.setdefault("x_local", torch.zeros((N, K)), REAL, dev)
        mu = loc_u.detach() if idx is None else loc_u.detach()[idx]
        if K == 1:
            sc_u = st.setdefault("x_scale", torch.ones((N, K)), POSITIVE, dev)
            sig_all = sc_u.detach().exp()
            sigma = sig_all if idx is None else sig_all[idx]

            def backward(r, w):
                _accum(loc_u, r["d_mu"] * w, idx)
                _accum(sc_u, r["d_sigma"] * sigma * w, idx)          # exp chain rule
            return dict(mu=mu, sigma=sigma), backward, [("x_local", loc_u, None), ("x_scale", sc_u, None)]
        if self.share_cov:
            sc_u = st.setdefault("x_scale", torch.eye(K), LOWER_CHOLESKY, dev)
            chol = sc_u.detach()

            def backward(r, w):
                _accum(loc_u, r["d_mu"] * w, idx)
                _accum(sc_u, r["d_chol"] * w)
        else:
            sc_u = st.setdefault("x_scale", lambda: torch.eye(K).repeat(N, 1, 1), LOWER_CHOLESKY, dev)
            chol = sc_u.detach() if idx is None else sc_u.detach()[idx]

            def backward(r, w):
                _accum(loc_u, r["d_mu"] * w, idx)
                _accum(sc_u, r["d_chol"] * w, idx)
        # the unconstrained (raw) factor goes straight in: LowerCholeskyTransform is fused
        return (dict(mu=mu, chol=chol, chol_raw=True), backward,
                [("x_local", loc_u, None), ("x_scale", sc_u, None)])


class VaeIRT(BaseIRT):                  # vi.py:659-693
    def __init__(self, hidden_dim=64, *args, **kwargs):
        super().__init__(*args, **kwargs)
        enc = NormEncoder if self.x_feature == 1 else MvnEncoder
        self.encoder = enc(self.item_size, self.x_feature, hidden_dim).to(self.device)
        # encoder inside the CUDA path (False: torch autograd end to end).  K = 1: NormEncoder forward / backward
        # kernels (csrc/enc_norm*.cu).  K > 1: MvnEncoder's fc22 + tril scatter + LowerCholeskyTransform fused into
        # the sampler and its backward (csrc/enc_mvn.cu), fc21 with it; fc1 stays in torch.
        self.fused_encoder = True

    def _guide(self, idx, B):
        get_param_store().register_module("encoder", self.encoder)
        sites = [(f"encoder$$${n}", p, None) for n, p in self.encoder.named_parameters()]
        if self.x_feature == 1 and self._norm_encoder_fusable():
            mu, sigma, enc_backward = self._norm_encoder_kernels(idx)

            def backward(r, w):
                enc_backward(r["d_mu"], r["d_sigma"], w)
            return dict(mu=mu, sigma=sigma), backward, sites
        if self.x_feature > 1 and self._mvn_encoder_fusable():
            # both heads of MvnEncoder (fc21: location, fc22: factor) are evaluated inside the call; torch keeps fc1
            enc = self.encoder
            hidden = enc.softplus(enc.fc1(self._encoder_input(idx)))

            def backward(r, w):
                _accum(enc.fc22.weight, r["d_w22"] * w)
                _accum(enc.fc22.bias, r["d_b22"] * w)
                _accum(enc.fc21.weight, r["d_w21"] * w)
                _accum(enc.fc21.bias, r["d_b21"] * w)
                torch.autograd.backward([hidden], [r["d_hidden"] * w])
            return (dict(mu=None, mvn_encoder=(hidden.detach(), enc.fc22.weight.detach(), enc.fc22.bias.detach(),
                                               enc.fc21.weight.detach(), enc.fc21.bias.detach())),
                    backward, sites)
        x_loc, x_sc = self.encoder(self._encoder_input(idx))
        if self.x_feature == 1:
            def backward(r, w):
                torch.autograd.backward([x_loc, x_sc], [r["d_mu"] * w, r["d_sigma"] * w])
            return dict(mu=x_loc.detach(), sigma=x_sc.detach()), backward, sites

        def backward(r, w):
            torch.autograd.backward([x_loc, x_sc], [r["d_mu"] * w, r["d_chol"] * w])
        return dict(mu=x_loc.detach(), chol=x_sc.detach(), chol_packed=True, chol_raw=True), backward, sites


# =========================================================================== CDM
class BaseCDM(BasePsy):                 # vi.py:726-782  (REINFORCE with a Bernoulli guide)
    CDM_FUN = ("dina", "dino")

    def __init__(self, q, model="dina", *args, **kwargs):
        super().__init__(*args, **kwargs)
        if model not in self.CDM_FUN:
            raise KeyError(model)
        self.q = q
        self.attr_size = q.size(0)
        self._model = model
        self._qmask = ops.qmask_from_q(q.to(self.device))

    def _gs(self):
        st, J, dev = get_param_store(), self.item_size, self.device
        g_u = st.setdefault("g", torch.zeros((1, J)) + 0.1, UNIT_INTERVAL, dev)
        s_u = st.setdefault("s", torch.zeros((1, J)) + 0.1, UNIT_INTERVAL, dev)
        return g_u, s_u, torch.sigmoid(g_u.detach()), torch.sigmoid(s_u.detach())

    def _attr_p(self, idx):
        raise NotImplementedError

    def loss_and_grads(self, cfg, data=None, want_grad=True):
        N, B = self.sample_size, min(self.subsample_size, self.sample_size)
        S = cfg.num_particles
        groups = [S] if B >= N else [1] * S
        g_u, s_u, g, s = self._gs()
        sites: List[Site] = [("g", g_u, None), ("s", s_u, None)]
        total = None
        for si in groups:
            idx, Bg, scale = self._plate()
            phi, backward, gsites = self._attr_p(idx)
            r = ops.cdm_score_elbo_grad(self._codes, self._qmask, g, s, phi.detach(), n_attrs=self.attr_size,
                                        n_particles=si, model=self._model, scale=scale, baseline=cfg.baseline,
                                        idx=idx, want_grad=want_grad,
                                        **self._noise_args("u", si, Bg, self.attr_size))
            w = 1.0 / len(groups)
            total = r["loss"] * w if total is None else total + r["loss"] * w
            if want_grad:
                _accum(g_u, r["d_g"] * g * (1 - g) * w)
                _accum(s_u, r["d_s"] * s * (1 - s) * w)
                backward(r, w)
            for site in gsites:
                if site[0] not in {n for n, _, _ in sites}:
                    sites.append(site)
        return self._finish(total, sites)

    def loss(self, cfg, data=None):
        return self.loss_and_grads(cfg, data, want_grad=False)[0]

    def model(self, data):
        raise RuntimeError("evaluated by the fused kernel; call fit() or SVI.step()")

    guide = model

    def _default_loss(self):
        return Trace_ELBO(num_particles=1)

    def fit(self, optim=None, loss=None, max_iter=5000, random_instance=None, evaluate_loss_each_iter=False,
            cuda_graph=None):
        """vi.py:758-782.  The reference re-evaluates the ELBO after every step (vi.py:770) only to
        print it; ``evaluate_loss_each_iter=True`` reproduces that, the default reports the step's loss."""
        optim = Adam({"lr": 1e-3}) if optim is None else optim
        loss = self._default_loss() if loss is None else loss
        svi = SVI(self.model, self.guide, optim=optim, loss=loss,
                  cuda_graph=self._use_graph(cuda_graph, optim, loss.num_particles))
        st = get_param_store()

        def postfix():
            if random_instance is None:
                return {}
            return {"g": "{0}".format(float((st.get("g").cpu() - random_instance.g.cpu()).abs().mean())),
                    "s": "{0}".format(float((st.get("s").cpu() - random_instance.s.cpu()).abs().mean()))}
        return self._fit_loop(svi, optim, max_iter, postfix, evaluate=evaluate_loss_each_iter)


class VCDM(BaseCDM):                    # vi.py:807-816
    def _attr_p(self, idx):
        st = get_param_store()
        p_u = st.setdefault("attr_p", torch.zeros((self.sample_size, self.attr_size)) + 0.5, UNIT_INTERVAL, self.device)
        phi_all = torch.sigmoid(p_u.detach())
        phi = phi_all if idx is None else phi_all[idx]

        def backward(r, w):
            _accum(p_u, r["d_attr_p"] * phi * (1 - phi) * w, idx)
        return phi, backward, [("attr_p", p_u, None)]


class VaeCDM(BaseCDM):                  # vi.py:785-804
    def __init__(self, hidden_dim=64, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.encoder = BinEncoder(self.item_size, self.attr_size, hidden_dim).to(self.device)

    def _attr_p(self, idx):
        get_param_store().register_module("encoder", self.encoder)
        x = self._data_dev if idx is None else self._data_dev[idx]          # vi.py:800: no NaN substitution
        phi = self.encoder(x)

        def backward(r, w):
            torch.autograd.backward([phi], [r["d_attr_p"] * w])
        return phi, backward, [(f"encoder$$${n}", p, None) for n, p in self.encoder.named_parameters()]


class VCCDM(BaseCDM):                   # vi.py:819-863  (exact enumeration of the 2^K patterns)
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.all_attr = self.get_all_attrs()

    def get_all_attrs(self):            # vi.py:825-837: row i = bits of i, LSB first
        K = self.attr_size
        i = torch.arange(2 ** K).unsqueeze(1)
        return ((i >> torch.arange(K).unsqueeze(0)) & 1).float()

    def _enum_extra(self, idx, B):
        return {}, (lambda r, w: None), []

    def loss_and_grads(self, cfg, data=None, want_grad=True):
        N, B = self.sample_size, min(self.subsample_size, self.sample_size)
        S = cfg.num_particles
        hod = isinstance(self, VCHoDina)
        groups = [S] if (B >= N or not hod) else [1] * S
        if not hod and B < N:
            groups = [1] * S
        g_u, s_u, g, s = self._gs()
        sites: List[Site] = [("g", g_u, None), ("s", s_u, None)]
        total = None
        for si in groups:
            idx, Bg, scale = self._plate()
            extra, backward, gsites = self._enum_extra(idx, Bg)
            r = ops.cdm_enum_elbo_grad(self._codes, self._qmask, g, s, n_attrs=self.attr_size,
                                       model="dina" if hod else self._model, scale=scale, idx=idx,
                                       n_particles=si, want_grad=want_grad, **extra,
                                       **self._noise_args("eps" if hod else None, si, Bg, 1))
            w = 1.0 / len(groups)
            total = r["loss"] * w if total is None else total + r["loss"] * w
            if want_grad:
                _accum(g_u, r["d_g"] * g * (1 - g) * w)
                _accum(s_u, r["d_s"] * s * (1 - s) * w)
                backward(r, w)
            for site in gsites:
                if site[0] not in {n for n, _, _ in sites}:
                    sites.append(site)
        return self._finish(total, sites)

    def _default_loss(self):
        return TraceEnum_ELBO(num_particles=1)

    def posterior(self):
        """Per-person posterior over attribute patterns at the current parameters (read-out)."""
        _, _, g, s = self._gs()
        extra, _, _ = self._enum_extra(None, self.sample_size)
        extra = {k: v for k, v in extra.items() if k != "want_posterior"}
        r = ops.cdm_enum_elbo_grad(self._codes, self._qmask, g, s, n_attrs=self.attr_size,
                                   model="dina" if isinstance(self, VCHoDina) else self._model,     # vi.py:905
                                   want_grad=False, want_posterior=True, **extra, **self._noise_args())
        return r["posterior"]


class VaeCCDM(VCCDM):                   # vi.py:866-891
    def __init__(self, hidden_dim=64, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.encoder = SoftmaxEncoder(self.item_size, self.all_attr.size(0), hidden_dim).to(self.device)

    def _enum_extra(self, idx, B):
        """vi.py:878-886: attr_p = encoder(data with NaN -> -1) -- a softmax over the BATCH axis (sic, vi.py:480) --
        becomes the pattern prior through Categorical(probs=attr_p): normalise over patterns, clamp, log (A.4).
        The missing cells stay in as y = -1 (no mask, vi.py:887-891): ``score_missing``."""
        get_param_store().register_module("encoder", self.encoder)
        attr_p = self.encoder(self._encoder_input(idx))
        prior = attr_p / attr_p.sum(-1, keepdim=True)
        eps = torch.finfo(torch.float32).eps
        log_prior = torch.log(prior.clamp(min=eps, max=1.0 - eps))
        scale = self.sample_size / B

        def backward(r, w):
            torch.autograd.backward([log_prior], [r["posterior"] * (-scale * w)])
        return (dict(log_prior=log_prior.detach(), score_missing=True, want_posterior=True), backward,
                [(f"encoder$$${n}", p, None) for n, p in self.encoder.named_parameters()])


class VCHoDina(VCCDM):                  # vi.py:894-958
    def _theta(self, idx):
        st, N, dev = get_param_store(), self.sample_size, self.device
        loc_u = st.setdefault("theta_local", torch.zeros((N, 1)), REAL, dev)
        sc_u = st.setdefault("theta_scale", torch.ones((N, 1)), POSITIVE, dev)
        sig_all = sc_u.detach().exp()
        mu = loc_u.detach() if idx is None else loc_u.detach()[idx]
        sigma = sig_all if idx is None else sig_all[idx]

        def backward(r, w):
            _accum(loc_u, r["d_mu"] * w, idx)
            _accum(sc_u, r["d_sigma"] * sigma * w, idx)
        return mu, sigma, backward, [("theta_local", loc_u, None), ("theta_scale", sc_u, None)]

    def _enum_extra(self, idx, B):
        st, K, dev = get_param_store(), self.attr_size, self.device
        l0_u = st.setdefault("lam0", torch.zeros((1, K)), REAL, dev)
        l1_u = st.setdefault("lam1", torch.ones((1, K)), POSITIVE, dev)
        lam1 = l1_u.detach().exp()
        mu, sigma, bw_theta, tsites = self._theta(idx)

        def backward(r, w):
            _accum(l0_u, r["d_lam0"] * w)
            _accum(l1_u, r["d_lam1"] * lam1 * w)
            bw_theta(r, w)
        return (dict(lam0=l0_u.detach(), lam1=lam1, mu=mu, sigma=sigma), backward,
                [("lam0", l0_u, None), ("lam1", l1_u, None)] + tsites)

    def fit(self, optim=None, loss=None, max_iter=50000, random_instance=None, evaluate_loss_each_iter=False,
            cuda_graph=None):
        optim = Adam({"lr": 5e-3}) if optim is None else optim                  # vi.py:936
        loss = TraceEnum_ELBO(num_particles=1) if loss is None else loss
        svi = SVI(self.model, self.guide, optim=optim, loss=loss,
                  cuda_graph=self._use_graph(cuda_graph, optim, loss.num_particles))
        st = get_param_store()

        def postfix():
            if random_instance is None:
                return {}
            ri = random_instance
            return {k: "{0}".format(float((st.get(k).cpu() - getattr(ri, k).cpu()).abs().mean()))
                    for k in ("g", "s", "lam0", "lam1")}
        return self._fit_loop(svi, optim, max_iter, postfix, evaluate=evaluate_loss_each_iter)


class VaeCHoDina(VCHoDina):             # vi.py:961-981
    def __init__(self, hidden_dim=64, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.encoder = NormEncoder(self.item_size, 1, hidden_dim).to(self.device)
        self.fused_encoder = True       # encoder forward / backward as CUDA kernels (False: torch autograd)

    def _theta(self, idx):
        get_param_store().register_module("encoder", self.encoder)
        sites = [(f"encoder$$${n}", p, None) for n, p in self.encoder.named_parameters()]
        if self._norm_encoder_fusable():
            mu, sigma, enc_backward = self._norm_encoder_kernels(idx)
            return mu, sigma, (lambda r, w: enc_backward(r["d_mu"], r["d_sigma"], w)), sites
        loc, sc = self.encoder(self._encoder_input(idx))

        def backward(r, w):
            torch.autograd.backward([loc, sc], [r["d_mu"] * w, r["d_sigma"] * w])
        return (loc.detach(), sc.detach(), backward,
                [(f"encoder$$${n}", p, None) for n, p in self.encoder.named_parameters()])

"""A minimal functional stand-in for the parts of ``pyro-ppl==1.4.0`` that ``/root/reference/vi.py`` uses,
written so that the reference's OWN ``model`` / ``guide`` methods can be executed unmodified in this container
(Pyro is not installable here: no network, not in /opt/wheelhouse).

TEST INFRASTRUCTURE ONLY -- used by ``tests/golden/make_golden_models.py`` to generate ELBO-level golden
vectors from the reference's code; nothing under ``vipsy_b200/`` imports it.

What is real and what is restated
---------------------------------
* real (executed from /root/reference/vi.py): every model / guide body -- parameter sites, inits and
  constraints, plates and their ``subsample_size`` / ``dim``, the link functions, the NaN masking rule, the
  encoders, the distributions chosen per site, ``to_event`` calls, the pattern table, HO-DINA's prior algebra;
* real (torch): ``torch.distributions`` for every log-probability and ``transform_to`` for every constraint;
* restated here (from the published Pyro 1.4.0 sources): the effect handlers
  (``trace`` / ``replay`` / ``plate`` scaling), ``Trace_ELBO`` (value and surrogate: pathwise for rsample
  sites, score function with ``log_r`` summed to the plate frame otherwise, no baselines) and
  ``TraceEnum_ELBO`` for one parallel-enumerated Categorical site inside one plate (log-sum-exp over the
  enumeration dim of the unscaled enumerated factors, plate scale applied after the contraction).

Sampling is done with explicit base noise (``eps`` for Normal / MultivariateNormal -- the same arithmetic as
``rsample`` --, ``u`` for Bernoulli: value = [u < probs]) which is recorded, so that the oracle and the CUDA
kernels can be fed the identical draws.
"""
from __future__ import annotations

import sys
import types
from contextlib import contextmanager

import torch
import torch.distributions as td
from torch.distributions import constraints, transform_to

td.Distribution.set_default_validate_args(False)        # Pyro 1.4 default; Bernoulli(probs=1.5) needs it (vi.py:753)

_PARAMS = {}            # name -> (unconstrained leaf, constraint)
_S = None               # active recording state


class _State:
    def __init__(self, mode, guide=None, enumerate_sites=()):
        self.mode, self.guide = mode, guide
        self.enumerate_sites = set(enumerate_sites)
        self.sites = {}             # name -> dict
        self.params = {}            # name -> dict(value, unconstrained, free)
        self.plates = {}            # name -> idx
        self.scale = 1.0
        self.free = None
        self.noise = {}             # name -> recorded base noise


def clear_param_store():
    _PARAMS.clear()


def get_param_store():
    return _PARAMS


def param(name, init=None, constraint=constraints.real):
    if name not in _PARAMS:
        v = init() if callable(init) else init
        u = transform_to(constraint).inv(v.detach().clone().float()).clone().requires_grad_(True)
        _PARAMS[name] = (u, constraint)
    u, c = _PARAMS[name]
    val = transform_to(c)(u)
    val.unconstrained = lambda: u
    if _S is not None:
        _S.params[name] = dict(value=val, unconstrained=u, free=_S.free)
    return val


def module(name, nn_module, update_module_params=False):
    for pn, p in nn_module.named_parameters():
        key = f"{name}$$${pn}"
        if key not in _PARAMS:
            _PARAMS[key] = (p, constraints.real)
        if _S is not None:
            _S.params[key] = dict(value=p, unconstrained=p, free=_S.free)
    return nn_module


@contextmanager
def plate(name, size, subsample_size=None, dim=None, **_):
    st = _S
    if st.mode == "model" and st.guide is not None and name in st.guide.plates:
        idx = st.guide.plates[name]                          # replay: the model sees the guide's subsample
    elif subsample_size is None or subsample_size >= size:
        idx = torch.arange(size)
    else:
        idx = torch.randperm(size)[:subsample_size]
    st.plates[name] = idx
    old = st.scale
    st.scale = old * (size / len(idx))
    st.plate_dim = dim
    try:
        yield idx
    finally:
        st.scale = old


class _Dist:
    """torch distribution + Pyro's ``to_event``."""

    def __init__(self, base, event=0):
        self.base, self.event = base, event

    def to_event(self, n=None):
        return _Dist(self.base, self.event + (n or 0))

    @property
    def has_rsample(self):
        return self.base.has_rsample

    def log_prob(self, value):
        lp = self.base.log_prob(value)
        for _ in range(self.event):
            lp = lp.sum(-1)
        return lp


def _wrap(cls):
    def make(*a, **k):
        return _Dist(cls(*a, **k, validate_args=False))
    return make


def _draw(name, d):
    """Explicit-noise sampling with the arithmetic of torch's (r)sample; the noise is recorded."""
    b = d.base
    if isinstance(b, td.Normal):
        eps = torch.randn(b.loc.shape)
        _S.noise[name] = eps
        return b.loc + eps * b.scale
    if isinstance(b, td.MultivariateNormal):
        eps = torch.randn(b.loc.shape)
        _S.noise[name] = eps
        return b.loc + torch.matmul(b.scale_tril, eps.unsqueeze(-1)).squeeze(-1)
    if isinstance(b, td.Bernoulli):
        u = torch.rand(b.probs.shape)
        _S.noise[name] = u
        return (u < b.probs).to(b.probs.dtype)
    raise NotImplementedError(type(b))


def sample(name, d, obs=None, infer=None):
    st = _S
    enumerated = False
    if obs is not None:
        value = obs
    elif st.mode == "model" and st.guide is not None and name in st.guide.sites:
        value = st.guide.sites[name]["value"]
    elif name in st.enumerate_sites:
        value = d.base.enumerate_support(expand=False)       # (C, 1): enumeration dim left of the plate dim
        enumerated = True
    else:
        value = _draw(name, d).detach() if not d.has_rsample else _draw(name, d)
    lp = d.log_prob(value)
    st.sites[name] = dict(value=value, log_prob=lp, scale=st.scale, observed=obs is not None, fn=d,
                          rsample=d.has_rsample, enumerated=enumerated)
    return value


# --------------------------------------------------------------------------- poutine / infer / optim shims
class Messenger:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        self._old = _S.free
        self._process_message_free()
        return self

    def __exit__(self, *a):
        _S.free = self._old

    def _process_message_free(self):
        msg = {"free": _S.free}
        self._process_message(msg)
        _S.free = msg["free"]


def config_enumerate(fn):
    fn._enumerate = True
    return fn


def _run(fn, mode, guide=None, enumerate_sites=(), args=()):
    global _S
    st = _State(mode, guide, enumerate_sites)
    old, _S = _S, st
    try:
        fn(*args)
    finally:
        _S = old
    return st


def _sum_to_plate(x, plate_ndim):
    while x.dim() > plate_ndim:
        x = x.sum(-1)
    return x


class Trace_ELBO:
    enumerate_sites = ()

    def __init__(self, num_particles=1, **_):
        self.num_particles = num_particles
        self.records = []           # per particle: dict(idx, noise, sites)

    def _particle(self, model, guide, args):
        g = _run(guide, "guide", args=args)
        m = _run(model, "model", guide=g, enumerate_sites=self.enumerate_sites, args=args)
        return g, m

    def _elbo(self, g, m):
        elbo = 0.0
        surrogate = 0.0
        for s in m.sites.values():
            t = (s["scale"] * s["log_prob"]).sum()
            elbo = elbo + t
            surrogate = surrogate + t
        log_r = None
        for name, s in g.sites.items():
            lq = s["log_prob"]
            elbo = elbo - (s["scale"] * lq).sum()
            if s["rsample"]:
                surrogate = surrogate - (s["scale"] * lq).sum()            # entropy term; score function is 0
            else:
                if log_r is None:                                            # _compute_log_r, summed to the plate frame
                    nd = lq.dim()
                    log_r = 0.0
                    for mn, ms in m.sites.items():
                        term = ms["scale"] * ms["log_prob"]
                        if not ms["observed"]:
                            term = term - g.sites[mn]["scale"] * g.sites[mn]["log_prob"]
                        log_r = log_r + _sum_to_plate(term.detach(), nd)
                surrogate = surrogate + (log_r * lq).sum()                  # score function unscaled (A.1)
        return elbo, surrogate

    def loss_and_grads(self, model, guide, *args):
        loss = 0.0
        self.records = []
        for _ in range(self.num_particles):
            g, m = self._particle(model, guide, args)
            elbo, surrogate = self._elbo(g, m)
            (-surrogate / self.num_particles).backward()
            loss += -float(elbo.detach()) / self.num_particles
            self.records.append(dict(idx={k: v.clone() for k, v in {**m.plates, **g.plates}.items()},
                                     noise={k: v.clone() for k, v in g.noise.items()},
                                     params={**g.params, **m.params}))
        return loss


class TraceEnum_ELBO(Trace_ELBO):
    enumerate_sites = ("attr_idx",)

    def _elbo(self, g, m):
        # one parallel-enumerated site inside one plate: the factors that carry the enumeration dim are summed
        # unscaled, contracted by log-sum-exp over that dim, and the plate scale is applied after (A.3)
        enum_scale = None
        joint = 0.0
        elbo = 0.0
        for s in m.sites.values():
            lp = s["log_prob"]
            if s["enumerated"] or lp.dim() >= 2 and lp.shape[0] > 1 and self._has_enum_dim(m, lp):
                joint = joint + lp
                enum_scale = s["scale"]
            else:
                elbo = elbo + (s["scale"] * lp).sum()
        elbo = elbo + enum_scale * torch.logsumexp(joint, dim=0).sum()
        for s in g.sites.values():
            assert s["rsample"]
            elbo = elbo - (s["scale"] * s["log_prob"]).sum()
        return elbo, elbo                                   # every guide site is reparameterised: surrogate = elbo

    @staticmethod
    def _has_enum_dim(m, lp):
        C = next(s["value"].shape[0] for s in m.sites.values() if s["enumerated"])
        return lp.dim() == 2 and lp.shape[0] == C


def install():
    """Register this module under the names ``vi.py`` imports."""
    me = sys.modules[__name__]

    def mod(name):
        mm = types.ModuleType(name)
        sys.modules[name] = mm
        return mm

    pyro = mod("pyro")
    for n in ("param", "module", "plate", "sample", "clear_param_store", "get_param_store"):
        setattr(pyro, n, getattr(me, n))
    d = mod("pyro.distributions")
    for n in ("Bernoulli", "Categorical", "Normal", "MultivariateNormal"):
        setattr(d, n, _wrap(getattr(td, n)))
    d.constraints = constraints
    sys.modules["pyro.distributions.constraints"] = constraints
    pyro.distributions = d
    pout = mod("pyro.poutine")
    msg = mod("pyro.poutine.messenger")
    msg.Messenger = Messenger
    pout.messenger = msg
    pyro.poutine = pout
    infer = mod("pyro.infer")

    class _Unused:
        def __init__(self, *a, **k):
            pass

    infer.SVI, infer.Predictive, infer.Importance = _Unused, _Unused, _Unused
    infer.Trace_ELBO, infer.TraceEnum_ELBO, infer.config_enumerate = Trace_ELBO, TraceEnum_ELBO, config_enumerate
    util = mod("pyro.infer.util")
    util.torch_item = lambda x: x if isinstance(x, float) else x.item()
    infer.util = util
    optim = mod("pyro.optim")
    optim.Adam, optim.PyroLRScheduler = _Unused, _Unused
    pyro.infer, pyro.optim = infer, optim
    return pyro

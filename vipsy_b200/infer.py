"""Host-side mirror of the Pyro names vipsy's ``fit`` API relies on:
the global param store with constraint transforms, per-parameter Adam with dict or
callable arguments, MultiStepLR, the ELBO configuration objects and the patched
``SVI.step`` of vi.py:503-516 (free-mask multiply, optimiser step, zero grads).

The arithmetic of ``loss_and_grads`` itself is not here: it is the fused CUDA path
(``vipsy_b200.ops``), which the estimator classes in ``vipsy_b200.models`` call.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Tuple, Union

import torch

# --------------------------------------------------------------------------- constraints
REAL, POSITIVE, UNIT_INTERVAL, LOWER_CHOLESKY = "real", "positive", "unit_interval", "lower_cholesky"


class constraints:                      # noqa: N801  (mirrors pyro.distributions.constraints)
    real = REAL
    positive = POSITIVE
    unit_interval = UNIT_INTERVAL
    lower_cholesky = LOWER_CHOLESKY

    @staticmethod
    def interval(lo, hi):
        if (lo, hi) != (0, 1):
            raise NotImplementedError("only interval(0, 1) is used by vipsy (vi.py:748-749)")
        return UNIT_INTERVAL


def transform(constraint: str, u: torch.Tensor) -> torch.Tensor:
    """``transform_to(constraint)(unconstrained)``."""
    if constraint == REAL:
        return u
    if constraint == POSITIVE:
        return u.exp()
    if constraint == UNIT_INTERVAL:
        return torch.sigmoid(u)
    if constraint == LOWER_CHOLESKY:
        return u.tril(-1) + torch.diag_embed(torch.diagonal(u, dim1=-2, dim2=-1).exp())
    raise ValueError(constraint)


def inverse_transform(constraint: str, v: torch.Tensor) -> torch.Tensor:
    if constraint == REAL:
        return v.clone()
    if constraint == POSITIVE:
        return v.log()
    if constraint == UNIT_INTERVAL:
        return torch.log(v) - torch.log1p(-v)
    if constraint == LOWER_CHOLESKY:
        return v.tril(-1) + torch.diag_embed(torch.diagonal(v, dim1=-2, dim2=-1).log())
    raise ValueError(constraint)


# --------------------------------------------------------------------------- param store
class ParamStore:
    def __init__(self):
        self._u: Dict[str, torch.Tensor] = {}
        self._c: Dict[str, str] = {}

    def clear(self):
        self._u.clear()
        self._c.clear()

    def __contains__(self, name):
        return name in self._u

    def names(self):
        return list(self._u)

    def setdefault(self, name: str, init, constraint: str = REAL, device=None) -> torch.Tensor:
        """First call stores transform^-1(init) as an unconstrained leaf; later calls ignore init."""
        if name not in self._u:
            if init is None:
                raise KeyError(f"param {name!r} has not been initialised")
            v = init() if callable(init) else init
            v = v.detach().to(device=device, dtype=torch.float32)
            self._u[name] = inverse_transform(constraint, v).contiguous().requires_grad_(True)
            self._c[name] = constraint
        return self._u[name]

    def unconstrained(self, name: str) -> torch.Tensor:
        return self._u[name]

    def constraint(self, name: str) -> str:
        return self._c[name]

    def get(self, name: str) -> torch.Tensor:
        """Constrained value (what ``pyro.param(name)`` returns), detached."""
        with torch.no_grad():
            return transform(self._c[name], self._u[name])

    def register_module(self, prefix: str, module: torch.nn.Module):
        for n, p in module.named_parameters():
            key = f"{prefix}$$${n}"
            if key not in self._u:
                self._u[key] = p
                self._c[key] = REAL


_STORE = ParamStore()


def get_param_store() -> ParamStore:
    return _STORE


def clear_param_store():
    _STORE.clear()


def param(name: str, init=None, constraint: str = REAL) -> torch.Tensor:
    """``pyro.param``: constrained value of a stored parameter (detached read-back)."""
    if name not in _STORE:
        _STORE.setdefault(name, init, constraint)
    return _STORE.get(name)


# --------------------------------------------------------------------------- optimisers
class PyroOptim:
    """One ``torch.optim`` instance per parameter tensor, created lazily."""

    def __init__(self, optim_constructor, optim_args: Union[dict, Callable], clip_args=None):
        self.ctor, self.args = optim_constructor, optim_args
        self.optim_objs: Dict[torch.Tensor, torch.optim.Optimizer] = {}
        self._names: Dict[torch.Tensor, str] = {}

    def _args_for(self, name: str) -> dict:
        if callable(self.args):
            module = name.split("$$$")[0]
            pname = name.split("$$$")[-1]
            return dict(self.args(module, pname))
        return dict(self.args)

    def _make(self, p, name):
        return self.ctor([p], **self._args_for(name))

    def __call__(self, params: Iterable[Tuple[str, torch.Tensor]]):
        for name, p in params:
            if p not in self.optim_objs:
                self.optim_objs[p] = self._make(p, name)
                self._names[p] = name
            if p.grad is not None:
                self.optim_objs[p].step()


class Adam(PyroOptim):
    def __init__(self, optim_args, clip_args=None):
        super().__init__(torch.optim.Adam, optim_args, clip_args)


class PyroLRScheduler(PyroOptim):
    """``pyro.optim.PyroLRScheduler``: wraps every per-parameter optimiser in a scheduler;
    ``fit`` calls ``.step()`` once per iteration (vi.py:639-640), so milestones are iterations."""

    def __init__(self, scheduler_constructor, optim_args: dict, clip_args=None):
        a = dict(optim_args)
        self.sched_ctor = scheduler_constructor
        inner = a.pop("optimizer")
        inner_args = a.pop("optim_args")
        self.kwargs = a
        super().__init__(inner, inner_args, clip_args)
        self.schedulers: Dict[torch.Tensor, object] = {}

    def _make(self, p, name):
        opt = super()._make(p, name)
        self.schedulers[p] = self.sched_ctor(opt, **self.kwargs)
        return opt

    def step(self, *args, **kwargs):
        for s in self.schedulers.values():
            s.step(*args, **kwargs)


class MultiStepLR(PyroLRScheduler):
    def __init__(self, optim_args, clip_args=None):
        super().__init__(torch.optim.lr_scheduler.MultiStepLR, optim_args, clip_args)


class StepLR(PyroLRScheduler):
    def __init__(self, optim_args, clip_args=None):
        super().__init__(torch.optim.lr_scheduler.StepLR, optim_args, clip_args)


# --------------------------------------------------------------------------- ELBO objects
class Trace_ELBO:                                           # noqa: N801
    """``pyro.infer.Trace_ELBO(num_particles=S)`` routed to the fused B200 path.

    ``estimator``: "pathwise" is what the reference executes for its Normal / MVN guides;
    "score" is the REINFORCE estimator its README describes (SURVEY Finding 1), with an
    optional fixed ``baseline``.  Bernoulli-guide CDMs are always score-function."""
    enumerate = False

    def __init__(self, num_particles: int = 1, estimator: str = "pathwise", baseline: float = 0.0, **_ignored):
        self.num_particles = int(num_particles)
        self.estimator, self.baseline = estimator, float(baseline)


class TraceEnum_ELBO(Trace_ELBO):                           # noqa: N801
    enumerate = True


# --------------------------------------------------------------------------- SVI
class SVI:
    """The reference's patched SVI (vi.py:503-516): loss_and_grads, multiply ``param.grad`` by
    the site's ``free`` mask, optimiser step, zero grads, return the loss as a float."""

    def __init__(self, model, guide, optim: PyroOptim, loss: Trace_ELBO):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss
        self.owner = model.__self__

    def step(self, *args, **kwargs) -> float:
        loss, sites = self.owner.loss_and_grads(self.loss, *args, **kwargs)
        params: List[Tuple[str, torch.Tensor]] = []
        for name, p, free in sites:
            if free is not None and p.grad is not None:
                p.grad = free.to(p.grad.dtype) * p.grad          # vi.py:511-512
            params.append((name, p))
        self.optim(params)
        for _, p in params:                                      # pyro.infer.util.zero_grads
            p.grad = None
        return float(loss)

    def evaluate_loss(self, *args, **kwargs) -> float:
        with torch.no_grad():
            return float(self.owner.loss(self.loss, *args, **kwargs))

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
    """One ``torch.optim`` instance per parameter tensor, created lazily.

    ``fused_step`` is the device-resident form of the same update for ``torch.optim.Adam`` with default
    options: one multi-tensor kernel (``ops.adam_step``) applies the free mask, the Adam update and the
    gradient zeroing for every parameter of the step; the step count and the learning rates live in device
    memory, so the whole SVI iteration can be captured in a CUDA graph.  The per-parameter ``torch.optim``
    objects still exist (``optim_objs``; schedulers step them), but only as the source of the hyper-parameters."""

    def __init__(self, optim_constructor, optim_args: Union[dict, Callable], clip_args=None):
        self.ctor, self.args = optim_constructor, optim_args
        # pyro.optim.PyroOptim: {'clip_norm': x} and / or {'clip_value': y}, a dict or a callable
        # (module_name, param_name) -> dict; applied per parameter tensor right before its optimiser step
        self.clip_args = clip_args
        self.optim_objs: Dict[torch.Tensor, torch.optim.Optimizer] = {}
        self._names: Dict[torch.Tensor, str] = {}
        self._fused: Dict[torch.Tensor, dict] = {}
        self._step_dev: Optional[torch.Tensor] = None
        self._steps_host = 0

    # ---- device-resident Adam ------------------------------------------------------------------------------
    _FUSED_DEFAULTS = (("weight_decay", 0), ("amsgrad", False), ("maximize", False))

    @classmethod
    def _fusable_args(cls, a: dict) -> bool:
        return all(a.get(k, dflt) == dflt for k, dflt in cls._FUSED_DEFAULTS)

    def supports_fused(self, names: Optional[Iterable[str]] = None) -> bool:
        """Can ``fused_step`` reproduce this optimiser?  torch.optim.Adam with default options (no weight decay,
        amsgrad or maximize) and no gradient clipping.  Per-parameter callable arguments can only be resolved once
        the parameter names are known: pass ``names`` (SVI does, on its first step)."""
        if self.ctor is not torch.optim.Adam or self.clip_args is not None:
            return False
        if not callable(self.args):
            return self._fusable_args(self.args)
        return all(self._fusable_args(self._args_for(n)) for n in (names or ()))

    def _fused_state(self, name: str, p: torch.Tensor) -> dict:
        st = self._fused.get(p)
        if st is None:
            if p not in self.optim_objs:
                self.optim_objs[p] = self._make(p, name)
                self._names[p] = name
            g = self.optim_objs[p].param_groups[0]
            if g.get("weight_decay", 0) != 0 or g.get("amsgrad", False) or g.get("maximize", False):
                raise NotImplementedError("fused Adam covers torch.optim.Adam's default options only")
            st = dict(param=p.data, exp_avg=torch.zeros_like(p.data), exp_avg_sq=torch.zeros_like(p.data),
                      lr_dev=torch.full((1,), float(g["lr"]), dtype=torch.float32, device=p.device),
                      lr=float(g["lr"]), beta1=float(g["betas"][0]), beta2=float(g["betas"][1]), eps=float(g["eps"]),
                      step_offset=self._steps_host, group=g)
            self._fused[p] = st
        return st

    def fused_step(self, sites):
        """sites: (name, leaf, free mask or None) with ``leaf.grad`` set.  Stream-ordered, no host sync."""
        from . import ops
        live = [(n, p, f) for n, p, f in sites if p.grad is not None]
        if not live:
            return
        dev = live[0][1].device
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        tensors = []
        for name, p, free in live:
            st = dict(self._fused_state(name, p))
            st["grad"] = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if free is not None:
                fm = st.get("free_u8")
                if fm is None:
                    fm = free.to(device=p.device, dtype=torch.uint8).expand_as(p).contiguous()
                    self._fused[p]["free_u8"] = fm
                st["free_mask"] = fm
            tensors.append(st)
        self._step_dev.add_(1)
        self._steps_host += 1          # NOTE: under graph replay the host count is advanced by SVI
        ops.adam_step(tensors, self._step_dev, zero_grad=False)

    def sync_lr(self):
        """Copy the (possibly scheduler-modified) learning rates of the torch optimizer objects to the device."""
        for p, st in self._fused.items():
            lr = float(st["group"]["lr"])
            if lr != st["lr"]:
                st["lr"] = lr
                st["lr_dev"].fill_(lr)

    def _args_for(self, name: str) -> dict:
        if callable(self.args):
            module = name.split("$$$")[0]
            pname = name.split("$$$")[-1]
            return dict(self.args(module, pname))
        return dict(self.args)

    def _make(self, p, name):
        return self.ctor([p], **self._args_for(name))

    def _clip(self, name: str, p: torch.Tensor):
        if self.clip_args is None:
            return
        if callable(self.clip_args):
            ca = dict(self.clip_args(name.split("$$$")[0], name.split("$$$")[-1]))
        else:
            ca = dict(self.clip_args)
        unknown = set(ca) - {"clip_norm", "clip_value"}
        if unknown:
            raise ValueError(f"clip_args: unknown keys {sorted(unknown)} (clip_norm, clip_value)")
        if ca.get("clip_norm") is not None:
            torch.nn.utils.clip_grad_norm_([p], ca["clip_norm"])
        if ca.get("clip_value") is not None:
            torch.nn.utils.clip_grad_value_([p], ca["clip_value"])

    def __call__(self, params: Iterable[Tuple[str, torch.Tensor]]):
        for name, p in params:
            if p not in self.optim_objs:
                self.optim_objs[p] = self._make(p, name)
                self._names[p] = name
            if p.grad is not None:
                self._clip(name, p)
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
        import warnings
        with warnings.catch_warnings():          # fused mode never calls the torch optimizer's own step()
            warnings.simplefilter("ignore", UserWarning)
            for s in self.schedulers.values():
                s.step(*args, **kwargs)
        self.sync_lr()


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
    the site's ``free`` mask, optimiser step, zero grads, return the loss as a float.

    ``cuda_graph=True``: the whole iteration -- minibatch draw, noise, encoder forward, fused
    ELBO + gradient kernels, constraint chain rules, encoder backward, free mask, Adam, gradient zeroing --
    runs on the device with no host round trip and is captured once into a CUDA graph; every later ``step``
    is one graph launch.  The reference's default usage (subsample_size ~ 100, tens of thousands of
    iterations) is bound by launch and interpreter overhead, not by arithmetic; this removes it.  Needs
    ``torch.optim.Adam`` with default options (what ``pyro.optim.Adam`` wraps)."""

    WARMUP = 2          # eager iterations before capture (lazy parameter / optimiser / workspace creation)

    _NEEDS = ("the device-resident iteration needs torch.optim.Adam with default options (no weight_decay / amsgrad "
              "/ maximize) and no clip_args; pass cuda_graph=False for the eager torch.optim path")

    def __init__(self, model, guide, optim: PyroOptim, loss: Trace_ELBO, cuda_graph: Optional[bool] = False):
        """cuda_graph: True = device-resident captured iteration or an error; False = eager torch.optim objects;
        None = captured when the optimiser allows it, else eager (decided at the latest on the first step, when
        per-parameter callable optimiser arguments can be resolved)."""
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss
        self.owner = model.__self__
        self._strict = cuda_graph is True
        self.cuda_graph = (cuda_graph is None or bool(cuda_graph)) and (optim.supports_fused() or callable(optim.args))
        if self._strict and not self.cuda_graph:
            raise NotImplementedError(self._NEEDS)
        self._fused_checked = False
        self._graph = None
        self._static_loss = None
        self._eager_steps = 0

    # ---- eager form, torch.optim objects (the reference's order of operations) -------------------------------
    def _step_eager(self, *args, **kwargs) -> torch.Tensor:
        return self._finish_eager(*self.owner.loss_and_grads(self.loss, *args, **kwargs))

    def _finish_eager(self, loss, sites) -> torch.Tensor:
        params: List[Tuple[str, torch.Tensor]] = []
        for name, p, free in sites:
            if free is not None and p.grad is not None:
                p.grad = free.to(p.grad.dtype) * p.grad          # vi.py:511-512
            params.append((name, p))
        self.optim(params)
        for _, p in params:                                      # pyro.infer.util.zero_grads
            p.grad = None
        return loss

    # ---- device-resident form ----------------------------------------------------------------------------------
    def _step_device(self, *args, **kwargs) -> torch.Tensor:
        loss, sites = self.owner.loss_and_grads(self.loss, *args, **kwargs)
        if not self._fused_checked:                              # first (eager warm-up) step: names are known now
            self._fused_checked = True
            if not self.optim.supports_fused([n for n, _, _ in sites]):
                if self._strict:
                    raise NotImplementedError(self._NEEDS)
                self.cuda_graph = False                          # auto mode: this and every later step run eagerly
                return self._finish_eager(loss, sites)
        self.optim.fused_step(sites)                             # free mask + Adam in one launch
        for _, p, _ in sites:
            p.grad = None
        return loss

    def step_async(self, *args, **kwargs) -> torch.Tensor:
        """One iteration; returns the loss as a device tensor without synchronising."""
        if not self.cuda_graph:
            return self._step_eager(*args, **kwargs)
        owner = self.owner
        owner._device_rng = True                                 # noise / minibatch counters in device memory
        if self._graph is None:
            if self._eager_steps < self.WARMUP:
                self._eager_steps += 1
                return self._step_device(*args, **kwargs)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            host_steps, calls0 = self.optim._steps_host, owner._calls
            with torch.cuda.graph(g):
                self._static_loss = self._step_device(*args, **kwargs)
            self.optim._steps_host = host_steps                  # capture does not execute
            self._calls_per_iter, owner._calls = owner._calls - calls0, calls0
            self._graph = g
        self._graph.replay()
        self.optim._steps_host += 1
        owner._calls += self._calls_per_iter                     # the noise calls the replay just made on the device
        return self._static_loss

    def step(self, *args, **kwargs) -> float:
        return float(self.step_async(*args, **kwargs))

    def evaluate_loss(self, *args, **kwargs) -> float:
        with torch.no_grad():
            return float(self.owner.loss(self.loss, *args, **kwargs))

"""CPU oracle for vipsy's per-iteration ELBO + gradient evaluation.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vipsy_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker
(or the reported CPU baseline), never as the thing shipped.

What it restates (all citations into /root/reference):

* link functions ``irt_1pl..irt_4pl``            vi.py:22-66
* ``dina`` / ``dino`` (incl. the DINO ordering quirk) vi.py:69-100
* the pattern table ``VCCDM.get_all_attrs``        vi.py:825-837
* ``BaseIRT.model`` + ``_get_p_data`` (NaN rule)   vi.py:574-625
* ``VIRT.guide`` / ``VaeIRT.guide``                vi.py:673-723
* ``BaseCDM.model`` + ``VCDM.guide`` (REINFORCE)   vi.py:745-756, 807-816
* ``VCCDM.model`` / ``VaeCCDM.model``              vi.py:839-857, 872-891
* ``VCHoDina.model`` / ``.guide``                  vi.py:897-934
* the ``free`` mask of the patched ``SVI.step``    vi.py:503-516

plus the semantics of the un-vendored dependency that owns the arithmetic,
``pyro-ppl==1.4.0`` (requirements.txt:1) ``Trace_ELBO`` / ``TraceEnum_ELBO`` and
``torch.distributions`` (Bernoulli-by-probs clamp, Categorical floor), as
written down in SURVEY.md Appendix A.

PARITY STATUS.  The reference's own tests contain no assertion, golden vector or fixed seed for this
path (test.py has 0 ``assert``; seeds come from ``random.random()``, test.py:159) and Pyro itself is not
installable here, so nothing the reference ships pins the numbers.  What pins this oracle instead:

* ``tests/golden/make_golden.py``: the link functions, DINA/DINO eta tables, the pattern table and the encoders'
  output layout against the reference's own ``vi.py`` functions (imported under a ``pyro`` stub);
* ``tests/golden/make_golden_models.py``: ELBO-level vectors (loss + the gradient of every unconstrained
  parameter, 18 cases over all eight estimator classes, with NaNs, subsampling, several particles) produced by
  the reference's OWN ``model`` / ``guide`` methods, executed unmodified under ``tests/golden/minipyro.py`` --
  a functional stand-in for the pieces of Pyro 1.4.0 they call.  The oracle reproduces all of them
  (``tests/test_oracle_reference_models.py``), and so do the CUDA kernels directly
  (``tests/test_kernels_vs_reference_models_gpu.py``).

What remains a restatement -- and is therefore still "parity unpinned" in the strict sense -- is the
stand-in's own semantics: Pyro's ``trace`` / ``replay`` / ``plate`` scaling and the ``Trace_ELBO`` /
``TraceEnum_ELBO`` estimators (written from the published 1.4.0 sources).  The
model code, distributions, constraints, masks and link functions on both sides of that boundary are the
reference's and torch's own.

Everything takes *constrained* parameter values (what ``pyro.param`` returns)
and returns gradients w.r.t. those values, exactly like the CUDA C-ABI does;
the chain rule through sigmoid/exp/lower_cholesky stays in torch on both sides.
``dtype=torch.float32`` reproduces the reference arithmetic; ``torch.float64``
is the arbiter the 1e-5 tolerance is measured against (same formulas, fp32
clamp constants).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

EPS32 = float(torch.finfo(torch.float32).eps)  # 1.1920929e-07, torch clamp_probs eps (A.4)
LOG_EPS32 = math.log(EPS32)                     # -15.942385
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)

IRT_MODELS = ("irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl")


# --------------------------------------------------------------------------
# torch.distributions numerics
# --------------------------------------------------------------------------
def bernoulli_logprob(p: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """``Bernoulli(probs=p).log_prob(y)`` with argument validation off.

    torch: probs -> clamp(eps, 1-eps) -> logits = log(p) - log1p(-p) ->
    -binary_cross_entropy_with_logits(logits, y).  The clamp constant is the
    *fp32* eps in every dtype so that the fp64 arbiter computes the same
    function as the fp32 reference.
    """
    pc = p.clamp(min=EPS32, max=1.0 - EPS32)
    logits = torch.log(pc) - torch.log1p(-pc)
    # -BCEWithLogits = y*z - softplus(z) = (y-1)*z + logsigmoid(z)   (torch 2.x form)
    return (y - 1.0) * logits + torch.nn.functional.logsigmoid(logits)


def categorical_logits(w: torch.Tensor) -> torch.Tensor:
    """``Categorical(probs=w).logits``: normalise, floor at eps, log (A.4)."""
    p = w / w.sum(-1, keepdim=True)
    return torch.log(p.clamp(min=EPS32, max=1.0 - EPS32))


def normal_logprob(x, loc, scale):
    return -((x - loc) ** 2) / (2.0 * scale ** 2) - torch.log(scale) - HALF_LOG_2PI


def mvn_tril_logprob(x, loc, L):
    """MultivariateNormal(loc, scale_tril=L).log_prob(x) via triangular solve (A.4)."""
    diff = (x - loc).unsqueeze(-1)
    if L.dim() == 2:
        Lb = L.expand(diff.shape[0], *L.shape)
    else:
        Lb = L
    z = torch.linalg.solve_triangular(Lb, diff, upper=False).squeeze(-1)
    half_log_det = torch.diagonal(Lb, dim1=-2, dim2=-1).log().sum(-1)
    k = x.shape[-1]
    return -0.5 * (z ** 2).sum(-1) - half_log_det - k * HALF_LOG_2PI


# --------------------------------------------------------------------------
# L2: psychometric link functions
# --------------------------------------------------------------------------
def irt_prob(model: str, x, a, b, c=None, d=None, D: float = 1.0):
    """Response probability matrix B x J (vi.py:22-66)."""
    if model == "irt_1pl":
        return torch.sigmoid(D * (x + b))
    s = torch.sigmoid(D * (x @ a + b))
    if model == "irt_2pl":
        return s
    if model == "irt_3pl":
        return c + (1.0 - c) * s
    if model == "irt_4pl":
        return c + (d - c) * s
    raise ValueError(f"unknown IRT model {model!r}")


def all_patterns(K: int, dtype=torch.float32) -> torch.Tensor:
    """2^K x K table, row i = binary digits of i, LSB first (vi.py:825-837)."""
    idx = torch.arange(2 ** K).unsqueeze(1)
    return ((idx >> torch.arange(K).unsqueeze(0)) & 1).to(dtype)


def cdm_eta(model: str, attr: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Ideal-response indicator, exactly 0/1 (vi.py:78-81 / 95-98).

    DINA: eta = [attr.q == sum q^2].  DINO as executed: the reference first sets
    every entry *below* the requirement count to 1 and then every entry *equal*
    to the count to 0 -- in place, so a freshly written 1 is wiped again when
    the count itself is 1 (and an untouched 0 is wiped when it is 0).
    """
    need = (q ** 2).sum(0)
    if model == "dina":
        have = attr @ q
        return (have == need).to(attr.dtype)
    if model == "dino":
        lack = (1.0 - attr) @ q
        eta = torch.where(lack < need, torch.ones_like(lack), lack)
        eta = torch.where(eta == need, torch.zeros_like(eta), eta)
        return eta
    raise ValueError(f"unknown CDM model {model!r}")


def cdm_prob(model: str, attr, q, g, s):
    """p = (1-s)^eta * g^(1-eta) (vi.py:82 / 99)."""
    eta = cdm_eta(model, attr, q)
    return (1.0 - s) ** eta * g ** (1.0 - eta)


def tril_from_packed(packed: torch.Tensor, K: int) -> torch.Tensor:
    """Scatter a (..., K(K+1)/2) vector row-major into (..., K, K) (vi.py:452-454)."""
    idx = torch.tril_indices(K, K)
    out = packed.new_zeros(packed.shape[:-1] + (K, K))
    out[..., idx[0], idx[1]] = packed
    return out


def lower_cholesky(raw: torch.Tensor) -> torch.Tensor:
    """``LowerCholeskyTransform``: tril(raw,-1) + diag(exp(diag raw)) (A.4)."""
    return raw.tril(-1) + torch.diag_embed(torch.diagonal(raw, dim1=-2, dim2=-1).exp())


def _mask_nan(p, y):
    """vi.py:621-624: where y is NaN, y := 0 and p := 0."""
    nan = torch.isnan(y)
    if nan.any():
        y = torch.where(nan, torch.zeros_like(y), y)
        p = torch.where(nan, torch.zeros_like(p), p)
    return p, y, nan


def norm_encoder(y, w1, b1, w21, b21, w22, b22, dtype=torch.float64):
    """``NormEncoder.forward`` (vi.py:417-435) on the guide's input (vi.py:679-682: NaN -> -1):
    hidden = softplus(fc1(x)); returns (fc21(hidden), exp(fc22(hidden))), each B x 1.  Differentiable in the six
    weight tensors (pass leaves to get their gradients by autograd)."""
    x = y.to(dtype)
    x = torch.where(torch.isnan(x), torch.full_like(x, -1.0), x)
    hidden = torch.nn.functional.softplus(x @ w1.to(dtype).t() + b1.to(dtype))
    return hidden @ w21.to(dtype).t() + b21.to(dtype), torch.exp(hidden @ w22.to(dtype).t() + b22.to(dtype))


def _leaf(t, dtype):
    return None if t is None else t.detach().to(dtype).clone().requires_grad_(True)


def _grad(t):
    if t is None:
        return None
    return torch.zeros_like(t) if t.grad is None else t.grad.detach().clone()


# --------------------------------------------------------------------------
# IRT: Trace_ELBO with Normal / MultivariateNormal guide  (Appendix A.1/A.2)
# --------------------------------------------------------------------------
def irt_elbo_grad(
    y: torch.Tensor,                    # B x J, NaN = missing (rows already gathered by idx)
    mu: torch.Tensor,                   # B x K
    scale_param: torch.Tensor,          # K==1: sigma B x 1;  K>1: L (K x K shared | B x K x K)
    eps: torch.Tensor,                  # S x B x K injected standard-normal noise
    a: Optional[torch.Tensor],          # K x J (None for 1PL)
    b: torch.Tensor,                    # 1 x J
    c: Optional[torch.Tensor] = None,   # 1 x J
    d: Optional[torch.Tensor] = None,   # 1 x J
    *,
    model: str = "irt_2pl",
    D: float = 1.0,
    scale: float = 1.0,                 # N / B of the `data` plate
    estimator: str = "pathwise",        # "pathwise" = what fit() executes; "score" = REINFORCE
    baseline: float = 0.0,
    a_free: Optional[torch.Tensor] = None,
    chol_raw: bool = False,             # scale_param holds raw (unconstrained) Cholesky entries
    dtype=torch.float64,
) -> Dict[str, Optional[torch.Tensor]]:
    """One ``Trace_ELBO(num_particles=S).loss_and_grads`` for BaseIRT.model with
    the VIRT / VaeIRT guide, on explicit noise.  Returns loss and d(loss)/d(input).
    """
    S, B, K = eps.shape
    y = y.to(dtype)
    mu_, sp_ = _leaf(mu, dtype), _leaf(scale_param, dtype)
    a_, b_, c_, d_ = _leaf(a, dtype), _leaf(b, dtype), _leaf(c, dtype), _leaf(d, dtype)
    eps = eps.to(dtype)
    loss_value = 0.0
    total = 0.0
    for s in range(S):
        if K == 1:
            sig = sp_
            x = mu_ + sig * eps[s]
        else:
            L = lower_cholesky(sp_) if chol_raw else sp_.tril()
            x = mu_ + (L @ eps[s].unsqueeze(-1)).squeeze(-1) if L.dim() == 3 else mu_ + eps[s] @ L.T
        if estimator == "score":
            x = x.detach()
        # guide site 'x' (vi.py:705/715/723/684/693)
        if K == 1:
            logq = normal_logprob(x, mu_, sig).squeeze(-1)               # plate dim=-2 -> B x 1
            logp = normal_logprob(x, torch.zeros_like(x), torch.ones_like(x)).squeeze(-1)
        else:
            logq = mvn_tril_logprob(x, mu_, L)
            logp = -0.5 * (x ** 2).sum(-1) - K * HALF_LOG_2PI             # MVN(0, I) (vi.py:601-611)
        # model site 'y' (vi.py:595/615) through _get_p_data
        p = irt_prob(model, x, a_, b_, c_, d_, D)
        p, y0, _ = _mask_nan(p, y)
        ll = bernoulli_logprob(p, y0).sum(-1)
        elbo_b = scale * (ll + logp - logq)
        if estimator == "pathwise":
            surrogate = elbo_b.sum()
        elif estimator == "score":
            w = elbo_b.detach() - baseline
            surrogate = (scale * ll).sum() + (w * logq).sum()
        else:
            raise ValueError(estimator)
        total = total + (-surrogate / S)
        loss_value += float(-elbo_b.detach().sum() / S)
    total.backward()
    out = {
        "loss": loss_value,
        "d_mu": _grad(mu_), "d_scale": _grad(sp_),
        "d_a": _grad(a_), "d_b": _grad(b_), "d_c": _grad(c_), "d_d": _grad(d_),
    }
    if a_free is not None and out["d_a"] is not None:
        out["d_a"] = out["d_a"] * a_free.to(dtype)   # vi.py:511-512
    return out


# --------------------------------------------------------------------------
# Enumerated CDMs: TraceEnum_ELBO (Appendix A.3)
# --------------------------------------------------------------------------
def cdm_enum_elbo_grad(
    y: torch.Tensor,                # B x J, NaN = missing
    q: torch.Tensor,                # K x J 0/1
    g: torch.Tensor, s: torch.Tensor,   # 1 x J, in (0,1)
    *,
    model: str = "dina",
    scale: float = 1.0,
    prior_probs: Optional[torch.Tensor] = None,   # B x C (VaeCCDM encoder output); None = uniform
    mask_missing: bool = True,      # VCCDM masks (vi.py:852-856); VaeCCDM does not (vi.py:881-891)
    dtype=torch.float64,
):
    """``TraceEnum_ELBO.loss_and_grads`` for VCCDM.model (vi.py:839-857)."""
    K = q.shape[0]
    C = 2 ** K
    y = y.to(dtype)
    g_, s_ = _leaf(g, dtype), _leaf(s, dtype)
    pp_ = _leaf(prior_probs, dtype)
    patt = all_patterns(K, dtype)
    allp = cdm_prob(model, patt, q.to(dtype), g_, s_)               # C x J
    B = y.shape[0]
    if pp_ is None:
        prior = torch.full((B, C), 1.0 / C, dtype=dtype)
    else:
        prior = pp_
    logprior = categorical_logits(prior)                             # B x C
    p = allp.unsqueeze(1).expand(C, B, y.shape[1])                   # all_p[attr_idx] -> C x B x J
    if mask_missing:
        nan = torch.isnan(y)
        y0 = torch.where(nan, torch.zeros_like(y), y)
        p = torch.where(nan.unsqueeze(0), torch.zeros_like(p), p)
    else:
        y0 = torch.where(torch.isnan(y), torch.full_like(y, -1.0), y)
    ll = bernoulli_logprob(p, y0.unsqueeze(0)).sum(-1)               # C x B
    marg = torch.logsumexp(logprior.T + ll, dim=0)                   # B
    loss = -(scale * marg).sum()
    loss.backward()
    post = torch.softmax((logprior.T + ll).detach(), dim=0).T        # B x C posterior over patterns
    return {"loss": float(loss.detach()), "d_g": _grad(g_), "d_s": _grad(s_),
            "d_prior": _grad(pp_), "posterior": post}


def hodina_elbo_grad(
    y: torch.Tensor, q: torch.Tensor, g: torch.Tensor, s: torch.Tensor,
    lam0: torch.Tensor, lam1: torch.Tensor,       # 1 x K (lam1 positive, constrained value)
    mu: torch.Tensor, sigma: torch.Tensor,        # B x 1
    eps: torch.Tensor,                            # S x B x 1
    *, scale: float = 1.0, dtype=torch.float64,
):
    """``TraceEnum_ELBO.loss_and_grads`` for VCHoDina (vi.py:897-934); always DINA."""
    S, B, _ = eps.shape
    K = q.shape[0]
    y = y.to(dtype)
    eps = eps.to(dtype)
    g_, s_, l0_, l1_ = _leaf(g, dtype), _leaf(s, dtype), _leaf(lam0, dtype), _leaf(lam1, dtype)
    mu_, sg_ = _leaf(mu, dtype), _leaf(sigma, dtype)
    patt = all_patterns(K, dtype)
    nan = torch.isnan(y)
    y0 = torch.where(nan, torch.zeros_like(y), y)
    total = 0.0
    for i in range(S):
        theta = mu_ + sg_ * eps[i]
        allp = cdm_prob("dina", patt, q.to(dtype), g_, s_)          # C x J (vi.py:905)
        attr_p = torch.sigmoid(theta @ l1_ + l0_)                    # B x K (vi.py:911)
        w = torch.exp(torch.log(attr_p) @ patt.T + torch.log(1 - attr_p) @ (1 - patt.T))  # vi.py:912
        logprior = categorical_logits(w)                             # B x C
        p = allp.unsqueeze(1).expand(-1, B, -1)
        p = torch.where(nan.unsqueeze(0), torch.zeros_like(p), p)
        ll = bernoulli_logprob(p, y0.unsqueeze(0)).sum(-1)           # C x B
        marg = torch.logsumexp(logprior.T + ll, dim=0)
        logp = normal_logprob(theta, torch.zeros_like(theta), torch.ones_like(theta)).sum(-1)
        logq = normal_logprob(theta, mu_, sg_).sum(-1)
        total = total - (scale * (marg + logp - logq)).sum() / S
    total.backward()
    return {"loss": float(total.detach()), "d_g": _grad(g_), "d_s": _grad(s_),
            "d_lam0": _grad(l0_), "d_lam1": _grad(l1_), "d_mu": _grad(mu_), "d_sigma": _grad(sg_)}


# --------------------------------------------------------------------------
# REINFORCE CDM: Trace_ELBO with a Bernoulli guide (Appendix A.5)
# --------------------------------------------------------------------------
def cdm_score_elbo_grad(
    y: torch.Tensor, q: torch.Tensor, g: torch.Tensor, s: torch.Tensor,
    attr_p: torch.Tensor,           # B x K guide probabilities (constrained)
    attr: torch.Tensor,             # S x B x K sampled 0/1 attribute vectors
    *, model: str = "dina", scale: float = 1.0, baseline: float = 0.0, dtype=torch.float64,
):
    """``Trace_ELBO.loss_and_grads`` for BaseCDM.model + VCDM.guide
    (vi.py:745-756, 809-816): the only score-function path the reference executes.
    Prior is ``Bernoulli(probs=1.5)`` (sic, vi.py:753); no NaN handling.
    """
    S = attr.shape[0]
    y = y.to(dtype)
    g_, s_, ph_ = _leaf(g, dtype), _leaf(s, dtype), _leaf(attr_p, dtype)
    total = 0.0
    value = 0.0
    for i in range(S):
        at = attr[i].to(dtype)
        logq = bernoulli_logprob(ph_, at).sum(-1)
        logp = bernoulli_logprob(torch.full_like(at, 1.5), at).sum(-1)
        p = cdm_prob(model, at, q.to(dtype), g_, s_)
        ll = bernoulli_logprob(p, y).sum(-1)
        elbo_b = scale * (logp + ll - logq)
        w = elbo_b.detach() - baseline
        surrogate = (scale * ll).sum() + (w * logq).sum()
        total = total - surrogate / S
        value += float(-elbo_b.detach().sum() / S)
    total.backward()
    return {"loss": value, "d_g": _grad(g_), "d_s": _grad(s_), "d_attr_p": _grad(ph_)}


# --------------------------------------------------------------------------
# Counter-based noise: Philox4x32-10 + Box-Muller, restated so that tests can
# check the in-kernel generator of the CUDA path bit-for-bit on the integers.
# --------------------------------------------------------------------------
PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter, key):
    """counter: 4 python ints (32-bit), key: 2 ints -> 4 ints (Salmon et al. 2011)."""
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = (((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
                          ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF)
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3

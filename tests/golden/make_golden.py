"""Generate tests/golden/golden_v1.npz from the reference's own code.

Run ONLY in the builder container (needs /root/reference, which does not exist
on the GPU box):   python tests/golden/make_golden.py

``vi.py`` is imported unmodified under a stub of the absent ``pyro`` package, so
every *pure-torch* piece of the reference is executed for real:
``irt_1pl..irt_4pl`` (vi.py:22-66), ``dina``/``dino`` (vi.py:69-100),
``VCCDM.get_all_attrs`` (vi.py:825-837), ``BaseIRT._get_p_data`` (vi.py:617-625),
the staircase ``a_free``/``a0`` constructor logic (vi.py:566-572) and
``MvnEncoder`` / ``NormEncoder`` (vi.py:417-455).  The estimator level
(Pyro 1.4.0 ``Trace_ELBO`` / ``TraceEnum_ELBO``, not installable here) is
written out with ``torch.distributions`` on top of those reference outputs,
following SURVEY.md Appendix A, in fp32 exactly as the reference would run it.
The resulting vectors pin the oracle (tests/test_oracle_golden.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.distributions as td

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_pyro():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self

    pyro = mod("pyro")
    pyro.distributions = mod("pyro.distributions")
    for n in ("Bernoulli", "Categorical", "Normal", "MultivariateNormal"):
        setattr(pyro.distributions, n, getattr(td, n))
    pyro.distributions.constraints = td.constraints
    sys.modules["pyro.distributions.constraints"] = td.constraints
    pyro.poutine = mod("pyro.poutine")
    msg = mod("pyro.poutine.messenger")
    msg.Messenger = _Any
    pyro.poutine.messenger = msg
    infer = mod("pyro.infer")
    for n in ("SVI", "Trace_ELBO", "TraceEnum_ELBO", "Predictive", "Importance"):
        setattr(infer, n, _Any)
    infer.config_enumerate = lambda f: f
    util = mod("pyro.infer.util")
    util.torch_item = lambda x: x if isinstance(x, float) else x.item()
    infer.util = util
    optim = mod("pyro.optim")
    optim.Adam = _Any
    optim.PyroLRScheduler = _Any
    pyro.infer, pyro.optim = infer, optim
    return pyro


def load_reference():
    _stub_pyro()
    spec = importlib.util.spec_from_file_location("vi_ref", os.path.join(REF, "vi.py"))
    vi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vi)
    return vi


def bern(p):
    return td.Bernoulli(probs=p, validate_args=False)


def main():
    vi = load_reference()
    out = {}
    gen = torch.Generator().manual_seed(20261017)

    def randn(*s):
        return torch.randn(*s, generator=gen)

    def rand(*s):
        return torch.rand(*s, generator=gen)

    # ---- fixtures -------------------------------------------------------
    lsat = np.loadtxt(os.path.join(REF, "lsat.dat"))
    ex52 = np.loadtxt(os.path.join(REF, "ex5.2.dat"))
    out["lsat_y"] = lsat.astype(np.uint8)
    out["ex52_y"] = ex52.astype(np.uint8)

    # ---- L2 link functions ----------------------------------------------
    B, J, K = 37, 11, 3
    x1, xk = randn(B, 1), randn(B, K)
    a1, ak = rand(1, J) * 2.5 + 0.5, rand(K, J) * 2.5 + 0.5
    b = randn(1, J)
    c, d = rand(1, J) * 0.15 + 0.05, rand(1, J) * 0.15 + 0.8
    out.update(link_x1=x1, link_xk=xk, link_a1=a1, link_ak=ak, link_b=b, link_c=c, link_d=d)
    out["link_1pl"] = vi.irt_1pl(x1, b, 1.702)
    out["link_2pl_k1"] = vi.irt_2pl(x1, a1, b, 1.0)
    out["link_2pl_k3"] = vi.irt_2pl(xk, ak, b, 1.702)
    out["link_3pl_k3"] = vi.irt_3pl(xk, ak, b, c, 1.0)
    out["link_4pl_k1"] = vi.irt_4pl(x1, a1, b, c, d, 1.0)
    out["link_4pl_k3"] = vi.irt_4pl(xk, ak, b, c, d, 1.702)

    # ---- CDM tables -----------------------------------------------------
    for Kc in (3, 5):
        data = torch.zeros(4, 9)
        qm = (rand(Kc, 9) < 0.5).float()
        qm[:, 0] = 0                      # item needing no attribute
        qm[:, 1] = 0; qm[0, 1] = 1        # item needing exactly one attribute
        m = vi.VCCDM(q=qm, data=data)
        patt = m.get_all_attrs()
        gg, ss = rand(1, 9) * 0.3, rand(1, 9) * 0.3
        out[f"cdm{Kc}_q"], out[f"cdm{Kc}_patterns"] = qm, patt
        out[f"cdm{Kc}_g"], out[f"cdm{Kc}_s"] = gg, ss
        out[f"cdm{Kc}_dina_p"] = vi.dina(patt, qm, gg, ss)
        out[f"cdm{Kc}_dino_p"] = vi.dino(patt, qm, gg, ss)

    # ---- constructor staircase + NaN rule --------------------------------
    yk = (rand(20, 8) < 0.6).float()
    yk[rand(20, 8) < 0.25] = float("nan")
    m = vi.VIRT(data=yk, model="irt_4pl", x_feature=3)
    out["stair_a_free"] = m.a_free.to(torch.uint8)
    out["stair_a0"] = m.a0
    xs = randn(20, 3)
    a3 = m.a0 * (rand(3, 8) + 0.5)
    kw = dict(x=xs, a=a3, b=randn(1, 8), c=rand(1, 8) * 0.2, d=1 - rand(1, 8) * 0.2, D=1.0)
    p_masked, y_masked = m._get_p_data(yk, torch.arange(20), kw)
    out.update(nan_y=yk, nan_x=xs, nan_a=a3, nan_b=kw["b"], nan_c=kw["c"], nan_d=kw["d"],
               nan_p=p_masked, nan_y0=y_masked)

    # ---- encoder output layout -------------------------------------------
    torch.manual_seed(7)
    enc = vi.MvnEncoder(8, 3, 5)
    yin = torch.where(torch.isnan(yk), torch.full_like(yk, -1.0), yk)
    with torch.no_grad():
        loc, tril = enc(yin)
        packed = enc.fc22(enc.softplus(enc.fc1(yin)))
    out.update(enc_loc=loc, enc_tril=tril, enc_packed=packed)

    # ---- ELBO-level goldens (reference L2 + torch.distributions, fp32) ----
    def irt_case(name, model, K, S, missing, share=False, scale=1.0):
        Bn, Jn = 24, 7
        y = (rand(Bn, Jn) < 0.55).float()
        if missing:
            y[rand(Bn, Jn) < missing] = float("nan")
        obj = vi.VIRT(data=y, model=model, x_feature=K, share_cov=share)
        mu = (randn(Bn, K) * 0.5).requires_grad_(True)
        if K == 1:
            sp = (torch.exp(randn(Bn, 1) * 0.2)).requires_grad_(True)
        else:
            raw = randn(K, K) * 0.2 if share else randn(Bn, K, K) * 0.2
            sp = (raw.tril(-1) + torch.diag_embed(torch.diagonal(raw, dim1=-2, dim2=-1).exp())
                  ).detach().requires_grad_(True)
        a = ((obj.a0 if K > 1 else torch.ones(1, Jn)) * (rand(K, Jn) * 2 + 0.5)).requires_grad_(True)
        bb = randn(1, Jn).requires_grad_(True)
        cc = (rand(1, Jn) * 0.2 + 0.02).requires_grad_(True)
        dd = (1 - rand(1, Jn) * 0.2).requires_grad_(True)
        eps = randn(S, Bn, K)
        kw = {"b": bb, "D": 1.702}
        if model != "irt_1pl":
            kw["a"] = a
        if model in ("irt_3pl", "irt_4pl"):
            kw["c"] = cc
        if model == "irt_4pl":
            kw["d"] = dd
        loss = 0.0
        for s in range(S):
            if K == 1:
                q = td.Normal(mu, sp)
                x = mu + sp * eps[s]
                logq = q.log_prob(x).squeeze(-1)
                logp = td.Normal(torch.zeros(Bn, 1), torch.ones(Bn, 1)).log_prob(x).squeeze(-1)
            else:
                L = sp.tril()
                q = td.MultivariateNormal(mu, scale_tril=L)
                x = mu + (eps[s] @ L.T if share else (L @ eps[s].unsqueeze(-1)).squeeze(-1))
                logq = q.log_prob(x)
                logp = td.MultivariateNormal(torch.zeros(Bn, K), scale_tril=torch.eye(K)).log_prob(x)
            p, y0 = obj._get_p_data(y, torch.arange(Bn), dict(kw, x=x))
            ll = bern(p).log_prob(y0).sum(-1)
            loss = loss - (scale * (ll + logp - logq)).sum() / S
        loss.backward()
        out[f"{name}_y"], out[f"{name}_mu"], out[f"{name}_sp"] = y, mu.detach(), sp.detach()
        out[f"{name}_eps"] = eps
        out[f"{name}_a"], out[f"{name}_b"] = a.detach(), bb.detach()
        out[f"{name}_c"], out[f"{name}_d"] = cc.detach(), dd.detach()
        out[f"{name}_loss"] = loss.detach()
        out[f"{name}_d_mu"], out[f"{name}_d_sp"] = mu.grad, sp.grad
        out[f"{name}_d_b"] = bb.grad
        if model != "irt_1pl":
            ga = a.grad
            if obj.a_free is not None:
                ga = ga * obj.a_free
            out[f"{name}_d_a"] = ga
        if model in ("irt_3pl", "irt_4pl"):
            out[f"{name}_d_c"] = cc.grad
        if model == "irt_4pl":
            out[f"{name}_d_d"] = dd.grad
        out[f"{name}_meta"] = np.array([K, S, 1.702, scale])

    irt_case("irt1", "irt_1pl", 1, 2, 0.0)
    irt_case("irt2", "irt_2pl", 1, 3, 0.3, scale=4.0)
    irt_case("irt3", "irt_3pl", 1, 1, 0.0)
    irt_case("irt4", "irt_4pl", 1, 4, 0.2)
    irt_case("mirt2p", "irt_2pl", 3, 2, 0.2)
    irt_case("mirt4s", "irt_4pl", 3, 2, 0.0, share=True, scale=2.5)

    def cdm_case(name, model, hod):
        Bn, Jn, Kc = 17, 9, 4
        qm = (rand(Kc, Jn) < 0.5).float()
        qm[:, 0] = 0
        qm[:, 1] = 0; qm[1, 1] = 1
        y = (rand(Bn, Jn) < 0.5).float()
        y[rand(Bn, Jn) < 0.2] = float("nan")
        cls = vi.VCHoDina if hod else vi.VCCDM
        obj = cls(q=qm, data=y, model=model)
        patt = obj.all_attr
        gg = (rand(1, Jn) * 0.3 + 0.02).requires_grad_(True)
        ss = (rand(1, Jn) * 0.3 + 0.02).requires_grad_(True)
        nan = torch.isnan(y)
        y0 = torch.where(nan, torch.zeros_like(y), y)
        S = 2 if hod else 1
        scale = 3.0
        out[f"{name}_y"], out[f"{name}_q"] = y, qm
        out[f"{name}_g"], out[f"{name}_s"] = gg.detach(), ss.detach()
        loss = 0.0
        if hod:
            lam0 = randn(1, Kc).requires_grad_(True)
            lam1 = (rand(1, Kc) * 2 + 0.5).requires_grad_(True)
            mu = (randn(Bn, 1) * 0.5).requires_grad_(True)
            sg = torch.exp(randn(Bn, 1) * 0.2).requires_grad_(True)
            eps = randn(S, Bn, 1)
        for i in range(S):
            all_p = (vi.dina if hod else vi.BaseCDM.CDM_FUN[model])(patt, qm, gg, ss)
            if hod:
                theta = mu + sg * eps[i]
                attr_p = torch.sigmoid(theta.mm(lam1) + lam0)
                w = torch.exp(torch.log(attr_p).mm(patt.T) + torch.log(1 - attr_p).mm(1 - patt.T))
                logprior = td.Categorical(probs=w).logits
            else:
                logprior = td.Categorical(probs=torch.zeros(Bn, patt.size(0)) + 1 / patt.size(0)).logits
            p = all_p.unsqueeze(1).expand(-1, Bn, -1)
            p = torch.where(nan.unsqueeze(0), torch.zeros_like(p), p)
            ll = bern(p).log_prob(y0.unsqueeze(0)).sum(-1)
            marg = torch.logsumexp(logprior.T + ll, 0)
            term = marg
            if hod:
                term = term + td.Normal(0.0, 1.0).log_prob(theta).sum(-1) - td.Normal(mu, sg).log_prob(theta).sum(-1)
            loss = loss - (scale * term).sum() / S
        loss.backward()
        out[f"{name}_loss"] = loss.detach()
        out[f"{name}_d_g"], out[f"{name}_d_s"] = gg.grad, ss.grad
        if hod:
            out.update({f"{name}_lam0": lam0.detach(), f"{name}_lam1": lam1.detach(),
                        f"{name}_mu": mu.detach(), f"{name}_sigma": sg.detach(), f"{name}_eps": eps,
                        f"{name}_d_lam0": lam0.grad, f"{name}_d_lam1": lam1.grad,
                        f"{name}_d_mu": mu.grad, f"{name}_d_sigma": sg.grad})

    cdm_case("dina", "dina", False)
    cdm_case("dino", "dino", False)
    cdm_case("hodina", "dina", True)

    # REINFORCE CDM (vi.py:745-756, 809-816)
    Bn, Jn, Kc = 19, 8, 3
    qm = (rand(Kc, Jn) < 0.5).float()
    y = (rand(Bn, Jn) < 0.5).float()
    gg = (rand(1, Jn) * 0.3 + 0.02).requires_grad_(True)
    ss = (rand(1, Jn) * 0.3 + 0.02).requires_grad_(True)
    phi = (rand(Bn, Kc) * 0.8 + 0.1).requires_grad_(True)
    attr = (rand(2, Bn, Kc) < phi.detach()).float()
    loss = 0.0
    for i in range(2):
        logq = bern(phi).log_prob(attr[i]).sum(-1)
        logp = bern(torch.ones(Bn, Kc) + 0.5).log_prob(attr[i]).sum(-1)
        ll = bern(vi.dina(attr[i], qm, gg, ss)).log_prob(y).sum(-1)
        elbo_b = 2.0 * (logp + ll - logq)
        loss = loss - ((2.0 * ll).sum() + (elbo_b.detach() * logq).sum()) / 2
    loss.backward()
    out.update(rf_y=y, rf_q=qm, rf_g=gg.detach(), rf_s=ss.detach(), rf_phi=phi.detach(), rf_attr=attr,
               rf_d_g=gg.grad, rf_d_s=ss.grad, rf_d_phi=phi.grad)

    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **arrays)
    print(f"wrote {len(arrays)} arrays, {os.path.getsize(os.path.join(HERE, 'golden_v1.npz'))} bytes")


if __name__ == "__main__":
    main()

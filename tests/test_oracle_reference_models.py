"""CPU: the oracle against ELBO-level golden vectors produced by the REFERENCE'S OWN model / guide methods
(vi.py, unmodified) executed under tests/golden/minipyro.py -- see tests/golden/make_golden_models.py.

Same data, parameter values, plate subsample and base noise in; loss and the gradient of every unconstrained
parameter out.  The reference ran in fp32 (its native dtype), the oracle arbitrates in fp64: agreement is
required to 1e-4 of each tensor's scale (the REINFORCE weights are O(1e3), their fp32 products carry ~1e-5).
"""
import os

import numpy as np
import pytest
import torch

from oracle import vipsy_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-4


@pytest.fixture(scope="module")
def gm():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_models_v1.npz")))


def T(x):
    return torch.from_numpy(np.asarray(x))


def close(got, ref, what, tol=TOL):
    got, ref = torch.as_tensor(got).double().reshape(-1), torch.as_tensor(ref).double().reshape(-1)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = max(float(ref.abs().max()), 1e-6)
    err = float((got - ref).abs().max())
    assert err <= tol * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


class Case:
    def __init__(self, gm, key):
        self.gm, self.key = gm, key
        self.y = T(gm[f"{key}/data"])
        self.S = int(gm[f"{key}/S"])
        self.loss = float(gm[f"{key}/loss"])
        self.N = self.y.shape[0]
        self.idx = T(gm[f"{key}/idx/0"]).long() if f"{key}/idx/0" in gm else torch.arange(self.N)
        self.B = len(self.idx)
        self.scale = self.N / self.B

    def u(self, name):
        return T(self.gm[f"{self.key}/param_u/{name}"])

    def g(self, name):
        return T(self.gm[f"{self.key}/grad_u/{name}"])

    def noise(self, site):
        return torch.stack([T(self.gm[f"{self.key}/noise/{site}/{s}"]) for s in range(self.S)])

    def has(self, name):
        return f"{self.key}/param_u/{name}" in self.gm

    def encoder(self, cls, *args):
        enc = cls(*args)
        sd = {k.split("$$$")[1]: T(v) for k, v in self.gm.items()
              if k.startswith(f"{self.key}/param_u/encoder$$$")}
        enc.load_state_dict(sd)
        return enc

    def check_encoder_grads(self, enc):
        for n, p in enc.named_parameters():
            close(p.grad, self.g(f"encoder$$${n}"), f"{self.key} encoder {n}", 5e-4)


def irt_items(c, model):
    a = c.u("a") if model != "irt_1pl" else None
    b = c.u("b")
    cc = torch.sigmoid(c.u("c")) if model in ("irt_3pl", "irt_4pl") else None
    dd = torch.sigmoid(c.u("d")) if model == "irt_4pl" else None
    return a, b, cc, dd


def check_irt_items(c, model, ref, cc, dd):
    close(ref["d_b"], c.g("b"), f"{c.key} b")
    if model != "irt_1pl":
        close(ref["d_a"], c.g("a"), f"{c.key} a")
    if cc is not None:
        close(ref["d_c"] * cc * (1 - cc), c.g("c"), f"{c.key} c")
    if dd is not None:
        close(ref["d_d"] * dd * (1 - dd), c.g("d"), f"{c.key} d")


def scatter_rows(c, g):
    full = torch.zeros((c.N,) + tuple(g.shape[1:]), dtype=torch.float64)
    full[c.idx] = g.double()
    return full


@pytest.mark.parametrize("key,model,D", [("virt_irt_1pl_k1_S3", "irt_1pl", 1.0), ("virt_irt_2pl_k1_S3", "irt_2pl", 1.0),
                                         ("virt_irt_3pl_k1_S3", "irt_3pl", 1.702), ("virt_irt_4pl_k1_S3", "irt_4pl", 1.0),
                                         ("virt_irt_2pl_k1_sub15", "irt_2pl", 1.0)])
def test_virt_k1(gm, key, model, D):
    c = Case(gm, key)
    a, b, cc, dd = irt_items(c, model)
    sigma_all = c.u("x_scale").exp()
    ref = o.irt_elbo_grad(c.y[c.idx], c.u("x_local")[c.idx], sigma_all[c.idx], c.noise("x"), a, b, cc, dd, model=model,
                          D=D, scale=c.scale)
    close(ref["loss"], c.loss, f"{key} loss")
    check_irt_items(c, model, ref, cc, dd)
    close(scatter_rows(c, ref["d_mu"]), c.g("x_local"), f"{key} x_local")
    close(scatter_rows(c, ref["d_scale"] * sigma_all[c.idx]), c.g("x_scale"), f"{key} x_scale")


@pytest.mark.parametrize("key,model,shared", [("virt_irt_2pl_k3_person_S2", "irt_2pl", False),
                                              ("virt_irt_4pl_k3_shared_S2", "irt_4pl", True)])
def test_virt_k3(gm, key, model, shared):
    c = Case(gm, key)
    a, b, cc, dd = irt_items(c, model)
    free = T(gm[f"{key}/a_free"]).bool()
    ref = o.irt_elbo_grad(c.y, c.u("x_local"), c.u("x_scale"), c.noise("x"), a, b, cc, dd, model=model, a_free=free,
                          chol_raw=True)
    close(ref["loss"], c.loss, f"{key} loss")
    check_irt_items(c, model, ref, cc, dd)
    assert float((c.g("a") * (~free)).abs().max()) == 0.0
    close(ref["d_mu"], c.g("x_local"), f"{key} x_local")
    close(ref["d_scale"].tril(), c.g("x_scale"), f"{key} x_scale (raw lower_cholesky)")


def test_vaeirt_k1(gm):
    from vipsy_b200.models import NormEncoder
    key = "vaeirt_irt_3pl_k1_S2"
    c = Case(gm, key)
    enc = c.encoder(NormEncoder, c.y.shape[1], 1, 5)
    yin = torch.where(torch.isnan(c.y), torch.full_like(c.y, -1.0), c.y)
    loc, sc = enc(yin)
    a, b, cc, dd = irt_items(c, "irt_3pl")
    ref = o.irt_elbo_grad(c.y, loc.detach(), sc.detach(), c.noise("x"), a, b, cc, dd, model="irt_3pl")
    close(ref["loss"], c.loss, f"{key} loss")
    check_irt_items(c, "irt_3pl", ref, cc, dd)
    torch.autograd.backward([loc, sc], [ref["d_mu"].float(), ref["d_scale"].float()])
    c.check_encoder_grads(enc)


def test_vaeirt_k3_subsampled(gm):
    from vipsy_b200.models import MvnEncoder
    key = "vaeirt_irt_2pl_k3_sub12"
    c = Case(gm, key)
    enc = c.encoder(MvnEncoder, c.y.shape[1], 3, 5)
    yb = c.y[c.idx]
    loc, packed = enc(torch.where(torch.isnan(yb), torch.full_like(yb, -1.0), yb))
    raw = o.tril_from_packed(packed.detach(), 3)                  # the reference's dense scatter (vi.py:452-454)
    a, b, _, _ = irt_items(c, "irt_2pl")
    free = T(gm[f"{key}/a_free"]).bool()
    ref = o.irt_elbo_grad(yb, loc.detach(), raw, c.noise("x"), a, b, model="irt_2pl", scale=c.scale, a_free=free,
                          chol_raw=True)
    close(ref["loss"], c.loss, f"{key} loss")
    check_irt_items(c, "irt_2pl", ref, None, None)
    ti = torch.tril_indices(3, 3)
    torch.autograd.backward([loc, packed], [ref["d_mu"].float(), ref["d_scale"][:, ti[0], ti[1]].float()])
    c.check_encoder_grads(enc)


@pytest.mark.parametrize("key,model", [("vcdm_dina_S2", "dina"), ("vcdm_dino_S2", "dino")])
def test_vcdm_reinforce(gm, key, model):
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s, phi = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s")), torch.sigmoid(c.u("attr_p"))
    attr = (c.noise("attr") < phi).float()
    ref = o.cdm_score_elbo_grad(c.y, q, g, s, phi, attr, model=model)
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    close(ref["d_s"] * s * (1 - s), c.g("s"), f"{key} s")
    close(ref["d_attr_p"] * phi * (1 - phi), c.g("attr_p"), f"{key} attr_p")


def test_vaecdm_subsampled(gm):
    from vipsy_b200.models import BinEncoder
    key = "vaecdm_dina_sub10"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    enc = c.encoder(BinEncoder, c.y.shape[1], q.shape[0], 5)
    yb = c.y[c.idx]
    phi = enc(yb)
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    attr = (c.noise("attr") < phi.detach()).float()
    ref = o.cdm_score_elbo_grad(yb, q, g, s, phi.detach(), attr, model="dina", scale=c.scale)
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    phi.backward(ref["d_attr_p"].float())
    c.check_encoder_grads(enc)


@pytest.mark.parametrize("key,model", [("vccdm_dina", "dina"), ("vccdm_dino", "dino"), ("vccdm_dina_sub16", "dina")])
def test_vccdm_enumerated(gm, key, model):
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    ref = o.cdm_enum_elbo_grad(c.y[c.idx], q, g, s, model=model, scale=c.scale)
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    close(ref["d_s"] * s * (1 - s), c.g("s"), f"{key} s")


def test_vaeccdm(gm):
    from vipsy_b200.models import SoftmaxEncoder
    key = "vaeccdm_dina"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    enc = c.encoder(SoftmaxEncoder, c.y.shape[1], 2 ** q.shape[0], 5)
    attr_p = enc(torch.where(torch.isnan(c.y), torch.full_like(c.y, -1.0), c.y))
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    ref = o.cdm_enum_elbo_grad(c.y, q, g, s, model="dina", prior_probs=attr_p.detach(), mask_missing=False)
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    close(ref["d_s"] * s * (1 - s), c.g("s"), f"{key} s")
    attr_p.backward(ref["d_prior"].float())
    gmax = max(float(c.g(f"encoder$$${n}").abs().max()) for n, _ in enc.named_parameters())
    for n, p in enc.named_parameters():            # fc2.bias has an exactly-zero gradient (softmax over the batch axis)
        assert float((p.grad.double() - c.g(f"encoder$$${n}").double()).abs().max()) <= 5e-4 * gmax, n


def test_vchodina(gm):
    key = "vchodina_S2"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    lam1, sigma = c.u("lam1").exp(), c.u("theta_scale").exp()
    ref = o.hodina_elbo_grad(c.y, q, g, s, c.u("lam0"), lam1, c.u("theta_local"), sigma, c.noise("theta"))
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    close(ref["d_s"] * s * (1 - s), c.g("s"), f"{key} s")
    close(ref["d_lam0"], c.g("lam0"), f"{key} lam0")
    close(ref["d_lam1"] * lam1, c.g("lam1"), f"{key} lam1")
    close(ref["d_mu"], c.g("theta_local"), f"{key} theta_local")
    close(ref["d_sigma"] * sigma, c.g("theta_scale"), f"{key} theta_scale")


def test_vaechodina_subsampled(gm):
    from vipsy_b200.models import NormEncoder
    key = "vaechodina_sub9"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    enc = c.encoder(NormEncoder, c.y.shape[1], 1, 5)
    yb = c.y[c.idx]
    loc, sc = enc(torch.where(torch.isnan(yb), torch.full_like(yb, -1.0), yb))
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    lam1 = c.u("lam1").exp()
    ref = o.hodina_elbo_grad(yb, q, g, s, c.u("lam0"), lam1, loc.detach(), sc.detach(), c.noise("theta"), scale=c.scale)
    close(ref["loss"], c.loss, f"{key} loss")
    close(ref["d_lam1"] * lam1, c.g("lam1"), f"{key} lam1")
    close(ref["d_g"] * g * (1 - g), c.g("g"), f"{key} g")
    torch.autograd.backward([loc, sc], [ref["d_mu"].float(), ref["d_sigma"].float()])
    c.check_encoder_grads(enc)

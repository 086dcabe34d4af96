"""GPU: the CUDA kernels (through the C ABI) against the golden vectors produced by the reference's own
model / guide methods under tests/golden/minipyro.py -- no oracle in between.  Same data, parameters,
subsample and base noise; loss and unconstrained-parameter gradients must match the reference's fp32 run
to 1e-4 of each tensor's scale."""
import pytest
import torch

from tests.test_oracle_reference_models import Case, T, close, gm, irt_items, scatter_rows  # noqa: F401
from tests.util import codes_from_float

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


def cu(t):
    return None if t is None else t.cuda()


@pytest.mark.parametrize("key,model,D", [("virt_irt_1pl_k1_S3", "irt_1pl", 1.0), ("virt_irt_2pl_k1_S3", "irt_2pl", 1.0),
                                         ("virt_irt_3pl_k1_S3", "irt_3pl", 1.702), ("virt_irt_4pl_k1_S3", "irt_4pl", 1.0),
                                         ("virt_irt_2pl_k1_sub15", "irt_2pl", 1.0)])
def test_k1_kernels_vs_reference_model(ops, gm, key, model, D):
    c = Case(gm, key)
    a, b, cc, dd = irt_items(c, model)
    sigma_all = c.u("x_scale").exp()
    r = ops.irt_elbo_grad(codes_from_float(c.y).cuda(), cu(c.u("x_local")[c.idx]), sigma=cu(sigma_all[c.idx]),
                          eps=cu(c.noise("x")), a=cu(a), b=cu(b), c=cu(cc), d=cu(dd), model=model, D=D, scale=c.scale,
                          idx=cu(c.idx))
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_b"].cpu(), c.g("b"), f"{key} b")
    if model != "irt_1pl":
        close(r["d_a"].cpu(), c.g("a"), f"{key} a")
    if cc is not None:
        close(r["d_c"].cpu() * cc * (1 - cc), c.g("c"), f"{key} c")
    if dd is not None:
        close(r["d_d"].cpu() * dd * (1 - dd), c.g("d"), f"{key} d")
    close(scatter_rows(c, r["d_mu"].cpu()), c.g("x_local"), f"{key} x_local")
    close(scatter_rows(c, r["d_sigma"].cpu() * sigma_all[c.idx]), c.g("x_scale"), f"{key} x_scale")


@pytest.mark.parametrize("key,model", [("virt_irt_2pl_k3_person_S2", "irt_2pl"), ("virt_irt_4pl_k3_shared_S2", "irt_4pl")])
def test_mirt_kernels_vs_reference_model(ops, gm, key, model):
    c = Case(gm, key)
    a, b, cc, dd = irt_items(c, model)
    free = T(gm[f"{key}/a_free"]).bool()
    r = ops.irt_elbo_grad(codes_from_float(c.y).cuda(), cu(c.u("x_local")), chol=cu(c.u("x_scale")), eps=cu(c.noise("x")),
                          a=cu(a), b=cu(b), c=cu(cc), d=cu(dd), model=model, a_free=cu(free), chol_raw=True)
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_a"].cpu(), c.g("a"), f"{key} a")
    close(r["d_b"].cpu(), c.g("b"), f"{key} b")
    close(r["d_mu"].cpu(), c.g("x_local"), f"{key} x_local")
    close(r["d_chol"].cpu().tril(), c.g("x_scale"), f"{key} x_scale (raw lower_cholesky)")


@pytest.mark.parametrize("key,model", [("vcdm_dina_S2", "dina"), ("vcdm_dino_S2", "dino")])
def test_reinforce_kernel_vs_reference_model(ops, gm, key, model):
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s, phi = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s")), torch.sigmoid(c.u("attr_p"))
    r = ops.cdm_score_elbo_grad(codes_from_float(c.y).cuda(), q.cuda(), g.cuda(), s.cuda(), phi.cuda(), u=cu(c.noise("attr")),
                                model=model)
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_g"].cpu() * g * (1 - g), c.g("g"), f"{key} g")
    close(r["d_s"].cpu() * s * (1 - s), c.g("s"), f"{key} s")
    close(r["d_attr_p"].cpu() * phi * (1 - phi), c.g("attr_p"), f"{key} attr_p")


@pytest.mark.parametrize("key,model", [("vccdm_dina", "dina"), ("vccdm_dino", "dino"), ("vccdm_dina_sub16", "dina")])
def test_enumeration_kernel_vs_reference_model(ops, gm, key, model):
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    r = ops.cdm_enum_elbo_grad(codes_from_float(c.y).cuda(), q.cuda(), g.cuda(), s.cuda(), model=model, scale=c.scale,
                               idx=cu(c.idx))
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_g"].cpu() * g * (1 - g), c.g("g"), f"{key} g")
    close(r["d_s"].cpu() * s * (1 - s), c.g("s"), f"{key} s")


def test_hodina_kernel_vs_reference_model(ops, gm):
    key = "vchodina_S2"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    lam1, sigma = c.u("lam1").exp(), c.u("theta_scale").exp()
    r = ops.cdm_enum_elbo_grad(codes_from_float(c.y).cuda(), q.cuda(), g.cuda(), s.cuda(), lam0=cu(c.u("lam0")),
                               lam1=cu(lam1), mu=cu(c.u("theta_local")), sigma=cu(sigma), eps=cu(c.noise("theta")))
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_g"].cpu() * g * (1 - g), c.g("g"), f"{key} g")
    close(r["d_lam0"].cpu(), c.g("lam0"), f"{key} lam0")
    close(r["d_lam1"].cpu() * lam1, c.g("lam1"), f"{key} lam1")
    close(r["d_mu"].cpu(), c.g("theta_local"), f"{key} theta_local")
    close(r["d_sigma"].cpu() * sigma, c.g("theta_scale"), f"{key} theta_scale")


def test_vaeccdm_kernel_vs_reference_model(ops, gm):
    from vipsy_b200.models import SoftmaxEncoder
    key = "vaeccdm_dina"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    enc = c.encoder(SoftmaxEncoder, c.y.shape[1], 2 ** q.shape[0], 5)
    attr_p = enc(torch.where(torch.isnan(c.y), torch.full_like(c.y, -1.0), c.y))
    prior = attr_p / attr_p.sum(-1, keepdim=True)
    eps = torch.finfo(torch.float32).eps
    log_prior = torch.log(prior.clamp(min=eps, max=1 - eps))
    g, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    r = ops.cdm_enum_elbo_grad(codes_from_float(c.y).cuda(), q.cuda(), g.cuda(), s.cuda(), model="dina",
                               log_prior=log_prior.detach().cuda(), score_missing=True, want_posterior=True)
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_g"].cpu() * g * (1 - g), c.g("g"), f"{key} g")
    close(r["d_s"].cpu() * s * (1 - s), c.g("s"), f"{key} s")
    log_prior.backward(-r["posterior"].cpu())
    gmax = max(float(c.g(f"encoder$$${n}").abs().max()) for n, _ in enc.named_parameters())
    for n, p in enc.named_parameters():
        assert float((p.grad.double() - c.g(f"encoder$$${n}").double()).abs().max()) <= 5e-4 * gmax, n


@pytest.mark.parametrize("tc", [True, False])
def test_vaeirt_k1_encoder_kernels_vs_reference_model(ops, gm, tc):
    """VaeIRT, K = 1 (vi.py:673-684): NormEncoder kernels -> ELBO kernel -> encoder backward kernels, against the
    loss and gradients the reference's own guide + model produced (encoder weights included)."""
    from vipsy_b200.models import NormEncoder
    key = "vaeirt_irt_3pl_k1_S2"
    c = Case(gm, key)
    enc = c.encoder(NormEncoder, c.y.shape[1], 1, 5).cuda()
    codes = codes_from_float(c.y).cuda()
    mu, sigma, pre = ops.norm_encoder_forward(codes, enc, tensor_cores=tc)
    a, b, cc, dd = irt_items(c, "irt_3pl")
    r = ops.irt_elbo_grad(codes, mu.reshape(-1, 1), sigma=sigma.reshape(-1, 1), eps=cu(c.noise("x")), a=cu(a), b=cu(b),
                          c=cu(cc), d=cu(dd), model="irt_3pl")
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_b"].cpu(), c.g("b"), f"{key} b")
    close(r["d_a"].cpu(), c.g("a"), f"{key} a")
    close(r["d_c"].cpu() * cc * (1 - cc), c.g("c"), f"{key} c")
    g = ops.norm_encoder_backward(codes, enc, sigma, pre, r["d_mu"].reshape(-1), r["d_sigma"].reshape(-1), tensor_cores=tc)
    for n, _ in enc.named_parameters():
        close(g[n].cpu(), c.g(f"encoder$$${n}"), f"{key} encoder {n}", 5e-4)


def test_vaechodina_encoder_kernels_vs_reference_model(ops, gm):
    """VaeCHoDina (vi.py:961-981), subsampled: theta encoder kernels (minibatch through idx) around the enumeration
    kernel, against the reference's own loss and gradients."""
    from vipsy_b200.models import NormEncoder
    key = "vaechodina_sub9"
    c = Case(gm, key)
    q = T(gm["cdm/q"])
    enc = c.encoder(NormEncoder, c.y.shape[1], 1, 5).cuda()
    codes = codes_from_float(c.y).cuda()
    idx = cu(c.idx)
    mu, sigma, pre = ops.norm_encoder_forward(codes, enc, idx)
    gg, s = torch.sigmoid(c.u("g")), torch.sigmoid(c.u("s"))
    lam1 = c.u("lam1").exp()
    r = ops.cdm_enum_elbo_grad(codes, q.cuda(), gg.cuda(), s.cuda(), lam0=cu(c.u("lam0")), lam1=cu(lam1),
                               mu=mu.reshape(-1, 1), sigma=sigma.reshape(-1, 1), eps=cu(c.noise("theta")), idx=idx,
                               scale=c.scale)
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_lam1"].cpu() * lam1, c.g("lam1"), f"{key} lam1")
    close(r["d_g"].cpu() * gg * (1 - gg), c.g("g"), f"{key} g")
    g = ops.norm_encoder_backward(codes, enc, sigma, pre, r["d_mu"].reshape(-1), r["d_sigma"].reshape(-1), idx)
    for n, _ in enc.named_parameters():
        close(g[n].cpu(), c.g(f"encoder$$${n}"), f"{key} encoder {n}", 5e-4)


def test_vaeirt_k3_fused_mvn_encoder_vs_reference_model(ops, gm):
    """VaeIRT, K = 3, subsampled (vi.py:685-693): fc1 / fc21 in torch, fc22 + tril scatter + LowerCholeskyTransform +
    rsample fused into the CUDA path, against the loss and every gradient (fc22's weights included) that the
    reference's own guide + model produced."""
    from vipsy_b200.models import MvnEncoder
    key = "vaeirt_irt_2pl_k3_sub12"
    c = Case(gm, key)
    enc = c.encoder(MvnEncoder, c.y.shape[1], 3, 5).cuda()
    yb = c.y[c.idx].cuda()
    hidden = enc.softplus(enc.fc1(torch.where(torch.isnan(yb), torch.full_like(yb, -1.0), yb)))
    loc = enc.fc21(hidden)
    a, b, _, _ = irt_items(c, "irt_2pl")
    free = T(gm[f"{key}/a_free"]).bool()
    r = ops.irt_elbo_grad(codes_from_float(c.y).cuda(), loc.detach(),
                          mvn_encoder=(hidden.detach(), enc.fc22.weight.detach(), enc.fc22.bias.detach()),
                          eps=cu(c.noise("x")), a=cu(a), b=cu(b), model="irt_2pl", scale=c.scale, idx=cu(c.idx),
                          a_free=cu(free))
    close(r["loss"].cpu(), c.loss, f"{key} loss")
    close(r["d_a"].cpu(), c.g("a"), f"{key} a")
    close(r["d_b"].cpu(), c.g("b"), f"{key} b")
    torch.autograd.backward([hidden, loc], [r["d_hidden"], r["d_mu"]])
    enc.fc22.weight.grad, enc.fc22.bias.grad = r["d_w22"], r["d_b22"]
    for n, p in enc.named_parameters():
        close(p.grad.cpu(), c.g(f"encoder$$${n}"), f"{key} encoder {n}", 5e-4)

"""GPU parity: enumerated DINA / DINO / HO-DINA kernel (lattice transform) vs the oracle."""
import pytest
import torch

from oracle import vipsy_oracle as o
from tests.util import assert_close, codes_from_float

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


def make_cdm(B, J, K, seed, missing=0.1, q_p=0.5):
    g = torch.Generator().manual_seed(seed)
    q = (torch.rand(K, J, generator=g) < q_p).float()
    if J > 2:
        q[:, 0] = 0                      # item needing nothing
        q[:, 1] = 0; q[K // 2, 1] = 1    # item needing exactly one attribute
        q[:, 2] = q[:, J - 1]            # duplicated mask: exercises the scatter rounds
    y = (torch.rand(B, J, generator=g) < 0.5).float()
    if missing:
        y[torch.rand(B, J, generator=g) < missing] = float("nan")
    gg = torch.rand(1, J, generator=g) * 0.3 + 0.02
    ss = torch.rand(1, J, generator=g) * 0.3 + 0.02
    return q, y, gg, ss, g


@pytest.mark.parametrize("model", ["dina", "dino"])
@pytest.mark.parametrize("B,J,K", [(1, 3, 1), (37, 9, 3), (130, 20, 5), (64, 60, 10), (50, 100, 7), (33, 33, 10)])
def test_vccdm_parity(ops, model, B, J, K):
    q, y, gg, ss, _ = make_cdm(B, J, K, seed=B + J + K)
    ref = o.cdm_enum_elbo_grad(y, q, gg, ss, model=model, scale=2.0)
    got = ops.cdm_enum_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), model=model, scale=2.0,
                                 want_posterior=True)
    torch.cuda.synchronize()
    assert_close(got["loss"], ref["loss"], 1e-5, 1e-5, "loss")
    assert_close(got["d_g"], ref["d_g"], 1e-5, 1e-5, "d_g")
    assert_close(got["d_s"], ref["d_s"], 1e-5, 1e-5, "d_s")
    assert_close(got["posterior"], ref["posterior"], 1e-4, 1e-5, "posterior")


def test_eta_table_bit_exact(ops, golden):
    for K in (3, 5):
        q = torch.from_numpy(golden[f"cdm{K}_q"])
        patt = torch.from_numpy(golden[f"cdm{K}_patterns"])
        qm = ops.qmask_from_q(q.cuda())
        for model in ("dina", "dino"):
            eta = ops.cdm_eta_table(qm, K, model).cpu()
            assert torch.equal(eta.float(), o.cdm_eta(model, patt, q))
    g = torch.Generator().manual_seed(0)
    q = (torch.rand(10, 60, generator=g) < 0.5).float()
    patt = o.all_patterns(10)
    for model in ("dina", "dino"):
        eta = ops.cdm_eta_table(ops.qmask_from_q(q.cuda()), 10, model).cpu()
        assert torch.equal(eta.float(), o.cdm_eta(model, patt, q))


def test_golden_cdm_through_the_abi(ops, golden):
    g = golden
    t = lambda k: torch.from_numpy(g[k])
    for name in ("dina", "dino"):
        got = ops.cdm_enum_elbo_grad(codes_from_float(t(f"{name}_y")).cuda(), t(f"{name}_q").cuda(),
                                     t(f"{name}_g").cuda(), t(f"{name}_s").cuda(), model=name, scale=3.0)
        assert_close(got["loss"], t(f"{name}_loss"), 2e-5, 2e-5, name)
        assert_close(got["d_g"], t(f"{name}_d_g"), 2e-5, 2e-5, name)
        assert_close(got["d_s"], t(f"{name}_d_s"), 2e-5, 2e-5, name)
    n = "hodina"
    got = ops.cdm_enum_elbo_grad(codes_from_float(t(f"{n}_y")).cuda(), t(f"{n}_q").cuda(), t(f"{n}_g").cuda(),
                                 t(f"{n}_s").cuda(), lam0=t(f"{n}_lam0").cuda(), lam1=t(f"{n}_lam1").cuda(),
                                 mu=t(f"{n}_mu").cuda(), sigma=t(f"{n}_sigma").cuda(), eps=t(f"{n}_eps").cuda(),
                                 scale=3.0)
    for k in ("loss", "d_g", "d_s", "d_lam0", "d_lam1", "d_mu", "d_sigma"):
        assert_close(got[k], t(f"{n}_{k}"), 3e-5, 3e-5, k)


@pytest.mark.parametrize("B,J,K,S", [(40, 12, 4, 1), (97, 30, 6, 3), (64, 60, 10, 2), (10, 60, 10, 5)])
def test_hodina_parity(ops, B, J, K, S):
    q, y, gg, ss, g = make_cdm(B, J, K, seed=100 + K)
    lam0 = torch.randn(1, K, generator=g)
    lam1 = torch.rand(1, K, generator=g) * 2.5 + 0.5
    mu = torch.randn(B, 1, generator=g) * 0.7
    sigma = torch.exp(torch.randn(B, 1, generator=g) * 0.3)
    eps = torch.randn(S, B, 1, generator=g)
    ref = o.hodina_elbo_grad(y, q, gg, ss, lam0, lam1, mu, sigma, eps, scale=1.5)
    got = ops.cdm_enum_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), lam0=lam0.cuda(),
                                 lam1=lam1.cuda(), mu=mu.cuda(), sigma=sigma.cuda(), eps=eps.cuda(), scale=1.5)
    torch.cuda.synchronize()
    for k in ("loss", "d_g", "d_s", "d_lam0", "d_lam1", "d_mu", "d_sigma"):
        assert_close(got[k], ref[k], 2e-5, 2e-5, k)


def test_hodina_prior_floor(ops):
    """Steep lam1 pushes many pattern priors below eps: the Categorical floor
    changes the loss by ~1e-3 relative and must be reproduced, gradients included."""
    B, J, K = 80, 40, 10
    q, y, gg, ss, g = make_cdm(B, J, K, seed=7)
    lam0 = torch.randn(1, K, generator=g) * 2
    lam1 = torch.rand(1, K, generator=g) * 3 + 2.0
    mu = torch.randn(B, 1, generator=g)
    sigma = torch.ones(B, 1) * 0.5
    eps = torch.randn(2, B, 1, generator=g)
    ref = o.hodina_elbo_grad(y, q, gg, ss, lam0, lam1, mu, sigma, eps)
    got = ops.cdm_enum_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), lam0=lam0.cuda(),
                                 lam1=lam1.cuda(), mu=mu.cuda(), sigma=sigma.cuda(), eps=eps.cuda())
    for k in ("loss", "d_g", "d_s", "d_lam0", "d_lam1", "d_mu", "d_sigma"):
        assert_close(got[k], ref[k], 3e-5, 3e-5, k)


def test_cdm_deterministic_and_idx(ops):
    q, y, gg, ss, g = make_cdm(3000, 60, 10, seed=3)
    codes = codes_from_float(y).cuda()
    a = ops.cdm_enum_elbo_grad(codes, q.cuda(), gg.cuda(), ss.cuda())
    b = ops.cdm_enum_elbo_grad(codes, q.cuda(), gg.cuda(), ss.cuda())
    assert torch.equal(a["loss"], b["loss"]) and torch.equal(a["item_grads"], b["item_grads"])
    idx = torch.randperm(3000, generator=g)[:500]
    sub = ops.cdm_enum_elbo_grad(codes, q.cuda(), gg.cuda(), ss.cuda(), idx=idx.cuda(), scale=6.0)
    ref = o.cdm_enum_elbo_grad(y[idx], q, gg, ss, scale=6.0)
    assert_close(sub["loss"], ref["loss"], 1e-5, 1e-5)
    assert_close(sub["d_g"], ref["d_g"], 1e-5, 1e-5)


@pytest.mark.parametrize("model", ["dina", "dino"])
@pytest.mark.parametrize("B,J,K", [(37, 9, 3), (64, 60, 10), (50, 100, 7), (20, 33, 1)])
def test_vaeccdm_prior_and_scored_missing_cells(ops, model, B, J, K):
    """VaeCCDM.model (vi.py:872-891): per-person pattern prior through Categorical(probs) (normalise, clamp,
    log) and missing cells scored as y = -1 instead of masked; d loss / d log_prior = -scale * posterior."""
    q, y, gg, ss, g = make_cdm(B, J, K, seed=7 * B + K, missing=0.2)
    C = 2 ** K
    pp = torch.rand(B, C, generator=g) ** 4 + 1e-9
    pp[0, 0] = 1e-12                                   # a floored entry
    ref = o.cdm_enum_elbo_grad(y, q, gg, ss, model=model, scale=2.0, prior_probs=pp, mask_missing=False)
    prior = (pp.double() / pp.double().sum(-1, keepdim=True)).clamp(min=o.EPS32, max=1 - o.EPS32)
    got = ops.cdm_enum_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), model=model, scale=2.0,
                                 log_prior=prior.log().float().cuda(), score_missing=True, want_posterior=True)
    torch.cuda.synchronize()
    assert_close(got["loss"], ref["loss"], 1e-5, 1e-5, "loss")
    assert_close(got["d_g"], ref["d_g"], 1e-5, 1e-5, "d_g")
    assert_close(got["d_s"], ref["d_s"], 1e-5, 1e-5, "d_s")
    assert_close(got["posterior"], ref["posterior"], 1e-4, 1e-5, "posterior")

"""GPU: edge cases of the three entry points -- empty batch, one person / one item, fully missing matrices --
and the masked-cell constant (a NaN cell scores log(1 - eps), vi.py:621-624 under torch's clamp) in every family."""
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


@pytest.mark.parametrize("S", [1, 4])
def test_empty_batch_irt_uni(ops, S):
    y = torch.zeros((0, 64), dtype=torch.uint8, device="cuda")
    z = torch.zeros((0, 1), device="cuda")
    r = ops.irt_elbo_grad(y, z, sigma=z + 1, n_particles=S, a=torch.ones(1, 64, device="cuda"),
                          b=torch.zeros(1, 64, device="cuda"), model="irt_2pl")
    torch.cuda.synchronize()
    assert float(r["loss"]) == 0.0 and r["d_mu"].shape == (0, 1)
    assert float(r["d_a"].abs().max()) == 0.0 and float(r["d_b"].abs().max()) == 0.0


def test_empty_batch_mirt_and_cdm(ops):
    dev = "cuda"
    y = torch.zeros((0, 12), dtype=torch.uint8, device=dev)
    r = ops.irt_elbo_grad(y, torch.zeros((0, 3), device=dev), chol=torch.zeros((0, 3, 3), device=dev), n_particles=2,
                          a=torch.ones(3, 12, device=dev), b=torch.zeros(1, 12, device=dev), model="irt_2pl", chol_raw=True)
    assert float(r["loss"]) == 0.0
    q = (torch.rand(3, 12) < 0.5).float().to(dev)
    g = torch.full((1, 12), 0.1, device=dev)
    r = ops.cdm_enum_elbo_grad(y, q, g, g, batch=0)
    assert float(r["loss"]) == 0.0 and float(r["d_g"].abs().max()) == 0.0
    r = ops.cdm_score_elbo_grad(y, q, g, g, torch.zeros((0, 3), device=dev))
    torch.cuda.synchronize()
    assert float(r["loss"]) == 0.0


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_4pl"])
def test_one_person_one_item(ops, model):
    y = torch.tensor([[1.0]])
    mu, sg, eps = torch.tensor([[0.3]]), torch.tensor([[0.8]]), torch.tensor([[[0.5]], [[-1.0]]])
    a, b, c, d = torch.tensor([[1.2]]), torch.tensor([[-0.4]]), torch.tensor([[0.1]]), torch.tensor([[0.9]])
    kw = dict(a=a if model != "irt_1pl" else None, b=b, c=c if model == "irt_4pl" else None, d=d if model == "irt_4pl" else None)
    ref = o.irt_elbo_grad(y, mu, sg, eps, kw["a"], kw["b"], kw["c"], kw["d"], model=model)
    got = ops.irt_elbo_grad(codes_from_float(y).cuda(), mu.cuda(), sigma=sg.cuda(), eps=eps.cuda(),
                            **{k: (v.cuda() if v is not None else None) for k, v in kw.items()}, model=model)
    assert_close(got["loss"], ref["loss"], 1e-5, 1e-5, "loss")
    assert_close(got["d_mu"], ref["d_mu"], 1e-5, 1e-5, "d_mu")
    assert_close(got["d_b"], ref["d_b"], 1e-5, 1e-5, "d_b")


def test_masked_cell_constant_mirt(ops):
    """K > 1 with 97 % of the cells missing: the loss is dominated by prior terms, so the log(1-eps) per masked
    cell is visible at 1e-6 relative -- and must be there."""
    from tests.test_parity_mirt import make, packed_of
    cs = make(B=300, J=400, K=4, S=1, model="irt_2pl", seed=3, missing=0.97)
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], cs["raw"], cs["eps"], cs["a"], cs["b"], None, None, model="irt_2pl",
                          a_free=cs["free"], chol_raw=True)
    dev = "cuda"
    got = ops.irt_elbo_grad(codes_from_float(cs["y"]).to(dev), cs["mu"].to(dev), chol=packed_of(cs["raw"]).to(dev),
                            eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), model="irt_2pl",
                            a_free=cs["free"].to(dev), chol_packed=True, chol_raw=True)
    n_masked = int(torch.isnan(cs["y"]).sum())
    assert n_masked * 1.19e-7 > 3e-6 * abs(ref["loss"])          # the term is resolvable
    assert_close(got["loss"], ref["loss"], 1e-6, 1e-6, "loss")


@pytest.mark.parametrize("hod", [False, True])
def test_masked_cell_constant_cdm_enum(ops, hod):
    from tests.test_parity_cdm_enum import make_cdm
    q, y, gg, ss, g = make_cdm(400, 200, 4, seed=9, missing=0.985)
    kw, okw = {}, {}
    if hod:
        mu, sg = torch.randn(400, 1, generator=g) * 0.3, torch.exp(torch.randn(400, 1, generator=g) * 0.1)
        eps = torch.randn(1, 400, 1, generator=g)
        lam0, lam1 = torch.randn(1, 4, generator=g) * 0.5, torch.rand(1, 4, generator=g) + 0.5
        okw = dict(lam0=lam0, lam1=lam1, mu=mu, sigma=sg, eps=eps)
        kw = {k: v.cuda() for k, v in okw.items()}
    if hod:
        ref = o.hodina_elbo_grad(y, q, gg, ss, okw["lam0"], okw["lam1"], okw["mu"], okw["sigma"], okw["eps"])
    else:
        ref = o.cdm_enum_elbo_grad(y, q, gg, ss, model="dina")
    got = ops.cdm_enum_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), model="dina", **kw)
    n_masked = int(torch.isnan(y).sum())
    assert n_masked * 1.19e-7 > 3e-6 * abs(ref["loss"])
    assert_close(got["loss"], ref["loss"], 1e-6, 1e-6, "loss")


def test_fully_missing_matrix_every_k1_kernel(ops):
    """All cells NaN: loss = prior/entropy terms + N J log(1-eps); every K = 1 mapping (particle kernel, per-lane
    compaction, row, row-sparse, CSR) returns the same number."""
    g = torch.Generator().manual_seed(0)
    N, J = 500, 128
    y = torch.full((N, J), float("nan"))
    mu, sg = torch.randn(N, 1, generator=g) * 0.5, torch.exp(torch.randn(N, 1, generator=g) * 0.2)
    a, b = torch.rand(1, J, generator=g) + 0.5, torch.randn(1, J, generator=g)
    dev = "cuda"
    codes = codes_from_float(y).to(dev)
    for S in (1, 4):
        eps = torch.randn(S, N, 1, generator=g)
        ref = o.irt_elbo_grad(y, mu, sg, eps, a, b, None, None, model="irt_2pl", scale=2.0)
        kw = dict(sigma=sg.to(dev), eps=eps.to(dev), a=a.to(dev), b=b.to(dev), model="irt_2pl", scale=2.0)
        variants = [dict(row_kernel=False), dict(row_kernel=False, sparse=True), dict(), dict(sparse=True),
                    dict(obs=ops.observed_csr(codes)), dict(row_kernel=False, n_masked=N * J)]
        for v in variants:
            got = ops.irt_elbo_grad(codes, mu.to(dev), **kw, **v)
            # N J log(1-eps) is 1e-4 of this loss; the row kernel's lg2.approx of the all-missing word product
            # (always the same input 2^-4, so its half-ulp error does not average out) allows 3e-5 here
            assert_close(got["loss"], ref["loss"], 2e-5, 2e-5, f"loss S={S} {sorted(v)}")
            assert float(got["d_a"].abs().max()) == 0.0

"""GPU: the amortised multivariate-normal guide of VaeIRT, K > 1 (vi.py:685-693) with MvnEncoder's tail
(fc22, the tril scatter, LowerCholeskyTransform: vi.py:444, 449-455, 693) fused into the sampler and its backward
pass (csrc/enc_mvn.cu), through the C ABI (vipsy_mvn_encoder_sample, vipsy_irt_mvn_encoder_elbo_grad).

Arbiter: the fp64 CPU oracle for the ELBO side (oracle.irt_elbo_grad on the dense raw factor) with MvnEncoder's
fc22 restated in fp64 torch for the encoder side.  Tolerance 2e-5 of each tensor's largest magnitude (the K > 1
tolerance of DESIGN.md section 2), written below."""
import pytest
import torch

from oracle import vipsy_oracle as o
from tests.util import assert_close

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


def make(B, J, K, H, S, seed, model="irt_2pl", missing=0.1, wscale=0.2):
    g = torch.Generator().manual_seed(seed)
    P = K * (K + 1) // 2
    y = (torch.rand(B, J, generator=g) < 0.6).float()
    y[torch.rand(B, J, generator=g) < missing] = float("nan")
    d = dict(y=y, hidden=torch.nn.functional.softplus(torch.randn(B, H, generator=g)),
             w22=(torch.rand(P, H, generator=g) - 0.5) * wscale, b22=(torch.rand(P, generator=g) - 0.5) * wscale,
             mu=torch.randn(B, K, generator=g) * 0.5, a=torch.rand(K, J, generator=g) * 0.6,
             b=torch.randn(1, J, generator=g), eps=torch.randn(S, B, K, generator=g))
    if model in ("irt_3pl", "irt_4pl"):
        d["c"] = torch.rand(1, J, generator=g) * 0.2
    if model == "irt_4pl":
        d["d"] = 1 - torch.rand(1, J, generator=g) * 0.2
    return d


def oracle_side(d, K, model, scale, a_free=None):
    """fp64: raw = fc22(hidden) -> dense factor -> oracle ELBO + grads -> back through fc22."""
    hd, wd, bd = (d[n].double().requires_grad_(True) for n in ("hidden", "w22", "b22"))
    packed = hd @ wd.T + bd
    raw = o.tril_from_packed(packed, K)
    ref = o.irt_elbo_grad(d["y"], d["mu"], raw.detach(), d["eps"], d["a"], d["b"], d.get("c"), d.get("d"), model=model,
                          scale=scale, chol_raw=True, a_free=a_free)
    ti = torch.tril_indices(K, K)
    torch.autograd.backward([packed], [ref["d_scale"][:, ti[0], ti[1]].double()])
    ref.update(d_hidden=hd.grad, d_w22=wd.grad, d_b22=bd.grad)
    return ref


@pytest.mark.parametrize("B,K,H,S,inject", [(300, 10, 16, 1, True), (128, 8, 64, 1, False), (777, 37, 20, 2, True),
                                            (500, 3, 8, 1, False), (1000, 100, 64, 1, False), (260, 128, 64, 1, True),
                                            (64, 100, 64, 3, False), (5, 2, 1, 1, True)])
def test_sample_site_matches_fp64(ops, B, K, H, S, inject):
    """x = loc + L eps with L from fc22(hidden): the kernel's draws (the library's Philox stream, or injected) pushed
    through the fp64 restatement of vi.py:449-455 + LowerCholeskyTransform."""
    d = make(B, 4, K, H, S, seed=B + K)
    r = ops.mvn_encoder_sample(d["mu"].cuda(), d["hidden"].cuda(), d["w22"].cuda(), d["b22"].cuda(),
                               eps=d["eps"].cuda() if inject else None, n_particles=S, seed=11, offset=4, person_offset=3)
    eps = r["eps"].cpu()
    if inject:
        assert torch.equal(eps, d["eps"])
    else:
        assert torch.equal(r["eps"], ops.fill_normal(S, B, K, 11, 4, person_offset=3))     # the same stream as every other kernel
    raw = o.tril_from_packed(d["hidden"].double() @ d["w22"].double().T + d["b22"].double(), K)
    L = o.lower_cholesky(raw)
    x = d["mu"].double().unsqueeze(0) + torch.einsum("bij,sbj->sbi", L, eps.double())
    assert_close(r["x"], x, TOL, TOL, "x")
    assert_close(r["diag"], torch.diagonal(L, dim1=1, dim2=2), TOL, TOL, "diag L")
    assert_close(r["logdet"], torch.diagonal(raw, dim1=1, dim2=2).sum(1), TOL, TOL, "log det")


def test_sample_site_with_location_head(ops):
    """fc21 inside vipsy_mvn_encoder_sample: the same draws, x shifted by fc21(hidden) instead of a given mu."""
    B, K, H = 300, 21, 24
    d = make(B, 4, K, H, 2, seed=77)
    g = torch.Generator().manual_seed(3)
    w21, b21 = (torch.rand(K, H, generator=g) - 0.5) * 0.4, torch.randn(K, generator=g)
    loc = d["hidden"].double() @ w21.double().T + b21.double()
    cu = {k: v.cuda() for k, v in d.items()}
    r = ops.mvn_encoder_sample(None, cu["hidden"], cu["w22"], cu["b22"], w21=w21.cuda(), b21=b21.cuda(), n_particles=2, seed=4, offset=1)
    r0 = ops.mvn_encoder_sample(loc.float().cuda(), cu["hidden"], cu["w22"], cu["b22"], n_particles=2, seed=4, offset=1)
    assert torch.equal(r["eps"], r0["eps"]) and torch.equal(r["diag"], r0["diag"])
    raw = o.tril_from_packed(d["hidden"].double() @ d["w22"].double().T + d["b22"].double(), K)
    x = loc.unsqueeze(0) + torch.einsum("bij,sbj->sbi", o.lower_cholesky(raw), r["eps"].cpu().double())
    assert_close(r["x"], x, TOL, TOL, "x")


@pytest.mark.parametrize("B,J,K,H,S,model", [(300, 40, 10, 16, 1, "irt_2pl"), (128, 20, 8, 64, 1, "irt_2pl"),
                                             (257, 33, 37, 20, 2, "irt_2pl"), (500, 12, 3, 8, 1, "irt_2pl"),
                                             (200, 40, 12, 32, 3, "irt_3pl"), (130, 70, 9, 64, 1, "irt_4pl"),
                                             (150, 150, 24, 64, 1, "irt_2pl")])
def test_fused_guide_matches_oracle(ops, B, J, K, H, S, model):
    d = make(B, J, K, H, S, seed=J + K, model=model)
    free = torch.ones(K, J, dtype=torch.bool)
    for i in range(K):                                       # the staircase mask of vi.py:568-572
        free[i, J - i:] = False
    ref = oracle_side(d, K, model, 1.7, a_free=free)
    cu = {k: v.cuda() for k, v in d.items()}
    r = ops.irt_elbo_grad(ops.pack_responses(cu["y"]), cu["mu"], mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"]),
                          eps=cu["eps"], a=cu["a"], b=cu["b"], c=cu.get("c"), d=cu.get("d"), model=model, scale=1.7,
                          a_free=free.cuda())
    assert_close(r["loss"], ref["loss"], TOL, TOL, "loss")
    for n in ("d_mu", "d_a", "d_b", "d_hidden", "d_w22", "d_b22") + (("d_c",) if "c" in d else ()) + (("d_d",) if "d" in d else ()):
        assert_close(r[n], ref[n], TOL, TOL, n)
    assert not r["d_a"].cpu()[~free].any()


@pytest.mark.parametrize("B,J,K,H,S", [(1000, 100, 100, 64, 1), (500, 30, 4, 16, 2), (333, 20, 12, 8, 3),
                                       (200, 25, 6, 16, 2), (129, 40, 32, 64, 1), (1, 7, 8, 3, 1),
                                       (150, 30, 37, 20, 2), (300, 60, 50, 64, 1), (260, 20, 128, 64, 1)])
def test_fused_guide_equals_materialised_factor_k100(ops, B, J, K, H, S):
    """The library's own noise stream (K % 4 == 0: the draws travel to the backward kernels as pre-split fp16 pair
    words by bulk copy; otherwise as fp32 rows).  K = 100 is the BASELINE c3 shape scaled down: the fused path against
    the three-kernel path fed with the factor matrix fc22 would have written to HBM (fp64 fc22 and fp64 backward
    through it as the arbiter of the encoder side)."""
    d = make(B, J, K, H, S, seed=5)
    cu = {k: v.cuda() for k, v in d.items()}
    codes = ops.pack_responses(cu["y"])
    # poison what the allocator hands out next: the kernels must not depend on stale workspace / shared-memory contents
    junk = torch.full((1 << 22,), float("nan"), device="cuda")
    del junk
    kw = dict(a=cu["a"], b=cu["b"], model="irt_2pl", scale=200.0, seed=3, offset=9, n_particles=S)
    r = ops.irt_elbo_grad(codes, cu["mu"], mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"]), **kw)
    hd, wd, bd = (cu[n].double().requires_grad_(True) for n in ("hidden", "w22", "b22"))
    packed = hd @ wd.T + bd
    ref = ops.irt_elbo_grad(codes, cu["mu"], chol=packed.detach().float(), chol_packed=True, chol_raw=True, **kw)
    torch.autograd.backward([packed], [ref["d_chol"].double()])
    assert_close(r["loss"], ref["loss"], TOL, TOL, "loss")
    for n in ("d_mu", "d_a", "d_b"):
        assert_close(r[n], ref[n], TOL, TOL, n)
    assert_close(r["d_hidden"], hd.grad, TOL, TOL, "d_hidden")
    assert_close(r["d_w22"], wd.grad, TOL, TOL, "d_w22")
    assert_close(r["d_b22"], bd.grad, TOL, TOL, "d_b22")
    again = ops.irt_elbo_grad(codes, cu["mu"], mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"]), **kw)
    for n in ("loss", "d_mu", "d_a", "d_b", "d_hidden", "d_w22", "d_b22"):
        assert torch.equal(r[n], again[n]), f"{n} is not bit-identical run to run"


@pytest.mark.parametrize("B,J,K,H,S,model", [(300, 40, 10, 16, 1, "irt_2pl"), (257, 33, 37, 20, 2, "irt_2pl"), (500, 12, 3, 8, 1, "irt_2pl"),
                                             (130, 70, 9, 64, 1, "irt_4pl"), (260, 30, 100, 64, 1, "irt_2pl"), (64, 20, 128, 33, 1, "irt_2pl")])
def test_fused_location_head_matches_oracle(ops, B, J, K, H, S, model):
    """fc21 inside the call as well (mvn_encoder with five entries): the location rides in the padding slots of the
    diagonal units.  fp64 side: loc = fc21(hidden), raw = fc22(hidden), oracle ELBO, autograd back through both heads."""
    d = make(B, J, K, H, S, seed=3 * J + K, model=model)
    g = torch.Generator().manual_seed(K)
    d["w21"] = (torch.rand(K, H, generator=g) - 0.5) * 0.3
    d["b21"] = (torch.rand(K, generator=g) - 0.5) * 0.3
    hd, wd, bd, w1, b1 = (d[n].double().requires_grad_(True) for n in ("hidden", "w22", "b22", "w21", "b21"))
    packed, loc = hd @ wd.T + bd, hd @ w1.T + b1
    ref = o.irt_elbo_grad(d["y"], loc.detach(), o.tril_from_packed(packed, K).detach(), d["eps"], d["a"], d["b"], d.get("c"),
                          d.get("d"), model=model, scale=1.3, chol_raw=True)
    ti = torch.tril_indices(K, K)
    torch.autograd.backward([packed, loc], [ref["d_scale"][:, ti[0], ti[1]].double(), ref["d_mu"].double()])
    cu = {k: v.cuda() for k, v in d.items()}
    r = ops.irt_elbo_grad(ops.pack_responses(cu["y"]), None,
                          mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"], cu["w21"], cu["b21"]),
                          eps=cu["eps"], a=cu["a"], b=cu["b"], c=cu.get("c"), d=cu.get("d"), model=model, scale=1.3)
    assert_close(r["loss"], ref["loss"], TOL, TOL, "loss")
    for n in ("d_mu", "d_a", "d_b"):
        assert_close(r[n], ref[n], TOL, TOL, n)
    for n, t in (("d_hidden", hd), ("d_w22", wd), ("d_b22", bd), ("d_w21", w1), ("d_b21", b1)):
        assert_close(r[n], t.grad, TOL, TOL, n)
    # the library's noise stream: same call twice bit-identical, and equal to the three-entry form fed with loc
    kw = dict(a=cu["a"], b=cu["b"], c=cu.get("c"), d=cu.get("d"), model=model, scale=1.3, seed=5, offset=2, n_particles=S)
    codes = ops.pack_responses(cu["y"])
    r1 = ops.irt_elbo_grad(codes, None, mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"], cu["w21"], cu["b21"]), **kw)
    r2 = ops.irt_elbo_grad(codes, None, mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"], cu["w21"], cu["b21"]), **kw)
    r3 = ops.irt_elbo_grad(codes, loc.detach().float().cuda(), mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"]), **kw)
    for n in ("loss", "d_mu", "d_hidden", "d_w21", "d_b21", "d_w22"):
        assert torch.equal(r1[n], r2[n]), n
    assert_close(r1["loss"], r3["loss"], TOL, TOL, "loss (five vs three entries)")
    assert_close(r1["d_mu"], r3["d_mu"], TOL, TOL, "d_mu (five vs three entries)")
    assert_close(r1["d_w22"], r3["d_w22"], TOL, TOL, "d_w22 (five vs three entries)")


def test_value_only_and_minibatch(ops):
    """evaluate_loss (no gradient pointers) gives the same loss; a minibatch through idx reads the right rows."""
    B, J, K, H, N = 96, 30, 6, 16, 400
    d = make(N, J, K, H, 1, seed=8)
    g = torch.Generator().manual_seed(1)
    idx = torch.randperm(N, generator=g)[:B]
    sub = {k: (v[idx] if k in ("hidden", "mu") else v) for k, v in d.items()}
    sub["eps"] = d["eps"][:, :B]
    sub["y"] = d["y"][idx]
    ref = oracle_side(sub, K, "irt_2pl", N / B)
    cu = {k: v.cuda() for k, v in sub.items()}
    codes = ops.pack_responses(d["y"].cuda())
    kw = dict(mvn_encoder=(cu["hidden"], cu["w22"], cu["b22"]), eps=cu["eps"], a=cu["a"], b=cu["b"], scale=N / B,
              idx=idx.cuda())
    r = ops.irt_elbo_grad(codes, cu["mu"], **kw)
    v = ops.irt_elbo_grad(codes, cu["mu"], want_grad=False, **kw)
    assert_close(r["loss"], ref["loss"], TOL, TOL, "loss")
    assert float(v["loss"]) == float(r["loss"]) and "d_w22" not in v
    for n in ("d_mu", "d_a", "d_b", "d_hidden", "d_w22", "d_b22"):
        assert_close(r[n], ref[n], TOL, TOL, n)


def test_bad_arguments(ops):
    from vipsy_b200._native import VipsyError
    d = {k: v.cuda() for k, v in make(40, 10, 4, 8, 1, seed=2).items()}
    codes = ops.pack_responses(d["y"])
    with pytest.raises(ValueError):
        ops.irt_elbo_grad(codes, d["mu"], mvn_encoder=(d["hidden"][:10], d["w22"], d["b22"]), a=d["a"], b=d["b"])
    with pytest.raises(ValueError):
        ops.irt_elbo_grad(codes, d["mu"], mvn_encoder=(torch.zeros(40, 65).cuda(), torch.zeros(10, 65).cuda(), d["b22"]),
                          a=d["a"], b=d["b"])
    with pytest.raises(VipsyError):
        ops.irt_elbo_grad(codes, d["mu"], mvn_encoder=(d["hidden"], d["w22"], d["b22"]), a=d["a"], b=d["b"],
                          estimator="score")

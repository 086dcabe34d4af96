"""GPU parity: the fused K=1 IRT kernel (through the C ABI) against the oracle."""
import numpy as np
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


def make_case(B, J, S, model, missing=0.0, seed=0, sat=False):
    g = torch.Generator().manual_seed(seed)
    y = (torch.rand(B, J, generator=g) < 0.55).float()
    if missing:
        y[torch.rand(B, J, generator=g) < missing] = float("nan")
    mu = torch.randn(B, 1, generator=g) * (3.0 if sat else 0.7)
    sigma = torch.exp(torch.randn(B, 1, generator=g) * 0.3)
    eps = torch.randn(S, B, 1, generator=g)
    a = torch.rand(1, J, generator=g) * 2.5 + 0.5
    b = torch.randn(1, J, generator=g) * (3.0 if sat else 1.0)
    c = torch.rand(1, J, generator=g) * 0.15 + 0.05
    d = torch.rand(1, J, generator=g) * 0.15 + 0.8
    return dict(y=y, mu=mu, sigma=sigma, eps=eps, a=a, b=b, c=c, d=d, model=model)


def run_both(ops, cs, D=1.0, scale=1.0, estimator="pathwise", baseline=0.0, strict=False, idx=None, y_full=None,
             sparse=False):
    m = cs["model"]
    dev = "cuda"
    kw = dict(a=cs["a"] if m != "irt_1pl" else None, b=cs["b"],
              c=cs["c"] if m in ("irt_3pl", "irt_4pl") else None, d=cs["d"] if m == "irt_4pl" else None)
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], cs["sigma"], cs["eps"], kw["a"], kw["b"], kw["c"], kw["d"],
                          model=m, D=D, scale=scale, estimator=estimator, baseline=baseline)
    ysrc = cs["y"] if y_full is None else y_full
    codes = codes_from_float(ysrc).to(dev)
    got = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sigma=cs["sigma"].to(dev), eps=cs["eps"].to(dev),
                            **{k: (v.to(dev) if v is not None else None) for k, v in kw.items()},
                            model=m, D=D, scale=scale, estimator=estimator, baseline=baseline, strict=strict,
                            idx=None if idx is None else idx.to(dev), sparse=sparse)
    torch.cuda.synchronize()
    return got, ref


def check(got, ref, model, tol=1e-5):
    assert_close(got["loss"], ref["loss"], tol, tol, "loss")
    assert_close(got["d_mu"], ref["d_mu"], tol, tol, "d_mu")
    assert_close(got["d_sigma"], ref["d_scale"], tol, tol, "d_sigma")
    assert_close(got["d_b"], ref["d_b"], tol, tol, "d_b")
    if model != "irt_1pl":
        assert_close(got["d_a"], ref["d_a"], tol, tol, "d_a")
    if model in ("irt_3pl", "irt_4pl"):
        assert_close(got["d_c"], ref["d_c"], tol, tol, "d_c")
    if model == "irt_4pl":
        assert_close(got["d_d"], ref["d_d"], tol, tol, "d_d")


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("S", [1, 2, 3, 4, 8, 20, 32, 64])
def test_parity_models_and_particles(ops, model, S):
    cs = make_case(B=203, J=37, S=S, model=model, missing=0.15, seed=S)
    got, ref = run_both(ops, cs, D=1.0, scale=2.5)
    check(got, ref, model)


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
def test_parity_with_D_1702(ops, model):
    """D = 1.702 pushes |logit| to ~15: 1-P must not be formed by cancellation."""
    cs = make_case(B=203, J=37, S=4, model=model, missing=0.15, seed=77)
    got, ref = run_both(ops, cs, D=1.702, scale=2.5)
    check(got, ref, model)


@pytest.mark.parametrize("model", ["irt_2pl", "irt_4pl"])
@pytest.mark.parametrize("B,J", [(1, 1), (31, 5), (32, 100), (33, 7), (1000, 5), (4097, 100), (257, 333)])
def test_ragged_shapes(ops, model, B, J):
    cs = make_case(B=B, J=J, S=4, model=model, missing=0.1, seed=B + J)
    got, ref = run_both(ops, cs)
    check(got, ref, model)


def test_strict_and_fast_4pl_agree(ops):
    cs = make_case(B=777, J=100, S=32, model="irt_4pl", seed=5)
    fast, ref = run_both(ops, cs)
    strict, _ = run_both(ops, cs, strict=True)
    check(fast, ref, "irt_4pl")
    check(strict, ref, "irt_4pl")


def test_fast_path_falls_back_when_items_touch_the_clamp(ops):
    cs = make_case(B=300, J=20, S=8, model="irt_4pl", seed=6, sat=True)
    cs["c"][0, :5] = 0.0          # lower asymptote exactly 0: clamp reachable
    cs["d"][0, 5:9] = 1.0
    got, ref = run_both(ops, cs)
    check(got, ref, "irt_4pl", tol=2e-5)


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl"])
def test_saturated_regime(ops, model):
    """Large |logit|: clamp fires, zero-gradient cells, reference formula class."""
    cs = make_case(B=500, J=40, S=4, model=model, seed=11, sat=True)
    got, ref = run_both(ops, cs)
    check(got, ref, model, tol=2e-5)


@pytest.mark.parametrize("estimator,baseline", [("score", 0.0), ("score", -35.0)])
@pytest.mark.parametrize("model", ["irt_2pl", "irt_4pl"])
def test_score_function_estimator(ops, model, estimator, baseline):
    cs = make_case(B=301, J=29, S=8, model=model, missing=0.1, seed=3)
    got, ref = run_both(ops, cs, scale=1.5, estimator=estimator, baseline=baseline)
    check(got, ref, model, tol=2e-5)


def test_all_missing_rows_and_items(ops):
    cs = make_case(B=100, J=12, S=2, model="irt_2pl", missing=0.3, seed=9)
    cs["y"][:40] = float("nan")        # a whole warp of persons unobserved
    cs["y"][:, 3] = float("nan")       # an item nobody answered
    got, ref = run_both(ops, cs)
    check(got, ref, "irt_2pl")
    assert float(got["d_b"][0, 3]) == 0.0 and float(got["d_a"][0, 3]) == 0.0


def test_idx_gather_matches_contiguous(ops):
    cs = make_case(B=150, J=23, S=4, model="irt_3pl", missing=0.1, seed=21)
    g = torch.Generator().manual_seed(1)
    y_full = (torch.rand(900, 23, generator=g) < 0.5).float()
    idx = torch.randperm(900, generator=g)[:150]
    y_full[idx] = cs["y"]
    got, ref = run_both(ops, cs, idx=idx, y_full=y_full, scale=6.0)
    check(got, ref, "irt_3pl")


def test_golden_fixtures_through_the_abi(ops, golden):
    g = golden
    for name, model in [("irt1", "irt_1pl"), ("irt2", "irt_2pl"), ("irt3", "irt_3pl"), ("irt4", "irt_4pl")]:
        K, S, D, scale = g[f"{name}_meta"]
        t = lambda k: torch.from_numpy(g[f"{name}_{k}"])
        dev = "cuda"
        got = ops.irt_elbo_grad(codes_from_float(t("y")).to(dev), t("mu").to(dev), sigma=t("sp").to(dev),
                                eps=t("eps").to(dev), a=t("a").to(dev), b=t("b").to(dev), c=t("c").to(dev),
                                d=t("d").to(dev), model=model, D=float(D), scale=float(scale))
        tol = 2e-5   # golden is the reference's fp32 arithmetic
        assert_close(got["loss"], t("loss"), tol, tol, name)
        assert_close(got["d_mu"], t("d_mu"), tol, tol, name)
        assert_close(got["d_sigma"], t("d_sp"), tol, tol, name)
        assert_close(got["d_b"], t("d_b"), tol, tol, name)
        if model != "irt_1pl":
            assert_close(got["d_a"], t("d_a"), tol, tol, name)


def test_lsat_first_iteration_known_answer(ops, golden):
    """BASELINE config 1: lsat.dat 2PL at the reference's init (SURVEY 8c KAT)."""
    y = torch.from_numpy(golden["lsat_y"]).cuda()
    torch.manual_seed(0)
    eps = torch.randn(1000, 1)
    got = ops.irt_elbo_grad(y, torch.zeros(1000, 1).cuda(), sigma=torch.ones(1000, 1).cuda(),
                            eps=eps[None].cuda(), a=torch.ones(1, 5).cuda(), b=torch.zeros(1, 5).cuda(),
                            model="irt_2pl")
    assert abs(float(got["loss"]) - 4056.5154) < 4e-2
    assert_close(got["d_b"], torch.tensor([[-419.5977, -204.5977, -48.5977, -258.5977, -365.5977]]), 1e-5, 1e-5)
    assert_close(got["d_a"], torch.tensor([[205.3315, 198.5405, 226.3220, 225.7764, 212.9940]]), 1e-5, 1e-5)


def test_bitwise_deterministic(ops):
    cs = make_case(B=5000, J=64, S=8, model="irt_4pl", missing=0.05, seed=2)
    a, _ = run_both(ops, cs)
    b, _ = run_both(ops, cs)
    for k in ("loss", "d_mu", "d_sigma", "item_grads"):
        assert torch.equal(a[k], b[k]), k


def test_in_kernel_philox_equals_injected_noise(ops):
    cs = make_case(B=999, J=31, S=12, model="irt_4pl", seed=8)
    dev = "cuda"
    eps = ops.fill_normal(12, 999, 1, seed=1234, offset=7, person_offset=100)
    common = dict(sigma=cs["sigma"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), c=cs["c"].to(dev),
                  d=cs["d"].to(dev), model="irt_4pl")
    codes = codes_from_float(cs["y"]).to(dev)
    inj = ops.irt_elbo_grad(codes, cs["mu"].to(dev), eps=eps, **common)
    rng = ops.irt_elbo_grad(codes, cs["mu"].to(dev), n_particles=12, seed=1234, offset=7, person_offset=100, **common)
    for k in ("loss", "d_mu", "d_sigma", "d_b", "d_a", "d_c", "d_d"):
        assert_close(rng[k], inj[k], 1e-6, 1e-6, k)
    e = eps.cpu().double()
    assert abs(float(e.mean())) < 0.03 and abs(float(e.std()) - 1.0) < 0.03
    other = ops.fill_normal(12, 999, 1, seed=1234, offset=8, person_offset=100)
    assert not torch.equal(other, eps)


def test_pack_responses_kernel(ops):
    g = torch.Generator().manual_seed(4)
    y = (torch.rand(1237, 13, generator=g) < 0.5).float()
    y[torch.rand(1237, 13, generator=g) < 0.2] = float("nan")
    assert torch.equal(ops.pack_responses(y.cuda()).cpu(), codes_from_float(y))


def test_loss_only_call(ops):
    cs = make_case(B=100, J=10, S=4, model="irt_2pl", seed=1)
    dev = "cuda"
    full, ref = run_both(ops, cs)
    val = ops.irt_elbo_grad(codes_from_float(cs["y"]).to(dev), cs["mu"].to(dev), sigma=cs["sigma"].to(dev),
                            eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), model="irt_2pl",
                            want_grad=False)
    assert torch.equal(val["loss"], full["loss"]) and "d_mu" not in val


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("S,missing,J", [(1, 0.9, 200), (4, 0.9, 37), (32, 0.95, 101), (3, 0.0, 20), (8, 1.0, 16)])
def test_sparse_compaction_path(ops, model, S, missing, J):
    """VIPSY_IRT_FLAG_SPARSE (BASELINE config 4: 90 % missing): per-person compaction + fixed-point item sums
    give the same loss and gradients as the oracle -- incl. fully observed and fully missing matrices."""
    cs = make_case(B=333, J=J, S=S, model=model, missing=missing, seed=S + J)
    if missing >= 1.0:
        cs["y"][:] = float("nan")
    got, ref = run_both(ops, cs, D=1.702, scale=4.0, sparse=True)
    check(got, ref, model)


def test_sparse_matches_dense_and_is_deterministic(ops):
    cs = make_case(B=4000, J=200, S=2, model="irt_2pl", missing=0.9, seed=5)
    dev = "cuda"
    codes = codes_from_float(cs["y"]).to(dev)
    kw = dict(sigma=cs["sigma"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), model="irt_2pl", n_particles=2,
              seed=3, offset=9)
    dense = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    sp1 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sparse=True, **kw)
    sp2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sparse=True, **kw)
    for k in ("loss", "d_mu", "d_sigma", "d_a", "d_b"):
        assert_close(sp1[k], dense[k], 2e-6, 2e-6, k)
        assert torch.equal(sp1[k], sp2[k]), k


def test_sparse_idx_gather(ops):
    cs = make_case(B=150, J=24, S=4, model="irt_3pl", missing=0.8, seed=21)
    g = torch.Generator().manual_seed(1)
    y_full = (torch.rand(600, 24, generator=g) < 0.5).float()
    idx = torch.randperm(600, generator=g)[:150]
    y_full[idx] = cs["y"]
    got, ref = run_both(ops, cs, scale=4.0, idx=idx, y_full=y_full, sparse=True)
    check(got, ref, "irt_3pl")


# ---- S == 1 row kernels (irt_uni_row.cu): warp per person, lanes over items ---------------------------
@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("J,missing,sparse", [(64, 0.0, False), (100, 0.15, False), (200, 0.9, True), (200, 0.9, False),
                                              (256, 0.3, True), (132, 0.0, True), (128, 1.0, True), (128, 1.0, False)])
def test_row_kernels_parity(ops, model, J, missing, sparse):
    """One particle, J % 4 == 0, 64 <= J: dense (register-resident item sums) and sparse (ballot compaction)
    row kernels against the oracle, incl. fully observed / fully missing matrices and ragged last tiles."""
    cs = make_case(B=1237, J=J, S=1, model=model, missing=missing, seed=J + int(10 * missing))
    if missing >= 1.0:
        cs["y"][:] = float("nan")
    got, ref = run_both(ops, cs, D=1.702, scale=4.0, sparse=sparse)
    check(got, ref, model)


@pytest.mark.parametrize("model,J", [("irt_2pl", 512), ("irt_2pl", 388), ("irt_1pl", 300), ("irt_4pl", 256)])
@pytest.mark.parametrize("sparse", [False, True])
def test_row_kernels_wide(ops, model, J, sparse):
    cs = make_case(B=515, J=J, S=1, model=model, missing=0.5, seed=J)
    got, ref = run_both(ops, cs, scale=2.0, sparse=sparse)
    check(got, ref, model)


@pytest.mark.parametrize("sparse", [False, True])
def test_row_kernels_strict_4pl_and_saturation(ops, sparse):
    cs = make_case(B=700, J=100, S=1, model="irt_4pl", missing=0.2, seed=3, sat=True)
    cs["c"][0, :5] = 0.0          # items that can reach the clamp: the rational form must switch itself off
    cs["d"][0, 5:9] = 1.0
    got, ref = run_both(ops, cs, D=1.702, scale=1.5, sparse=sparse)
    check(got, ref, "irt_4pl", tol=2e-5)
    got, ref = run_both(ops, cs, D=1.702, scale=1.5, sparse=sparse, strict=True)
    check(got, ref, "irt_4pl", tol=2e-5)


@pytest.mark.parametrize("sparse", [False, True])
def test_row_kernels_idx_gather_and_loss_only(ops, sparse):
    cs = make_case(B=150, J=96, S=1, model="irt_3pl", missing=0.6, seed=21)
    g = torch.Generator().manual_seed(1)
    y_full = (torch.rand(600, 96, generator=g) < 0.5).float()
    idx = torch.randperm(600, generator=g)[:150]
    y_full[idx] = cs["y"]
    got, ref = run_both(ops, cs, scale=4.0, idx=idx, y_full=y_full, sparse=sparse)
    check(got, ref, "irt_3pl")
    dev = "cuda"
    val = ops.irt_elbo_grad(codes_from_float(y_full).to(dev), cs["mu"].to(dev), sigma=cs["sigma"].to(dev),
                            eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), c=cs["c"].to(dev),
                            model="irt_3pl", scale=4.0, idx=idx.to(dev), sparse=sparse, want_grad=False)
    assert torch.equal(val["loss"], got["loss"]) and "d_mu" not in val


def test_row_kernels_match_lane_kernels_same_philox_stream(ops):
    """In-kernel noise: the row kernels draw particle 0 of the same Philox stream as the lane-per-person
    kernels, so the two mappings agree to rounding on every output; and they are bit-reproducible."""
    cs = make_case(B=5000, J=200, S=1, model="irt_2pl", missing=0.9, seed=5)
    dev = "cuda"
    codes = codes_from_float(cs["y"]).to(dev)
    kw = dict(sigma=cs["sigma"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), model="irt_2pl", n_particles=1,
              seed=3, offset=9, person_offset=17)
    lane = ops.irt_elbo_grad(codes, cs["mu"].to(dev), row_kernel=False, **kw)
    for sparse in (False, True):
        r1 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sparse=sparse, **kw)
        r2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sparse=sparse, **kw)
        for k in ("loss", "d_mu", "d_sigma", "d_a", "d_b"):
            assert_close(r1[k], lane[k], 2e-6, 2e-6, k)
            assert torch.equal(r1[k], r2[k]), k


# ---- observed-cell CSR layout (irt_uni_csr.cu) -----------------------------------------------------------
def test_observed_csr_layout_is_exact(ops):
    g = torch.Generator().manual_seed(4)
    y = (torch.rand(777, 53, generator=g) < 0.5).float()
    y[torch.rand(777, 53, generator=g) < 0.8] = float("nan")
    y[5] = float("nan")                                   # an empty row
    y[6] = 1.0                                            # a full row
    codes = codes_from_float(y)
    rp, en, nnz = ops.observed_csr(codes.cuda())
    rp, en = rp.cpu(), en.cpu().to(torch.int32)
    obs = codes != 255
    assert nnz == int(obs.sum()) and torch.equal(rp[1:] - rp[:-1], obs.sum(1))
    rows, cols = obs.nonzero(as_tuple=True)               # row-major = item order within a row
    assert torch.equal(en[:nnz] & 0x7fff, cols.to(torch.int32))
    assert torch.equal((en[:nnz] >> 15).to(torch.uint8), codes[rows, cols])


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("J,missing", [(200, 0.9), (37, 0.5), (300, 0.97), (64, 0.0), (50, 1.0)])
def test_csr_path_parity(ops, model, J, missing):
    """S == 1 over the CSR of observed cells: every sub-warp width (8 / 16 / 32 lanes per person), empty and
    full rows, ragged last tile."""
    cs = make_case(B=1237, J=J, S=1, model=model, missing=missing, seed=J)
    if missing >= 1.0:
        cs["y"][:] = float("nan")
    m = cs["model"]
    dev = "cuda"
    kw = dict(a=cs["a"] if m != "irt_1pl" else None, b=cs["b"],
              c=cs["c"] if m in ("irt_3pl", "irt_4pl") else None, d=cs["d"] if m == "irt_4pl" else None)
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], cs["sigma"], cs["eps"], kw["a"], kw["b"], kw["c"], kw["d"],
                          model=m, D=1.702, scale=3.0)
    codes = codes_from_float(cs["y"]).to(dev)
    got = ops.irt_elbo_grad(codes, cs["mu"].to(dev), sigma=cs["sigma"].to(dev), eps=cs["eps"].to(dev),
                            **{k: (v.to(dev) if v is not None else None) for k, v in kw.items()},
                            model=m, D=1.702, scale=3.0, obs=ops.observed_csr(codes))
    check(got, ref, m)


def test_csr_path_idx_gather_deterministic(ops):
    cs = make_case(B=150, J=96, S=1, model="irt_2pl", missing=0.85, seed=2)
    g = torch.Generator().manual_seed(1)
    y_full = (torch.rand(600, 96, generator=g) < 0.5).float()
    y_full[torch.rand(600, 96, generator=g) < 0.85] = float("nan")
    idx = torch.randperm(600, generator=g)[:150]
    y_full[idx] = cs["y"]
    dev = "cuda"
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], cs["sigma"], cs["eps"], cs["a"], cs["b"], None, None,
                          model="irt_2pl", scale=4.0)
    codes = codes_from_float(y_full).to(dev)
    obs = ops.observed_csr(codes)
    kw = dict(sigma=cs["sigma"].to(dev), eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev),
              model="irt_2pl", scale=4.0, idx=idx.to(dev), obs=obs)
    r1 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    r2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    check(r1, ref, "irt_2pl")
    for k in ("loss", "d_mu", "d_sigma", "d_a", "d_b"):
        assert torch.equal(r1[k], r2[k]), k


@pytest.mark.parametrize("model", ["irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("J", [1, 5, 6, 8, 13, 20, 33, 50, 63, 64])
def test_row_kernel_short_rows(ops, model, J):
    """J <= 64: several persons per warp step (WP = 2, 4, 8, 16 words per row), rows of any length / alignment --
    the reference's own data sets have J = 5 (lsat.dat) and 6 (ex5.2.dat), its article experiment J = 50."""
    cs = make_case(B=1237, J=J, S=1, model=model, missing=0.15, seed=3 * J + 1)
    got, ref = run_both(ops, cs, D=1.702, scale=4.0)
    check(got, ref, model)
    lane, _ = run_both(ops, cs, D=1.702, scale=4.0)          # deterministic
    assert torch.equal(lane["d_b"], got["d_b"]) and torch.equal(lane["d_mu"], got["d_mu"])


@pytest.mark.parametrize("J,sparse", [(101, False), (203, False), (130, True), (66, True), (67, False), (511, False)])
def test_row_kernels_unaligned_rows(ops, J, sparse):
    """J % 4 != 0: every other row starts off a 4-byte boundary (funnel-shifted loads), the last word is partial."""
    cs = make_case(B=777, J=J, S=1, model="irt_2pl", missing=0.6 if sparse else 0.1, seed=J)
    got, ref = run_both(ops, cs, scale=2.0, sparse=sparse)
    check(got, ref, "irt_2pl")
    g = torch.Generator().manual_seed(1)
    y_full = (torch.rand(1500, J, generator=g) < 0.5).float()
    idx = torch.randperm(1500, generator=g)[:777]
    y_full[idx] = cs["y"]
    got2, _ = run_both(ops, cs, scale=2.0, idx=idx, y_full=y_full, sparse=sparse)
    for k in ("loss", "d_mu", "d_a", "d_b"):
        assert torch.equal(got2[k], got[k]), k


def test_large_launch_noise_in_shared_memory_variant(ops):
    """Launches of at least four rounds of the five-CTA grid keep the noise pairs in shared memory (csrc/irt_uni.cu,
    ESM).  The same persons evaluated in two half-size launches take the register variant: per-person outputs must be
    bit-identical (same Philox stream, same arithmetic), item gradients and loss equal to rounding."""
    from vipsy_b200 import workloads
    N, J, S = 420_000, 8, 32
    for model in ("irt_4pl", "irt_2pl"):
        w = workloads.uni_irt(N, J, model, missing=0.1, seed=3)
        kw = {k: w[k] for k in ("a", "b", "c", "d")[:int(model[4])]}
        full = ops.irt_elbo_grad(w["y"], w["mu"], sigma=w["sigma"], n_particles=S, model=model, seed=4, offset=2, **kw)
        h = N // 2
        parts = [ops.irt_elbo_grad(w["y"][s:s + h], w["mu"][s:s + h], sigma=w["sigma"][s:s + h], n_particles=S,
                                   model=model, seed=4, offset=2, person_offset=s, **kw) for s in (0, h)]
        assert torch.equal(full["d_mu"], torch.cat([p["d_mu"] for p in parts]))
        assert torch.equal(full["d_sigma"], torch.cat([p["d_sigma"] for p in parts]))
        assert_close(full["item_grads"], parts[0]["item_grads"].double() + parts[1]["item_grads"].double(), 1e-5, 1e-5, model)
        assert_close(full["loss"], parts[0]["loss"].double() + parts[1]["loss"].double(), 1e-9, 1e-9, model)


@pytest.mark.parametrize("model,S", [("irt_2pl", 32), ("irt_4pl", 8), ("irt_3pl", 2), ("irt_1pl", 4)])
def test_observed_cell_csr_with_several_particles(ops, model, S):
    """S > 1 with the observed-cell CSR given: the sparse kernel walks the CSR entries directly (no staging, no
    compaction).  Same per-person arithmetic in the same order as the compaction variant -> bit-identical per-person
    outputs; item gradients (exact fixed-point sums per CTA, different grids) and loss to rounding; and the dense
    kernel agrees to the parity tolerance."""
    from vipsy_b200 import workloads
    N, J = 20_011, 60
    w = workloads.uni_irt(N, J, model, missing=0.85, seed=21)
    w["y"][5] = 255                                                   # a person without a single response
    kw = {k: w[k] for k in (("b",) if model == "irt_1pl" else ("a", "b", "c", "d")[:int(model[4])])}
    common = dict(sigma=w["sigma"], n_particles=S, model=model, seed=4, offset=2, **kw)
    obs = ops.observed_csr(w["y"])
    csr = ops.irt_elbo_grad(w["y"], w["mu"], obs=obs, **common)
    comp = ops.irt_elbo_grad(w["y"], w["mu"], sparse=True, **common)
    dense = ops.irt_elbo_grad(w["y"], w["mu"], **common)
    assert torch.equal(csr["d_mu"], comp["d_mu"]) and torch.equal(csr["d_sigma"], comp["d_sigma"])
    assert_close(csr["item_grads"], comp["item_grads"], 1e-6, 1e-6, "item grads csr vs compaction")
    assert_close(csr["loss"], comp["loss"], 1e-9, 1e-9, "loss csr vs compaction")
    assert_close(csr["d_mu"], dense["d_mu"], 2e-5, 2e-5, "d_mu csr vs dense")
    assert_close(csr["item_grads"], dense["item_grads"], 2e-5, 2e-5, "item grads csr vs dense")
    assert_close(csr["loss"], dense["loss"], 1e-6, 1e-6, "loss csr vs dense")
    # minibatch: the CSR of the whole matrix is indexed through idx
    idx = torch.randperm(N, device="cuda")[:3001]
    mb = dict(sigma=w["sigma"][idx], n_particles=S, model=model, seed=4, offset=2, idx=idx, scale=N / 3001, **kw)
    csr_mb = ops.irt_elbo_grad(w["y"], w["mu"][idx], obs=obs, **mb)
    comp_mb = ops.irt_elbo_grad(w["y"], w["mu"][idx], sparse=True, **mb)
    assert torch.equal(csr_mb["d_mu"], comp_mb["d_mu"])
    assert_close(csr_mb["item_grads"], comp_mb["item_grads"], 1e-6, 1e-6, "minibatch item grads")
    assert_close(csr_mb["loss"], comp_mb["loss"], 1e-9, 1e-9, "minibatch loss")

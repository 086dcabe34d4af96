"""GPU parity: the K > 1 path (sample -> fused likelihood contraction -> guide gradients)."""
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


def staircase(K, J):
    free = torch.ones(K, J, dtype=torch.bool)
    for i in range(K):
        free[i, J - i:] = False          # vi.py:568-572
    return free


def make(B, J, K, S, model, seed, shared=False, missing=0.1):
    g = torch.Generator().manual_seed(seed)
    free = staircase(K, J)
    a = (torch.rand(K, J, generator=g) * 1.5 + 0.3) / K ** 0.5 * free
    b = torch.randn(1, J, generator=g)
    c = torch.rand(1, J, generator=g) * 0.15 + 0.05
    d = torch.rand(1, J, generator=g) * 0.15 + 0.8
    y = (torch.rand(B, J, generator=g) < 0.55).float()
    if missing:
        y[torch.rand(B, J, generator=g) < missing] = float("nan")
    mu = torch.randn(B, K, generator=g) * 0.6
    raw = torch.randn((K, K) if shared else (B, K, K), generator=g) * 0.15
    eps = torch.randn(S, B, K, generator=g)
    return dict(y=y, mu=mu, raw=raw, eps=eps, a=a, b=b, c=c, d=d, free=free, model=model, K=K)


def packed_of(dense):
    K = dense.shape[-1]
    idx = torch.tril_indices(K, K)
    return dense[..., idx[0], idx[1]].contiguous()


def run(ops, cs, layout="dense", raw=True, scale=1.0, D=1.0, idx=None, y_full=None, S=None):
    m, K = cs["model"], cs["K"]
    kw = dict(a=cs["a"], b=cs["b"], c=cs["c"] if m in ("irt_3pl", "irt_4pl") else None,
              d=cs["d"] if m == "irt_4pl" else None)
    chol_in = cs["raw"] if raw else o.lower_cholesky(cs["raw"])
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], chol_in, cs["eps"], kw["a"], kw["b"], kw["c"], kw["d"], model=m,
                          D=D, scale=scale, a_free=cs["free"], chol_raw=raw)
    ref_dchol = ref["d_scale"].tril()
    dev = "cuda"
    chol_dev = (packed_of(chol_in) if layout == "packed" else chol_in).to(dev)
    ysrc = cs["y"] if y_full is None else y_full
    got = ops.irt_elbo_grad(codes_from_float(ysrc).to(dev), cs["mu"].to(dev), chol=chol_dev, eps=cs["eps"].to(dev),
                            **{k: (v.to(dev) if v is not None else None) for k, v in kw.items()}, model=m, D=D,
                            scale=scale, a_free=cs["free"].to(dev), chol_packed=(layout == "packed"), chol_raw=raw,
                            idx=None if idx is None else idx.to(dev))
    torch.cuda.synchronize()
    tol = 2e-5
    assert_close(got["loss"], ref["loss"], tol, tol, "loss")
    assert_close(got["d_mu"], ref["d_mu"], tol, tol, "d_mu")
    gd = got["d_chol"].cpu()
    if layout == "packed":
        assert_close(gd, packed_of(ref_dchol), tol, tol, "d_chol(packed)")
    else:
        assert_close(gd.tril(), ref_dchol, tol, tol, "d_chol")
        assert float(gd.triu(1).abs().max()) == 0.0
    assert_close(got["d_a"], ref["d_a"], tol, tol, "d_a")
    assert float((got["d_a"].cpu() * (~cs["free"])).abs().max()) == 0.0
    assert_close(got["d_b"], ref["d_b"], tol, tol, "d_b")
    if m in ("irt_3pl", "irt_4pl"):
        assert_close(got["d_c"], ref["d_c"], tol, tol, "d_c")
    if m == "irt_4pl":
        assert_close(got["d_d"], ref["d_d"], tol, tol, "d_d")
    return got, ref


@pytest.mark.parametrize("K", [2, 3, 5, 8, 10, 17, 33, 100])
@pytest.mark.parametrize("model", ["irt_2pl", "irt_4pl"])
def test_per_person_packed_raw(ops, K, model):
    cs = make(B=70 if K < 50 else 40, J=max(2 * K, 23) if K < 50 else 130, K=K, S=2, model=model, seed=K)
    run(ops, cs, layout="packed", raw=True, scale=3.0)


@pytest.mark.parametrize("layout,raw", [("dense", True), ("dense", False), ("packed", False)])
@pytest.mark.parametrize("shared", [False, True])
def test_layouts_and_shared_covariance(ops, layout, raw, shared):
    cs = make(B=97, J=41, K=6, S=3, model="irt_3pl", seed=5, shared=shared)
    run(ops, cs, layout=layout, raw=raw, D=1.702)


@pytest.mark.parametrize("B,J,S", [(1, 2, 1), (63, 64, 1), (64, 65, 4), (130, 500, 1), (9, 7, 9)])
def test_ragged_tiles(ops, B, J, S):
    cs = make(B=B, J=J, K=min(4, J), S=S, model="irt_2pl", seed=B + J)
    run(ops, cs, layout="packed", raw=True)


def test_idx_gather(ops):
    cs = make(B=50, J=30, K=3, S=2, model="irt_2pl", seed=9)
    g = torch.Generator().manual_seed(2)
    y_full = (torch.rand(400, 30, generator=g) < 0.5).float()
    idx = torch.randperm(400, generator=g)[:50]
    y_full[idx] = cs["y"]
    run(ops, cs, layout="packed", raw=True, scale=8.0, idx=idx, y_full=y_full)


def test_golden_mirt_through_the_abi(ops, golden):
    t = lambda k: torch.from_numpy(golden[k])
    dev = "cuda"
    for name, model in (("mirt2p", "irt_2pl"), ("mirt4s", "irt_4pl")):
        K, S, D, scale = golden[f"{name}_meta"]
        a = t(f"{name}_a")
        got = ops.irt_elbo_grad(codes_from_float(t(f"{name}_y")).to(dev), t(f"{name}_mu").to(dev),
                                chol=t(f"{name}_sp").to(dev), eps=t(f"{name}_eps").to(dev), a=a.to(dev),
                                b=t(f"{name}_b").to(dev), c=t(f"{name}_c").to(dev), d=t(f"{name}_d").to(dev),
                                model=model, D=float(D), scale=float(scale), a_free=(a != 0).to(dev))
        tol = 3e-5
        assert_close(got["loss"], t(f"{name}_loss"), tol, tol, name)
        assert_close(got["d_mu"], t(f"{name}_d_mu"), tol, tol, name)
        assert_close(got["d_chol"].cpu().tril(), t(f"{name}_d_sp").tril(), tol, tol, name)
        assert_close(got["d_a"], t(f"{name}_d_a"), tol, tol, name)
        assert_close(got["d_b"], t(f"{name}_d_b"), tol, tol, name)


def test_mirt_philox_equals_injected(ops):
    cs = make(B=45, J=20, K=5, S=3, model="irt_2pl", seed=4)
    dev = "cuda"
    eps = ops.fill_normal(3, 45, 5, seed=99, offset=3, person_offset=10)
    common = dict(chol=packed_of(cs["raw"]).to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev), model="irt_2pl",
                  chol_packed=True, chol_raw=True)
    codes = codes_from_float(cs["y"]).to(dev)
    inj = ops.irt_elbo_grad(codes, cs["mu"].to(dev), eps=eps, **common)
    rng = ops.irt_elbo_grad(codes, cs["mu"].to(dev), n_particles=3, seed=99, offset=3, person_offset=10, **common)
    for k in ("loss", "d_mu", "d_chol", "d_a", "d_b"):
        assert_close(rng[k], inj[k], 1e-6, 1e-6, k)


def test_mirt_deterministic(ops):
    cs = make(B=300, J=90, K=12, S=2, model="irt_2pl", seed=6)
    dev = "cuda"
    kw = dict(chol=packed_of(cs["raw"]).to(dev), eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev),
              model="irt_2pl", chol_packed=True, chol_raw=True)
    codes = codes_from_float(cs["y"]).to(dev)
    r1 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    r2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    for k in ("loss", "d_mu", "d_chol", "item_grads"):
        assert torch.equal(r1[k], r2[k]), k


@pytest.mark.parametrize("K,S,layout", [(40, 3, "packed"), (40, 4, "dense"), (100, 5, "packed"), (65, 8, "packed"),
                                        (33, 7, "dense"), (100, 8, "packed")])
def test_streaming_guide_kernels_more_particles(ops, K, S, layout):
    """2 < S <= 8 with a per-person factor and K > 32: the multi-pass sampler (pairs of particles per pass over
    the factor) and the particle-loop guide-gradient stream."""
    cs = make(B=45, J=70, K=K, S=S, model="irt_2pl", seed=K + S)
    run(ops, cs, layout=layout, raw=True, scale=2.0)


@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("model,S", [("irt_2pl", 1), ("irt_3pl", 5), ("irt_4pl", 8), ("irt_2pl", 20)])
def test_fused_small_k_kernel(ops, K, model, S):
    """2 <= K <= 8, per-person factor: the single fused kernel (lane per person, warp-shuffle reductions),
    every particle-chunk remainder, packed and dense layouts, injected noise."""
    cs = make(B=75, J=37, K=K, S=S, model=model, seed=10 * K + S, missing=0.2)
    run(ops, cs, layout="packed" if K % 2 else "dense", raw=bool(S % 2), scale=2.0, D=1.702 if model == "irt_3pl" else 1.0)


def test_fused_small_k_matches_three_kernel_path_on_philox_noise(ops):
    """Same in-kernel Philox stream in both paths (noise_coord's K <= 32 layout): outputs agree to rounding;
    the fused kernel is bit-reproducible."""
    cs = make(B=700, J=60, K=4, S=6, model="irt_2pl", seed=3)
    dev = "cuda"
    kw = dict(chol=packed_of(cs["raw"]).to(dev), chol_packed=True, chol_raw=True, n_particles=6, a=cs["a"].to(dev),
              b=cs["b"].to(dev), model="irt_2pl", a_free=cs["free"].to(dev), seed=11, offset=5, person_offset=9)
    codes = codes_from_float(cs["y"]).to(dev)
    f1 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    f2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    u = ops.irt_elbo_grad(codes, cs["mu"].to(dev), fused=False, **kw)
    for k in ("loss", "d_mu", "d_chol", "d_a", "d_b"):
        assert torch.equal(f1[k], f2[k]), k
        assert_close(f1[k], u[k], 2e-5, 2e-5, k)


@pytest.mark.parametrize("K,model,S", [(2, "irt_2pl", 4), (3, "irt_4pl", 5), (4, "irt_3pl", 8), (6, "irt_2pl", 3), (8, "irt_2pl", 2)])
@pytest.mark.parametrize("raw,layout", [(True, "packed"), (False, "dense")])
def test_fused_small_k_score_function_estimator(ops, K, model, S, raw, layout):
    """REINFORCE for the MultivariateNormal guide (north star (d)): x constant, weights scale (ll + log p - log q)
    - baseline times the score of q (L^-T eps; tril(L^-T eps eps^T) - diag(1/L_ii)); item gradients unchanged."""
    cs = make(B=60, J=30, K=K, S=S, model=model, seed=7 * K + S, missing=0.15)
    m = cs["model"]
    kw = dict(a=cs["a"], b=cs["b"], c=cs["c"] if m in ("irt_3pl", "irt_4pl") else None, d=cs["d"] if m == "irt_4pl" else None)
    chol_in = cs["raw"] if raw else o.lower_cholesky(cs["raw"])
    ref = o.irt_elbo_grad(cs["y"], cs["mu"], chol_in, cs["eps"], kw["a"], kw["b"], kw["c"], kw["d"], model=m, scale=1.5,
                          a_free=cs["free"], chol_raw=raw, estimator="score", baseline=3.0)
    dev = "cuda"
    chol_dev = (packed_of(chol_in) if layout == "packed" else chol_in).to(dev)
    got = ops.irt_elbo_grad(codes_from_float(cs["y"]).to(dev), cs["mu"].to(dev), chol=chol_dev, eps=cs["eps"].to(dev),
                            **{k: (v.to(dev) if v is not None else None) for k, v in kw.items()}, model=m, scale=1.5,
                            a_free=cs["free"].to(dev), chol_packed=(layout == "packed"), chol_raw=raw,
                            estimator="score", baseline=3.0)
    tol = 3e-5
    assert_close(got["loss"], ref["loss"], tol, tol, "loss")
    assert_close(got["d_mu"], ref["d_mu"], tol, tol, "d_mu")
    rd = ref["d_scale"].tril()
    gd = got["d_chol"].cpu()
    assert_close(gd if layout == "dense" else gd, rd if layout == "dense" else packed_of(rd), tol, tol, "d_chol")
    assert_close(got["d_a"], ref["d_a"], tol, tol, "d_a")
    assert_close(got["d_b"], ref["d_b"], tol, tol, "d_b")


def test_bulk_fed_sampler_repeats_bit_for_bit_with_injected_noise(ops):
    """The bulk-copy fed sampler (csrc/mirt_stream.cu) hands successive uses of a shared-memory stage to different
    warp groups.  Injected noise skews the groups (its loads sit in front of the barrier wait); a group that ran a
    ring ahead once took the pending parity of the previous use for its own (wrong draws, then a hang).  Every
    kernel on the path is deterministic, so repeated evaluations must agree bit for bit -- and with the oracle."""
    from vipsy_b200 import workloads
    B, J, K = 60_000, 40, 100
    w = workloads.mirt(B, J, K)
    eps = ops.fill_normal(1, B, K, 3, 1)
    junk = torch.empty(200 << 20, dtype=torch.uint8, device="cuda")
    ref = None
    for i in range(12):
        junk.random_(0, 255)
        r = ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, eps=eps,
                              a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"])
        cur = {k: r[k].clone() for k in ("loss", "d_mu", "d_chol", "d_a", "d_b")}
        if ref is None:
            ref = cur
        for k in cur:
            assert torch.equal(ref[k], cur[k]), (i, k)
    # the draws themselves: x = mu + L eps against fp64 torch on a slice of persons
    n = 2000
    ti = torch.tril_indices(K, K, device="cuda")
    L = torch.zeros(n, K, K, dtype=torch.float64, device="cuda")
    L[:, ti[0], ti[1]] = w["chol_packed_raw"][-n:].double()
    L = L.tril(-1) + torch.diag_embed(torch.diagonal(L, dim1=1, dim2=2).exp())
    x = w["mu"][-n:].double() + torch.einsum("bij,bj->bi", L, eps[0, -n:].double())
    # d mu = -scale (sum_j a_j dlb/dlogit_j - x): check it through the likelihood gradient of the same slice
    y = w["y"][-n:]
    gl = (y.double().clamp(max=1) - torch.sigmoid(x @ w["a"].double() + w["b"].double())) * (y != 255).double()
    d_mu = -(gl @ w["a"].double().T - x)
    assert_close(ref["d_mu"][-n:], d_mu, 2e-5)

"""GPU parity of the tcgen05 contraction (csrc/mirt_tc.cu) behind vipsy_irt_elbo_grad for K >= 8, 2PL:
against the fp64 oracle, against the SIMT contraction, across ragged row / item tiles."""
import pytest
import torch

from tests.test_parity_mirt import make, packed_of, run
from tests.util import assert_close, codes_from_float

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("B,J,K,S", [(300, 300, 20, 2), (129, 128, 8, 1), (128, 129, 15, 3), (260, 500, 100, 1),
                                     (1, 9, 9, 1), (700, 257, 31, 1), (40, 130, 111, 2)])
def test_tc_vs_oracle_ragged_tiles(ops, B, J, K, S):
    cs = make(B=B, J=J, K=K, S=S, model="irt_2pl", seed=B + J + K)
    run(ops, cs, layout="packed", raw=True, scale=2.5, D=1.702 if K % 2 else 1.0)


def test_tc_matches_simt(ops):
    cs = make(B=500, J=200, K=24, S=2, model="irt_2pl", seed=77, missing=0.3)
    dev = "cuda"
    kw = dict(chol=packed_of(cs["raw"]).to(dev), eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev),
              model="irt_2pl", chol_packed=True, chol_raw=True, a_free=cs["free"].to(dev), scale=4.0)
    codes = codes_from_float(cs["y"]).to(dev)
    tc = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    simt = ops.irt_elbo_grad(codes, cs["mu"].to(dev), tensor_cores=False, **kw)
    for k in ("loss", "d_mu", "d_chol", "d_a", "d_b"):
        assert_close(tc[k], simt[k], 5e-6, 5e-6, k)
    tc2 = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    for k in ("loss", "d_mu", "d_chol", "item_grads"):
        assert torch.equal(tc[k], tc2[k]), k          # deterministic


def test_tc_idx_gather_and_shared_cov(ops):
    cs = make(B=150, J=140, K=12, S=2, model="irt_2pl", seed=3, shared=True)
    g = torch.Generator().manual_seed(2)
    y_full = (torch.rand(600, 140, generator=g) < 0.5).float()
    idx = torch.randperm(600, generator=g)[:150]
    y_full[idx] = cs["y"]
    run(ops, cs, layout="dense", raw=True, scale=4.0, idx=idx, y_full=y_full)


def test_tc_saturated_logits(ops):
    """K = 100 at the reference's init (a = ones, vi.py:567): logits of +-30, clamp fires, zero gradient there."""
    cs = make(B=200, J=140, K=100, S=1, model="irt_2pl", seed=11)
    cs["a"] = torch.ones_like(cs["a"]) * cs["free"]
    cs["mu"] = cs["mu"] * 3.0
    run(ops, cs, layout="packed", raw=True)


@pytest.mark.parametrize("B,K,S", [(37, 100, 1), (64, 33, 2), (5, 16, 1), (131, 64, 2), (19, 127, 1), (257, 18, 1)])
def test_streamed_guide_kernels(ops, B, K, S):
    """Warp-per-person streaming sampler / guide-gradient kernels (per-person factors, S <= 2), packed and dense."""
    cs = make(B=B, J=40, K=K, S=S, model="irt_2pl", seed=B + K)
    run(ops, cs, layout="packed", raw=True, scale=2.0)
    run(ops, cs, layout="dense", raw=False, scale=0.5)


@pytest.mark.parametrize("model", ["irt_3pl", "irt_4pl"])
@pytest.mark.parametrize("B,J,K,S", [(300, 140, 10, 2), (129, 257, 20, 1), (64, 500, 100, 1)])
def test_tc_3pl_4pl(ops, model, B, J, K, S):
    """10-dim 3PL / 4PL are the reference's own MIRT settings (test.py:363-416); d_c, d_d come from the
    warp column-sum butterfly of the tensor-core epilogue."""
    cs = make(B=B, J=J, K=K, S=S, model=model, seed=B + J + K)
    run(ops, cs, layout="packed", raw=True, scale=1.5, D=1.702)
    dev = "cuda"
    kw = dict(chol=packed_of(cs["raw"]).to(dev), eps=cs["eps"].to(dev), a=cs["a"].to(dev), b=cs["b"].to(dev),
              c=cs["c"].to(dev), d=cs["d"].to(dev) if model == "irt_4pl" else None, model=model, chol_packed=True,
              chol_raw=True, a_free=cs["free"].to(dev), D=1.702)
    codes = codes_from_float(cs["y"]).to(dev)
    tc = ops.irt_elbo_grad(codes, cs["mu"].to(dev), **kw)
    simt = ops.irt_elbo_grad(codes, cs["mu"].to(dev), tensor_cores=False, **kw)
    for k in ("loss", "d_mu", "d_chol", "d_a", "d_b", "d_c") + (("d_d",) if model == "irt_4pl" else ()):
        assert_close(tc[k], simt[k], 1e-5, 1e-5, k)

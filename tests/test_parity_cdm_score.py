"""GPU parity: the REINFORCE CDM kernel (VCDM / VaeCDM path) vs the oracle."""
import math

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


@pytest.mark.parametrize("model", ["dina", "dino"])
@pytest.mark.parametrize("B,J,K,S", [(1, 1, 1, 1), (50, 9, 3, 2), (333, 60, 10, 4), (64, 100, 5, 1), (20, 40, 20, 3)])
def test_reinforce_parity(ops, model, B, J, K, S):
    g = torch.Generator().manual_seed(B + J + K)
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    y = (torch.rand(B, J, generator=g) < 0.5).float()
    gg = torch.rand(1, J, generator=g) * 0.3 + 0.02
    ss = torch.rand(1, J, generator=g) * 0.3 + 0.02
    phi = torch.rand(B, K, generator=g) * 0.9 + 0.05
    u = torch.rand(S, B, K, generator=g)
    attr = (u < phi).float()
    for baseline in (0.0, -20.0):
        ref = o.cdm_score_elbo_grad(y, q, gg, ss, phi, attr, model=model, scale=2.0, baseline=baseline)
        got = ops.cdm_score_elbo_grad(codes_from_float(y).cuda(), q.cuda(), gg.cuda(), ss.cuda(), phi.cuda(),
                                      u=u.cuda(), model=model, scale=2.0, baseline=baseline)
        torch.cuda.synchronize()
        assert_close(got["loss"], ref["loss"], 1e-5, 1e-5, "loss")
        assert_close(got["d_g"], ref["d_g"], 1e-5, 1e-5, "d_g")
        assert_close(got["d_s"], ref["d_s"], 1e-5, 1e-5, "d_s")
        assert_close(got["d_attr_p"], ref["d_attr_p"], 2e-5, 2e-5, "d_attr_p")


def test_golden_reinforce(ops, golden):
    t = lambda k: torch.from_numpy(golden[k])
    phi, attr = t("rf_phi"), t("rf_attr")
    # reconstruct uniforms that reproduce the recorded attribute draws
    u = torch.where(attr > 0, phi * 0.5, phi + (1 - phi) * 0.5)
    got = ops.cdm_score_elbo_grad(codes_from_float(t("rf_y")).cuda(), t("rf_q").cuda(), t("rf_g").cuda(),
                                  t("rf_s").cuda(), phi.cuda(), u=u.cuda(), model="dina", scale=2.0)
    assert_close(got["d_g"], t("rf_d_g"), 2e-5, 2e-5)
    assert_close(got["d_s"], t("rf_d_s"), 2e-5, 2e-5)
    assert_close(got["d_attr_p"], t("rf_d_phi"), 3e-5, 3e-5)


def test_missing_cell_is_nan_like_the_reference(ops):
    """BaseCDM.model has no NaN handling (vi.py:745-756): a missing response poisons the loss."""
    y = torch.zeros(4, 5); y[2, 3] = float("nan")
    q = torch.ones(2, 5)
    got = ops.cdm_score_elbo_grad(codes_from_float(y).cuda(), q.cuda(), torch.full((1, 5), .1).cuda(),
                                  torch.full((1, 5), .1).cuda(), torch.full((4, 2), .5).cuda(), n_particles=1)
    assert math.isnan(float(got["loss"]))


def test_philox_uniform_equals_injected(ops):
    g = torch.Generator().manual_seed(1)
    B, J, K, S = 200, 30, 6, 3
    q = (torch.rand(K, J, generator=g) < 0.5).float().cuda()
    y = (torch.rand(B, J, generator=g) < 0.5).to(torch.uint8).cuda()
    gg, ss = torch.full((1, J), 0.1).cuda(), torch.full((1, J), 0.2).cuda()
    phi = (torch.rand(B, K, generator=g) * 0.8 + 0.1).cuda()
    u = ops.fill_uniform(S, B, K, seed=5, offset=2, person_offset=7)
    a = ops.cdm_score_elbo_grad(y, q, gg, ss, phi, u=u)
    b = ops.cdm_score_elbo_grad(y, q, gg, ss, phi, n_particles=S, seed=5, offset=2, person_offset=7)
    for k in ("loss", "d_g", "d_s", "d_attr_p"):
        assert torch.equal(a[k], b[k]), k
    assert 0.45 < float(u.mean()) < 0.55

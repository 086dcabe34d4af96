"""The oracle against the vectors produced by the reference's own functions
(tests/golden/make_golden.py).  CPU only; pins oracle L2 + estimator restatement."""
import numpy as np
import pytest
import torch

from oracle import vipsy_oracle as o

T = torch.from_numpy


def t(g, k):
    return T(g[k])


def close(a, b, rtol=2e-6, atol_frac=2e-6):
    a = a.detach().to(torch.float64) if isinstance(a, torch.Tensor) else torch.tensor(a, dtype=torch.float64)
    b = b.to(torch.float64)
    atol = atol_frac * max(float(b.abs().max()), 1e-30)
    assert torch.allclose(a.reshape(b.shape), b, rtol=rtol, atol=atol), float((a.reshape(b.shape) - b).abs().max())


def test_link_functions_bit_exact(golden):
    g = golden
    x1, xk, a1, ak = t(g, "link_x1"), t(g, "link_xk"), t(g, "link_a1"), t(g, "link_ak")
    b, c, d = t(g, "link_b"), t(g, "link_c"), t(g, "link_d")
    assert torch.equal(o.irt_prob("irt_1pl", x1, None, b, D=1.702), t(g, "link_1pl"))
    assert torch.equal(o.irt_prob("irt_2pl", x1, a1, b), t(g, "link_2pl_k1"))
    assert torch.equal(o.irt_prob("irt_2pl", xk, ak, b, D=1.702), t(g, "link_2pl_k3"))
    assert torch.equal(o.irt_prob("irt_3pl", xk, ak, b, c), t(g, "link_3pl_k3"))
    assert torch.equal(o.irt_prob("irt_4pl", x1, a1, b, c, d), t(g, "link_4pl_k1"))
    assert torch.equal(o.irt_prob("irt_4pl", xk, ak, b, c, d, D=1.702), t(g, "link_4pl_k3"))


@pytest.mark.parametrize("K", [3, 5])
def test_cdm_tables_bit_exact(golden, K):
    g = golden
    patt = o.all_patterns(K)
    assert torch.equal(patt, t(g, f"cdm{K}_patterns"))
    q, gg, ss = t(g, f"cdm{K}_q"), t(g, f"cdm{K}_g"), t(g, f"cdm{K}_s")
    for m in ("dina", "dino"):
        p = o.cdm_prob(m, patt, q, gg, ss)
        assert torch.equal(p, t(g, f"cdm{K}_{m}_p"))
        eta = o.cdm_eta(m, patt, q)
        assert set(eta.unique().tolist()) <= {0.0, 1.0}
    # DINO quirk: items needing <= 1 attribute have eta == 0 everywhere
    assert o.cdm_eta("dino", patt, q)[:, :2].abs().sum() == 0
    # DINA: an item needing nothing is always answered "ideally"
    assert o.cdm_eta("dina", patt, q)[:, 0].min() == 1


def test_nan_rule_and_staircase(golden):
    g = golden
    y, x = t(g, "nan_y"), t(g, "nan_x")
    p = o.irt_prob("irt_4pl", x, t(g, "nan_a"), t(g, "nan_b"), t(g, "nan_c"), t(g, "nan_d"))
    p, y0, nan = o._mask_nan(p, y)
    assert torch.equal(p, t(g, "nan_p")) and torch.equal(y0, t(g, "nan_y0"))
    free = t(g, "stair_a_free")
    K, J = free.shape
    for i in range(K):
        assert free[i, J - i:].sum() == 0 and free[i, : J - i].min() == 1


def test_encoder_packed_layout(golden):
    g = golden
    assert torch.equal(o.tril_from_packed(t(g, "enc_packed"), 3), t(g, "enc_tril"))


IRT_CASES = [("irt1", "irt_1pl"), ("irt2", "irt_2pl"), ("irt3", "irt_3pl"), ("irt4", "irt_4pl"),
             ("mirt2p", "irt_2pl"), ("mirt4s", "irt_4pl")]


@pytest.mark.parametrize("name,model", IRT_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_irt_elbo_matches_golden(golden, name, model, dtype):
    g = golden
    K, S, D, scale = g[f"{name}_meta"]
    r = o.irt_elbo_grad(
        t(g, f"{name}_y"), t(g, f"{name}_mu"), t(g, f"{name}_sp"), t(g, f"{name}_eps"),
        None if model == "irt_1pl" else t(g, f"{name}_a"), t(g, f"{name}_b"),
        t(g, f"{name}_c") if model in ("irt_3pl", "irt_4pl") else None,
        t(g, f"{name}_d") if model == "irt_4pl" else None,
        model=model, D=float(D), scale=float(scale), dtype=dtype,
        a_free=(t(g, f"{name}_a") != 0) if K > 1 else None)
    tol = dict(rtol=2e-5, atol_frac=2e-5)   # golden is fp32 torch; arbiter is fp64
    close(r["loss"], t(g, f"{name}_loss"), **tol)
    close(r["d_mu"], t(g, f"{name}_d_mu"), **tol)
    close(r["d_scale"], t(g, f"{name}_d_sp"), **tol)
    close(r["d_b"], t(g, f"{name}_d_b"), **tol)
    if model != "irt_1pl":
        close(r["d_a"], t(g, f"{name}_d_a"), **tol)
    if model in ("irt_3pl", "irt_4pl"):
        close(r["d_c"], t(g, f"{name}_d_c"), **tol)
    if model == "irt_4pl":
        close(r["d_d"], t(g, f"{name}_d_d"), **tol)


@pytest.mark.parametrize("name,model", [("dina", "dina"), ("dino", "dino")])
def test_cdm_enum_matches_golden(golden, name, model):
    g = golden
    r = o.cdm_enum_elbo_grad(t(g, f"{name}_y"), t(g, f"{name}_q"), t(g, f"{name}_g"), t(g, f"{name}_s"),
                             model=model, scale=3.0)
    close(r["loss"], t(g, f"{name}_loss"), 2e-5, 2e-5)
    close(r["d_g"], t(g, f"{name}_d_g"), 2e-5, 2e-5)
    close(r["d_s"], t(g, f"{name}_d_s"), 2e-5, 2e-5)
    assert torch.allclose(r["posterior"].sum(-1), torch.ones(17, dtype=torch.float64))


def test_hodina_matches_golden(golden):
    g = golden
    n = "hodina"
    r = o.hodina_elbo_grad(t(g, f"{n}_y"), t(g, f"{n}_q"), t(g, f"{n}_g"), t(g, f"{n}_s"),
                           t(g, f"{n}_lam0"), t(g, f"{n}_lam1"), t(g, f"{n}_mu"), t(g, f"{n}_sigma"),
                           t(g, f"{n}_eps"), scale=3.0)
    for k in ("loss", "d_g", "d_s", "d_lam0", "d_lam1", "d_mu", "d_sigma"):
        close(r[k], t(g, f"{n}_{k}"), 2e-5, 2e-5)


def test_reinforce_cdm_matches_golden(golden):
    g = golden
    r = o.cdm_score_elbo_grad(t(g, "rf_y"), t(g, "rf_q"), t(g, "rf_g"), t(g, "rf_s"), t(g, "rf_phi"),
                              t(g, "rf_attr"), model="dina", scale=2.0)
    close(r["d_g"], t(g, "rf_d_g"), 2e-5, 2e-5)
    close(r["d_s"], t(g, "rf_d_s"), 2e-5, 2e-5)
    close(r["d_attr_p"], t(g, "rf_d_phi"), 2e-5, 2e-5)


def test_lsat_known_answer(golden):
    """First-iteration known answer recorded in SURVEY.md §8c / BASELINE.md §4."""
    y = T(golden["lsat_y"]).float()
    torch.manual_seed(0)
    eps = torch.randn(1000, 1)
    r = o.irt_elbo_grad(y, torch.zeros(1000, 1), torch.ones(1000, 1), eps[None], torch.ones(1, 5),
                        torch.zeros(1, 5), model="irt_2pl", dtype=torch.float32)
    assert abs(r["loss"] - 4056.5154) < 2e-2
    close(r["d_b"], torch.tensor([[-419.5977, -204.5977, -48.5977, -258.5977, -365.5977]]), 1e-5, 1e-5)
    close(r["d_a"], torch.tensor([[205.3315, 198.5405, 226.3220, 225.7764, 212.9940]]), 1e-5, 1e-5)


def test_score_and_pathwise_agree_in_expectation():
    """the score-function estimator is unbiased for the same gradient."""
    torch.manual_seed(1)
    B, J, S = 4, 6, 20000
    y = (torch.rand(B, J) < 0.5).float()
    mu, sg = torch.randn(B, 1) * 0.3, torch.ones(B, 1) * 0.7
    a, b = torch.rand(1, J) + 0.5, torch.randn(1, J)
    eps = torch.randn(S, B, 1)
    pw = o.irt_elbo_grad(y, mu, sg, eps, a, b, model="irt_2pl")
    sc = o.irt_elbo_grad(y, mu, sg, eps, a, b, model="irt_2pl", estimator="score", baseline=-6.0)
    assert abs(pw["loss"] - sc["loss"]) < 1e-9
    assert torch.allclose(pw["d_mu"], sc["d_mu"], atol=0.15)
    assert torch.allclose(pw["d_b"], sc["d_b"], atol=1e-9)


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert o.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert o.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert o.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)

"""GPU: accuracy at BASELINE.json's full sizes.  The tcgen05 accumulator adds with truncation, so a kernel that
keeps a sum in tensor memory over a whole CTA's worth of tiles drifts with the length of the chain -- a parity
test at a few hundred persons cannot see it.  These run the tensor-core kernels at 200 000 x 500 x K=100 (c3) and
1 000 000 x 100 (c2) and compare the accumulated gradients with fp64 torch on the device.  Tolerance 1e-5 of each
tensor's largest magnitude (half the K > 1 parity tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


def rel(x, y):
    return float((x.detach().double() - y.double()).abs().max() / y.double().abs().max())


@pytest.fixture(scope="module")
def c3(ops):
    from vipsy_b200 import workloads
    B, J, K = 200_000, 500, 100
    w = workloads.mirt(B, J, K)
    w["eps"] = ops.fill_normal(1, B, K, 3, 1)
    return w


def test_c3_item_gradients_of_the_tensor_core_likelihood(ops, c3):
    w = c3
    B, K = w["mu"].shape
    J = w["y"].shape[1]
    dev = w["y"].device
    r = ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, eps=w["eps"],
                          a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"])
    ti = torch.tril_indices(K, K, device=dev)
    da = torch.zeros(K, J, dtype=torch.float64, device=dev)
    db = torch.zeros(1, J, dtype=torch.float64, device=dev)
    a64, b64 = w["a"].double(), w["b"].double()
    for s in range(0, B, 20000):
        e = min(s + 20000, B)
        L = torch.zeros(e - s, K, K, dtype=torch.float64, device=dev)
        L[:, ti[0], ti[1]] = w["chol_packed_raw"][s:e].double()
        L = L.tril(-1) + torch.diag_embed(torch.diagonal(L, dim1=1, dim2=2).exp())
        x = w["mu"][s:e].double() + torch.einsum("bij,bj->bi", L, w["eps"][0, s:e].double())
        y = w["y"][s:e]
        gl = (y.double().clamp(max=1) - torch.sigmoid(x @ a64 + b64)) * (y != 255).double()   # d lb / d logit, 2PL
        da -= x.T @ gl
        db -= gl.sum(0, keepdim=True)
    da *= w["a_free"].double()
    assert rel(r["d_a"], da) < TOL and rel(r["d_b"], db) < TOL, (rel(r["d_a"], da), rel(r["d_b"], db))


def test_c3_fused_guide_gradients(ops, c3):
    w = c3
    B, K = w["mu"].shape
    H, P = 64, K * (K + 1) // 2
    dev = w["y"].device
    torch.manual_seed(0)
    hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev))
    w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.02
    b22 = torch.zeros(P, device=dev)
    rows = torch.arange(K, device=dev)
    b22[rows * (rows + 1) // 2 + rows] = -0.7
    kw = dict(eps=w["eps"], a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"])
    rf = ops.irt_elbo_grad(w["y"], w["mu"], mvn_encoder=(hidden, w22, b22), **kw)
    packed = hidden @ w22.T + b22                   # what fc22 would have written to HBM: 4 GB
    rm = ops.irt_elbo_grad(w["y"], w["mu"], chol=packed, chol_packed=True, chol_raw=True, **kw)
    dw = torch.zeros(P, H, dtype=torch.float64, device=dev)
    dh = torch.empty(B, H, dtype=torch.float64, device=dev)
    for s in range(0, B, 20000):
        e = min(s + 20000, B)
        dc = rm["d_chol"][s:e].double()
        dw += dc.T @ hidden[s:e].double()
        dh[s:e] = dc @ w22.double()
    errs = {"loss": abs(float(rf["loss"]) - float(rm["loss"])) / abs(float(rm["loss"])), "d_mu": rel(rf["d_mu"], rm["d_mu"]),
            "d_a": rel(rf["d_a"], rm["d_a"]), "d_w22": rel(rf["d_w22"], dw), "d_b22": rel(rf["d_b22"], rm["d_chol"].double().sum(0)),
            "d_hidden": rel(rf["d_hidden"], dh)}
    assert all(v < TOL for v in errs.values()), errs


def test_c2_norm_encoder_weight_gradients(ops):
    from vipsy_b200 import workloads
    from vipsy_b200.models import NormEncoder
    B, J, H = 1_000_000, 100, 64
    torch.manual_seed(1)
    w = workloads.uni_irt(B, J, "irt_4pl", missing=0.05)
    dev = w["y"].device
    enc = NormEncoder(J, 1, H).to(dev)
    mu, sigma, pre = ops.norm_encoder_forward(w["y"], enc)
    d_mu, d_sigma = torch.randn(B, device=dev), torch.randn(B, device=dev)
    g = ops.norm_encoder_backward(w["y"], enc, sigma, pre, d_mu, d_sigma, x_tiles=ops.norm_encoder_pack_tiles(w["y"]))
    x = torch.where(w["y"] == 255, torch.full((), -1.0, device=dev, dtype=torch.float64), w["y"].double())
    dhid = d_mu.double()[:, None] * enc.fc21.weight.double() + (d_sigma.double() * sigma.double())[:, None] * enc.fc22.weight.double()
    dpre = dhid * torch.sigmoid(pre.double())
    assert rel(g["fc1.weight"], dpre.T @ x) < TOL and rel(g["fc1.bias"], dpre.sum(0)) < TOL

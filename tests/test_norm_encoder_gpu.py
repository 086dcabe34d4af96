"""GPU: the NormEncoder kernels (csrc/enc_norm.cu -- vi.py:417-435 with the guide's NaN -> -1 input rule,
vi.py:679-682; tcgen05 variant csrc/enc_norm_tc.cu for J <= 127) against the fp64 oracle restatement: mu, sigma
forward; all six weight gradients backward; both kernel families.
Floating point: tolerance 2e-5 of the largest magnitude of each tensor (fp32 accumulation over J and B)."""
import pytest
import torch

from oracle import vipsy_oracle as o
from tests.util import codes_from_float

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


def _encoder(J, H, seed):
    from vipsy_b200.models import NormEncoder
    torch.manual_seed(seed)
    enc = NormEncoder(J, 1, H)
    with torch.no_grad():                      # move off the init so that softplus sees both tails
        for p in enc.parameters():
            p.mul_(3.0)
    return enc.cuda()


def _close(got, ref, name, tol=2e-5):
    ref = ref.to(torch.float64).cpu().reshape(-1)
    got = got.to(torch.float64).cpu().reshape(-1)
    assert got.shape == ref.shape, name
    scale = max(float(ref.abs().max()) if ref.numel() else 0.0, 1e-6)
    err = float((got - ref).abs().max()) if ref.numel() else 0.0
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("B,J,H,missing,sub", [
    (300, 7, 5, 0.2, None),          # the golden-vector shape class: tiny, H not a multiple of 4
    (1000, 100, 64, 0.1, None),      # c2 row shape
    (777, 101, 64, 0.1, None),       # unaligned rows
    (513, 130, 33, 0.3, None),       # two item chunks
    (400, 300, 16, 0.0, None),       # three item chunks, no missing cells
    (2000, 50, 64, 0.1, 333),        # minibatch through idx
    (1, 1, 1, 0.0, None),
    (40000, 64, 64, 0.05, None),     # more tiles than CTAs: persistent loop
    (5000, 127, 64, 0.1, None),      # widest tensor-core shape (item slot 127 carries the bias column)
    (3000, 16, 8, 0.1, None),
    (2500, 200, 64, 0.3, None),      # two 128-slot item blocks: tensor cores through the resident tiles only
    (1111, 255, 48, 0.1, None),      # widest resident shape (slot 255 = bias column)
    (900, 128, 64, 0.0, None),
])
@pytest.mark.parametrize("tc", [True, False])
def test_norm_encoder_forward_backward(ops, B, J, H, missing, sub, tc):
    g = torch.Generator().manual_seed(B * 31 + J)
    y = (torch.rand(B, J, generator=g) < 0.6).float()
    if missing:
        y[torch.rand(B, J, generator=g) < missing] = float("nan")
    enc = _encoder(J, H, B + J + H)
    codes = codes_from_float(y).cuda()
    idx = torch.randperm(B, generator=g)[:sub].cuda() if sub else None
    # tensor cores, full batch: once from the codes, once from the resident operand tiles
    xt = ops.norm_encoder_pack_tiles(codes) if (tc and idx is None and B > 1) else None
    mu, sigma, pre = ops.norm_encoder_forward(codes, enc, idx, tensor_cores=tc, x_tiles=xt)
    Bb = sub or B
    d_mu = torch.randn(Bb, generator=g).cuda()
    d_sigma = torch.randn(Bb, generator=g).cuda()
    got = ops.norm_encoder_backward(codes, enc, sigma, pre, d_mu, d_sigma, idx, grad_scale=0.5, tensor_cores=tc,
                                    x_tiles=xt)
    torch.cuda.synchronize()
    # oracle, fp64
    leaves = {n: p.detach().cpu().double().requires_grad_(True) for n, p in enc.named_parameters()}
    yb = y if idx is None else y[idx.cpu()]
    loc, sc = o.norm_encoder(yb, leaves["fc1.weight"], leaves["fc1.bias"], leaves["fc21.weight"], leaves["fc21.bias"],
                             leaves["fc22.weight"], leaves["fc22.bias"])
    torch.autograd.backward([loc, sc], [0.5 * d_mu.cpu().double().reshape(-1, 1), 0.5 * d_sigma.cpu().double().reshape(-1, 1)])
    _close(mu, loc.detach(), "mu")
    _close(sigma, sc.detach(), "sigma")
    for n, leaf in leaves.items():
        _close(got[n], leaf.grad, n)


def test_norm_encoder_resident_tiles_equal_on_the_fly(ops):
    """The resident operand tiles hold exactly what the kernels expand from the codes: identical results."""
    g = torch.Generator().manual_seed(11)
    y = (torch.rand(3001, 100, generator=g) < 0.6).float()
    y[torch.rand(3001, 100, generator=g) < 0.1] = float("nan")
    enc = _encoder(100, 64, 2)
    codes = codes_from_float(y).cuda()
    xt = ops.norm_encoder_pack_tiles(codes)
    d = torch.randn(3001, generator=g).cuda()
    out = []
    for tiles in (None, xt):
        mu, sigma, pre = ops.norm_encoder_forward(codes, enc, x_tiles=tiles)
        gr = ops.norm_encoder_backward(codes, enc, sigma, pre, d, d * 0.5, x_tiles=tiles)
        out.append((mu, sigma, pre, gr))
    torch.cuda.synchronize()
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])
    for n in out[0][3]:
        assert torch.equal(out[0][3][n], out[1][3][n]), n
    assert ops.norm_encoder_pack_tiles(torch.zeros((10, 256), dtype=torch.uint8, device="cuda")) is None
    assert ops.norm_encoder_pack_tiles(torch.zeros((300, 200), dtype=torch.uint8, device="cuda")).numel() == 3 * 65536


def test_norm_encoder_deterministic(ops):
    g = torch.Generator().manual_seed(5)
    y = (torch.rand(50000, 100, generator=g) < 0.6).float()
    enc = _encoder(100, 64, 1)
    codes = codes_from_float(y).cuda()
    mu, sigma, pre = ops.norm_encoder_forward(codes, enc)
    d = torch.randn(50000, generator=g).cuda()
    a = ops.norm_encoder_backward(codes, enc, sigma, pre, d, d * 0.3)
    b = ops.norm_encoder_backward(codes, enc, sigma, pre, d, d * 0.3)
    torch.cuda.synchronize()
    for n in a:
        assert torch.equal(a[n], b[n]), n


def test_norm_encoder_rejects_wide_hidden(ops):
    from vipsy_b200 import _native
    enc = _encoder(10, 65, 0)
    codes = torch.zeros((4, 10), dtype=torch.uint8, device="cuda")
    with pytest.raises(_native.VipsyError):
        ops.norm_encoder_forward(codes, enc)


def test_vaeirt_fused_encoder_matches_autograd_path(ops):
    """One SVI-level loss_and_grads of VaeIRT with the encoder kernels vs the torch-autograd encoder."""
    import vipsy_b200 as vb
    g = torch.Generator().manual_seed(9)
    y = (torch.rand(3000, 40, generator=g) < 0.6).float()
    y[torch.rand(3000, 40, generator=g) < 0.1] = float("nan")
    out = {}
    for fused in (True, False):
        vb.clear_param_store()
        torch.manual_seed(3)
        m = vb.VaeIRT(data=y, model="irt_3pl", hidden_dim=64, seed=7)
        m.fused_encoder = fused
        loss, sites = m.loss_and_grads(vb.Trace_ELBO(4))
        torch.cuda.synchronize()
        out[fused] = (float(loss), {n: p.grad.clone() for n, p, _ in sites})
    assert abs(out[True][0] - out[False][0]) <= 2e-6 * abs(out[False][0])
    for n, gr in out[False][1].items():
        _close(out[True][1][n], gr, n, tol=5e-5)


def test_vaechodina_fused_encoder_matches_autograd_path(ops):
    """VaeCHoDina (vi.py:961-981) holds the same NormEncoder for theta: kernels vs torch autograd, one loss_and_grads."""
    import vipsy_b200 as vb
    g = torch.Generator().manual_seed(21)
    K, J, N = 4, 30, 2000
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    q[0, q.sum(0) == 0] = 1.0
    y = (torch.rand(N, J, generator=g) < 0.55).float()
    y[torch.rand(N, J, generator=g) < 0.1] = float("nan")
    out = {}
    for fused in (True, False):
        vb.clear_param_store()
        torch.manual_seed(5)
        m = vb.VaeCHoDina(data=y, q=q, hidden_dim=32, seed=3)
        m.fused_encoder = fused
        loss, sites = m.loss_and_grads(vb.TraceEnum_ELBO(2))
        torch.cuda.synchronize()
        out[fused] = (float(loss), {n: p.grad.clone() for n, p, _ in sites})
    assert abs(out[True][0] - out[False][0]) <= 2e-6 * abs(out[False][0])
    assert set(out[True][1]) == set(out[False][1])
    for n, gr in out[False][1].items():
        _close(out[True][1][n], gr, n, tol=5e-5)

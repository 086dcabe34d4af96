"""Accuracy of the tensor-core accumulations at the BASELINE sizes (the tcgen05 accumulator adds with truncation,
so long chains of accumulating MMAs drift): item gradients of the K = 100 likelihood kernel and the fc22 gradients
of the fused guide at 200 000 persons, against fp64 torch on the device."""
import sys, torch
from vipsy_b200 import ops, build, workloads
build.build()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
K, J, H = 100, 500, 64
w = workloads.mirt(B, J, K)
def rel(x, y):
    return float((x.double() - y.double()).abs().max() / y.double().abs().max())
eps = ops.fill_normal(1, B, K, 3, 1)
r = ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, eps=eps, a=w["a"], b=w["b"],
                      model="irt_2pl", a_free=w["a_free"])
# fp64 reference of the item gradients from the kernel's own x (x is exact to 1e-7)
raw = w["chol_packed_raw"].double()
ti = torch.tril_indices(K, K, device=dev)
x = torch.empty(B, K, dtype=torch.float64, device=dev)
CH = 20000
da = torch.zeros(K, J, dtype=torch.float64, device=dev); db = torch.zeros(1, J, dtype=torch.float64, device=dev)
for s in range(0, B, CH):
    e = min(s + CH, B)
    L = torch.zeros(e - s, K, K, dtype=torch.float64, device=dev)
    L[:, ti[0], ti[1]] = raw[s:e]
    L = L.tril(-1) + torch.diag_embed(torch.diagonal(L, dim1=1, dim2=2).exp())
    xx = w["mu"][s:e].double() + torch.einsum("bij,bj->bi", L, eps[0, s:e].double())
    x[s:e] = xx
    a64, b64 = w["a"].double(), w["b"].double()
    p = torch.sigmoid(xx @ a64 + b64)
    y = w["y"][s:e]
    m = (y != 255).double()
    gl = (y.double().clamp(max=1) - p) * m          # d lb / d logit for the 2PL (unclamped cells)
    da -= xx.T @ gl
    db -= gl.sum(0, keepdim=True)
    del L
da *= w["a_free"].double()
print(f"B={B}: lik kernel  d_a {rel(r['d_a'], da):.2e}  d_b {rel(r['d_b'], db):.2e}", flush=True)
# fused guide: fc22 gradients
torch.manual_seed(0)
P = K * (K + 1) // 2
hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)); w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.02
b22 = torch.zeros(P, device=dev)
rows = torch.arange(K, device=dev); b22[rows * (rows + 1) // 2 + rows] = -0.7
rf = ops.irt_elbo_grad(w["y"], w["mu"], mvn_encoder=(hidden, w22, b22), eps=eps, a=w["a"], b=w["b"], model="irt_2pl",
                       a_free=w["a_free"])
packed = hidden @ w22.T + b22
rm = ops.irt_elbo_grad(w["y"], w["mu"], chol=packed, chol_packed=True, chol_raw=True, eps=eps, a=w["a"], b=w["b"],
                       model="irt_2pl", a_free=w["a_free"])
dw = torch.zeros(P, H, dtype=torch.float64, device=dev)
dh = torch.empty(B, H, dtype=torch.float64, device=dev)
for s in range(0, B, CH):
    e = min(s + CH, B)
    dc = rm["d_chol"][s:e].double()
    dw += dc.T @ hidden[s:e].double()
    dh[s:e] = dc @ w22.double()
db22 = rm["d_chol"].double().sum(0)
print(f"B={B}: fused guide loss {abs(float(rf['loss']) - float(rm['loss'])) / abs(float(rm['loss'])):.2e} d_mu {rel(rf['d_mu'], rm['d_mu']):.2e} "
      f"d_w22 {rel(rf['d_w22'], dw):.2e} d_b22 {rel(rf['d_b22'], db22):.2e} d_hidden {rel(rf['d_hidden'], dh):.2e} "
      f"d_a(fused vs fp64) {rel(rf['d_a'], da) if False else rel(rf['d_a'], rm['d_a']):.2e}", flush=True)
# NormEncoder backward at the c2 shape: d fc1.weight accumulates over every tile of a persistent CTA
from vipsy_b200.models import NormEncoder
Bn, Jn, Hn = 1_000_000, 100, 64
torch.manual_seed(1)
w2 = workloads.uni_irt(Bn, Jn, "irt_4pl", missing=0.05)
enc = NormEncoder(Jn, 1, Hn).to(dev)
mu, sigma, pre = ops.norm_encoder_forward(w2["y"], enc)
d_mu = torch.randn(Bn, device=dev); d_sigma = torch.randn(Bn, device=dev)
for tiles in (None, ops.norm_encoder_pack_tiles(w2["y"])):
    g = ops.norm_encoder_backward(w2["y"], enc, sigma, pre, d_mu, d_sigma, x_tiles=tiles)
    x = torch.where(w2["y"] == 255, torch.full((), -1.0, device=dev, dtype=torch.float64), w2["y"].double())
    dlog = d_sigma.double() * sigma.double()
    dhid = d_mu.double()[:, None] * enc.fc21.weight.double() + dlog[:, None] * enc.fc22.weight.double()
    dpre = dhid * torch.sigmoid(pre.double())
    dw1 = dpre.T @ x
    print(f"NormEncoder bwd 1M x 100 (tiles={'resident' if tiles is not None else 'on the fly'}): d fc1.weight {rel(g['fc1.weight'], dw1):.2e} "
          f"d fc1.bias {rel(g['fc1.bias'], dpre.sum(0)):.2e}", flush=True)

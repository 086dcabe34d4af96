"""Development check of the fused fc22 + sampler kernel against torch (run on the GPU box)."""
import sys, time
import torch
from vipsy_b200 import ops, build
build.build()
torch.manual_seed(0)
dev = "cuda"
def run(B, K, H, S=1, inject=False):
    P = K * (K + 1) // 2
    hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev))
    w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.25
    b22 = (torch.rand(P, device=dev) - 0.5) * 0.25
    mu = torch.randn(B, K, device=dev)
    eps = torch.randn(S, B, K, device=dev) if inject else None
    r = ops.mvn_encoder_sample(mu, hidden, w22, b22, eps=eps, n_particles=S, seed=7, offset=3)
    torch.cuda.synchronize()
    e = r["eps"].double()
    if inject:
        assert torch.equal(r["eps"], eps)
    else:
        ref = ops.fill_normal(S, B, K, 7, 3)
        assert torch.equal(r["eps"], ref), (r["eps"] - ref).abs().max()
    raw = hidden.double() @ w22.double().T + b22.double()
    L = torch.zeros(B, K, K, dtype=torch.float64, device=dev)
    ti = torch.tril_indices(K, K, device=dev)
    L[:, ti[0], ti[1]] = raw
    d = torch.diagonal(L, dim1=1, dim2=2)
    Lc = L.tril(-1) + torch.diag_embed(d.exp())
    x = mu.double().unsqueeze(0) + torch.einsum("bij,sbj->sbi", Lc, e)
    ex = (r["x"].double() - x).abs().max().item() / x.abs().max().item()
    ed = (r["diag"].double() - d.exp()).abs().max().item() / d.exp().abs().max().item()
    el = (r["logdet"].double() - d.sum(1)).abs().max().item() / d.sum(1).abs().max().item()
    print(f"B={B} K={K} H={H} S={S} inject={inject}: x {ex:.2e} diag {ed:.2e} logdet {el:.2e}", flush=True)
    assert ex < 2e-6 and ed < 2e-6 and el < 2e-6
for cfg in [(300, 10, 16), (128, 8, 64), (1000, 100, 64), (777, 37, 20), (260, 128, 64), (500, 3, 8), (64, 100, 64, 3)]:
    run(*cfg)
    run(*cfg, inject=True)
# timing at the c3 shape
B, K, H = 200_000, 100, 64
P = K * (K + 1) // 2
hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)); w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.25
b22 = torch.zeros(P, device=dev); mu = torch.randn(B, K, device=dev)
for _ in range(3):
    ops.mvn_encoder_sample(mu, hidden, w22, b22, seed=1, offset=1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(10):
    ops.mvn_encoder_sample(mu, hidden, w22, b22, seed=1, offset=i)
ev[1].record(); torch.cuda.synchronize()
print("c3 shape fused fc22+sample: %.3f ms per call (incl. prep + output allocs)" % (ev[0].elapsed_time(ev[1]) / 10))

"""Development check: fused amortised-MVN-guide path == the same evaluation with the factor materialised by torch."""
import torch
from vipsy_b200 import ops, build, workloads
build.build()
dev = "cuda"
torch.manual_seed(0)

def run(B, J, K, H, S=1, inject=True, model="irt_2pl"):
    P = K * (K + 1) // 2
    y = (torch.rand(B, J, device=dev) < 0.6).float()
    y[torch.rand(B, J, device=dev) < 0.1] = float("nan")
    codes = ops.pack_responses(y)
    hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)).requires_grad_(True)
    w22 = ((torch.rand(P, H, device=dev) - 0.5) * 0.2).requires_grad_(True)
    b22 = ((torch.rand(P, device=dev) - 0.5) * 0.2).requires_grad_(True)
    mu = (torch.randn(B, K, device=dev) * 0.5)
    a = torch.rand(K, J, device=dev) * 0.5
    b = torch.randn(1, J, device=dev)
    eps = torch.randn(S, B, K, device=dev) if inject else ops.fill_normal(S, B, K, 5, 2)
    kw = dict(a=a, b=b, model=model, scale=1.7, seed=5, offset=2)
    if model != "irt_2pl":
        kw["c"] = torch.rand(1, J, device=dev) * 0.2
    r = ops.irt_elbo_grad(codes, mu, mvn_encoder=(hidden, w22, b22), eps=eps if inject else None, n_particles=S, **kw)
    hd, wd, bd = (t.detach().double().requires_grad_(True) for t in (hidden, w22, b22))
    raw = hd @ wd.T + bd                                     # fp64 arbiter for the encoder side
    ref = ops.irt_elbo_grad(codes, mu, chol=raw.detach().float(), chol_packed=True, chol_raw=True, eps=eps, **kw)
    torch.autograd.backward([raw], [ref["d_chol"].double()])
    gh, gw, gb = hd.grad, wd.grad, bd.grad
    torch.cuda.synchronize()
    def rel(x, y):
        return float((x.double() - y.double()).abs().max() / y.double().abs().max())
    errs = {"loss": abs(float(r["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])),
            "d_mu": rel(r["d_mu"], ref["d_mu"]), "d_a": rel(r["d_a"], ref["d_a"]), "d_b": rel(r["d_b"], ref["d_b"]),
            "d_hidden": rel(r["d_hidden"], gh), "d_w22": rel(r["d_w22"], gw), "d_b22": rel(r["d_b22"], gb)}
    print(f"B={B} J={J} K={K} H={H} S={S} inject={inject} {model}:", {k: f"{v:.1e}" for k, v in errs.items()}, flush=True)
    assert all(v < 2e-5 for v in errs.values()), errs

for cfg in [(300, 40, 10, 16), (128, 20, 8, 64), (1000, 100, 100, 64), (777, 33, 37, 20), (500, 12, 3, 8), (260, 60, 128, 64)]:
    run(*cfg)
run(300, 40, 10, 16, inject=False)
run(200, 40, 12, 32, S=3)
run(300, 40, 10, 16, model="irt_3pl")
print("full path OK")

import torch
from vipsy_b200 import ops, build, workloads
build.build()
dev = "cuda"
B, J, K, H = 200_000, 500, 100, 64
w = workloads.mirt(B, J, K)
P = K * (K + 1) // 2
torch.manual_seed(0)
hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)); w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.02
b22 = torch.zeros(P, device=dev)
for i in range(2):
    ops.irt_elbo_grad(w["y"], w["mu"], mvn_encoder=(hidden, w22, b22), a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"], seed=1, offset=i)
torch.cuda.synchronize()

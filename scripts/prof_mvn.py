"""ncu target: the fused fc22 + sampler kernel (and later its backward) at the c3 shape."""
import torch
from vipsy_b200 import ops, build
build.build()
dev = "cuda"
B, K, H = 200_000, 100, 64
P = K * (K + 1) // 2
torch.manual_seed(0)
hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)); w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.25
b22 = torch.zeros(P, device=dev); mu = torch.randn(B, K, device=dev)
for i in range(2):
    ops.mvn_encoder_sample(mu, hidden, w22, b22, seed=1, offset=i)
torch.cuda.synchronize()

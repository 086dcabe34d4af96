"""Three launches of the c2 workload for ncu (kernel regex irt_uni)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
N, J, S = 1_000_000, 100, 32
model = sys.argv[1] if len(sys.argv) > 1 else "irt_4pl"
g = torch.Generator(device="cuda").manual_seed(1)
y = (torch.rand(N, J, device="cuda", generator=g) < 0.6).to(torch.uint8)
mu = torch.randn(N, 1, device="cuda", generator=g) * 0.5
sg = torch.exp(torch.randn(N, 1, device="cuda", generator=g) * 0.1)
a = torch.rand(1, J, device="cuda") * 2.5 + 0.5
b = torch.randn(1, J, device="cuda")
c = torch.rand(1, J, device="cuda") * 0.15 + 0.05
d = torch.rand(1, J, device="cuda") * 0.15 + 0.8
for i in range(3):
    ops.irt_elbo_grad(y, mu, sigma=sg, n_particles=S, a=a, b=b, c=c, d=d, model=model, seed=1, offset=i)
torch.cuda.synchronize()

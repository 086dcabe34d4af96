"""Two launches each of the S == 1 paths on the c4 shape for ncu: dense row kernel, CSR kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
N, J = 2_000_000, 200
g = torch.Generator(device="cuda").manual_seed(1)
y = (torch.rand(N, J, device="cuda", generator=g) < 0.6).to(torch.uint8)
y[torch.rand(N, J, device="cuda", generator=g) < 0.9] = 255
mu = torch.randn(N, 1, device="cuda", generator=g) * 0.5
sg = torch.exp(torch.randn(N, 1, device="cuda", generator=g) * 0.1)
a = torch.rand(1, J, device="cuda") * 2.5 + 0.5
b = torch.randn(1, J, device="cuda")
obs = ops.observed_csr(y)
for i in range(2):
    ops.irt_elbo_grad(y, mu, sigma=sg, n_particles=1, a=a, b=b, model="irt_2pl", seed=1, offset=i)
    ops.irt_elbo_grad(y, mu, sigma=sg, n_particles=1, a=a, b=b, model="irt_2pl", seed=1, offset=i, obs=obs)
torch.cuda.synchronize()

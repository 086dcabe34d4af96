"""ncu target: eager VaeIRT iterations at the c3 shape with the fused guide (launch list of one whole iteration)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer
N = int(os.environ.get("N", 200_000))
y = (torch.rand(N, 500, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0)) < 0.6).float()
vb.clear_param_store()
torch.manual_seed(0)
m = vb.VaeIRT(data=y, model="irt_2pl", x_feature=100)
svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-3}), loss=vb.Trace_ELBO(1), cuda_graph=False)
for _ in range(3):
    svi.step(m.data)
torch.cuda.synchronize()
torch.cuda.profiler.start()
svi.step(m.data)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

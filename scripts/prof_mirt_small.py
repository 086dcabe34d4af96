import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
K, J, S, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 500_000
w = workloads.mirt(N, J, K)
dchol = torch.zeros_like(w["chol_packed_raw"]); dmu = torch.empty_like(w["mu"])
for i in range(2):
    ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, n_particles=S,
                      a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"], seed=1, offset=i,
                      d_mu_out=dmu, d_scale_out=dchol, n_masked=0)
torch.cuda.synchronize()

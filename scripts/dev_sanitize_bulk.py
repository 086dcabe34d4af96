"""Small K > 32 evaluations through the bulk-fed sampler for compute-sanitizer (partial last groups included)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
for (N, K, J, S) in ((1501, 100, 40, 1), (1203, 50, 40, 1), (900, 64, 40, 2), (1303, 33, 40, 1), (777, 128, 40, 1)):
    w = workloads.mirt(N, J, K)
    for inject in (False, True):
        eps = ops.fill_normal(S, N, K, 3, 1) if inject else None
        r = ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, eps=eps,
                              n_particles=S, a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"], seed=1, offset=3)
        torch.cuda.synchronize()
        print(N, K, S, inject, float(r["loss"]), flush=True)

"""Fused small-K kernel vs the three-kernel path, K = 2..8 (500k persons, J = 100, S = 8 and S = 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
from scripts.dev_bench_uni import timeit
n, J = 500_000, 100
for K in (2, 3, 4, 5, 6, 7, 8):
    w = workloads.mirt(n, J, K)
    dchol = torch.zeros_like(w["chol_packed_raw"]); dmu = torch.empty_like(w["mu"])
    for S in (1, 8):
        res = []
        for fused in (True, False):
            off = [0]
            def run():
                off[0] += 1
                return ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True,
                                         n_particles=S, a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"], seed=1,
                                         offset=off[0], d_mu_out=dmu, d_scale_out=dchol, n_masked=0, fused=fused)
            res.append(timeit(run, it=10))
        print(f"K={K} J={J} S={S}: fused {res[0]:.3f} ms  three-kernel {res[1]:.3f} ms  ({n*J*S/res[0]*1e3:.2e} evals/s)", flush=True)

"""c3 shape at S = 1, 2, 8, 32 (per-kernel times via CUDA events around the whole evaluation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
from scripts.dev_bench_uni import timeit
N = 200_000
w = workloads.mirt(N, 500, 100)
dchol = torch.zeros_like(w["chol_packed_raw"]); dmu = torch.empty_like(w["mu"])
for S in (1, 2, 8, 32):
    off = [0]
    def run():
        off[0] += 1
        return ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True,
                                 n_particles=S, a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"],
                                 seed=1, offset=off[0], d_mu_out=dmu, d_scale_out=dchol)
    ms = timeit(run, it=5)
    fl = N * S * (6 * 100 * 500 + 2 * 100 * 101)
    print(f"c3 S={S}: {ms:.3f} ms  {N*500*S/ms*1e3:.3e} evals/s  {fl/ms/1e9:.1f} TFLOP/s (fp32-equivalent useful flops)", flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
from vipsy_b200.clocks import ClockSampler
from scripts.dev_bench_uni import timeit

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    cfgs = ((100, 500, 1),) if os.environ.get('VIPSY_C3_ONLY') else ((100, 500, 1), (10, 100, 8), (2, 50, 8))
    for (K, J, S) in cfgs:
        n = N if K == 100 else 500_000
        w = workloads.mirt(n, J, K)
        off = [0]
        dchol = torch.zeros_like(w["chol_packed_raw"])
        dmu = torch.empty_like(w["mu"])
        n_miss = int((w["y"] == 255).sum().item())
        def run():
            off[0] += 1
            return ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True,
                                     n_particles=S, a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"],
                                     seed=1, offset=off[0], d_mu_out=dmu, d_scale_out=dchol, n_masked=n_miss)
        with ClockSampler() as cs:
            ms = timeit(run, it=10)
        tri = K * (K + 1) // 2
        gb = n * (J + 8 * K + 8 * tri) / 1e9
        print(f"mirt K={K} J={J} S={S} N={n}: {ms:.3f} ms  {n*J*S/ms*1e3:.3e} evals/s  "
              f"algo {gb:.2f} GB -> {gb/ms*1e3:.0f} GB/s  clk={cs.summary()['sm_mhz']}", flush=True)

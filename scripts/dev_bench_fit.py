"""Whole-iteration timing (SVI.step: ELBO + grads + free mask + Adam + zero grads) eager vs CUDA graph."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer


def run(name, make, loss, iters=2000):
    for graph in (False, True):
        vb.clear_param_store()
        torch.manual_seed(0)
        m = make()
        svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-2}), loss=loss, cuda_graph=graph)
        for _ in range(5):
            svi.step(m.data)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            svi.step_async(m.data)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        print(f"{name}: {'graph' if graph else 'eager'} {dt*1e6:.1f} us/iteration", flush=True)


if __name__ == "__main__":
    g = torch.Generator().manual_seed(0)
    y5 = (torch.rand(1000, 5, generator=g) < 0.7).float()
    y100 = (torch.rand(100_000, 100, generator=g) < 0.6).float()
    K, J = 5, 100
    q = (torch.rand(K, J, generator=g) < 0.5).float(); q[0, q.sum(0) == 0] = 1
    yq = (torch.rand(1500, J, generator=g) < 0.5).float()
    run("c1 VIRT 2PL 1000x5 full batch S=1", lambda: vb.VIRT(data=y5, model="irt_2pl"), vb.Trace_ELBO(1))
    run("VIRT 2PL 100k x 100, subsample 100 (test.py:308)", lambda: vb.VIRT(data=y100, model="irt_2pl", subsample_size=100), vb.Trace_ELBO(1))
    run("VaeIRT 2PL 100k x 100, subsample 100 (test.py:288)", lambda: vb.VaeIRT(data=y100, model="irt_2pl", subsample_size=100), vb.Trace_ELBO(1))
    run("VaeIRT 2PL K=3 100k x 100, subsample 100 (test.py:343)", lambda: vb.VaeIRT(data=y100, model="irt_2pl", x_feature=3, subsample_size=100), vb.Trace_ELBO(1))
    run("VCCDM dina 1500 x 100 K=5 full (test.py:558)", lambda: vb.VCCDM(data=yq, q=q), vb.TraceEnum_ELBO(1), iters=500)
    run("VCHoDina 1500 x 100 K=5 full", lambda: vb.VCHoDina(data=yq, q=q), vb.TraceEnum_ELBO(1), iters=500)

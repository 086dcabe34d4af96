"""Full VaeIRT iteration at the c2 shape (1M x 100, 4PL, S=32, full batch): encoder + ELBO kernel + Adam."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer

N = int(os.environ.get("N", 1_000_000))
g = torch.Generator().manual_seed(0)
y = (torch.rand(N, 100, generator=g) < 0.6).float()
for model, S in (("irt_4pl", 32), ("irt_2pl", 1)):
    for graph in (False, True):
        vb.clear_param_store()
        torch.manual_seed(0)
        m = vb.VaeIRT(data=y, model=model)
        svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-3}), loss=vb.Trace_ELBO(S), cuda_graph=graph)
        for _ in range(4):
            svi.step(m.data)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            svi.step_async(m.data)
        torch.cuda.synchronize()
        print(f"VaeIRT {model} {N}x100 S={S} full batch: {'graph' if graph else 'eager'} "
              f"{(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/iteration", flush=True)
        del svi, m

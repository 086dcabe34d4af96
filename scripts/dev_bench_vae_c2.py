"""Full VaeIRT iteration at the c2 shape (1M x 100, 4PL, S=32, full batch): encoder + ELBO kernel + Adam.
N=2000000 J=200 MISSING=0.9: the c4 shape (2PL, S=1, observed-cell CSR kernel)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer

N, J, MISSING = int(os.environ.get("N", 1_000_000)), int(os.environ.get("J", 100)), float(os.environ.get("MISSING", 0))
g = torch.Generator(device="cuda").manual_seed(0)
y = (torch.rand(N, J, generator=g, device="cuda") < 0.6).float()
if MISSING:
    y[torch.rand(N, J, generator=g, device="cuda") < MISSING] = float("nan")
CASES = (("irt_2pl", 1),) if MISSING else (("irt_4pl", 32), ("irt_2pl", 1))
for model, S in CASES:
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
        print(f"VaeIRT {model} {N}x{J} missing={MISSING} S={S} full batch: {'graph' if graph else 'eager'} "
              f"{(time.perf_counter() - t0) / 20 * 1e3:.3f} ms/iteration", flush=True)
        del svi, m

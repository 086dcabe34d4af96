"""e2e (host buffers in, host results out) time of the c2 step vs the number of pipeline chunks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
from vipsy_b200.host_api import IrtHostStep
N, J, S = 1_000_000, 100, 32
w = workloads.uni_irt(N, J, "irt_4pl", seed=1234)
hy = ops.pack_responses_2bit_host(w["y"].cpu(), pin=True)
hmu = torch.empty(w["mu"].shape, dtype=torch.float32, pin_memory=True).copy_(w["mu"])
hsg = torch.empty(w["sigma"].shape, dtype=torch.float32, pin_memory=True).copy_(w["sigma"])
hp = {k: w[k].cpu() for k in ("a", "b", "c", "d")}
for nc in (1, 2, 3, 4, 6, 8, 12, 16):
    host = IrtHostStep(N, J, "irt_4pl", n_chunks=nc)
    for i in range(3):
        host(hy, hmu, hsg, hp["a"], hp["b"], hp["c"], hp["d"], n_particles=S, seed=1, offset=i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        host(hy, hmu, hsg, hp["a"], hp["b"], hp["c"], hp["d"], n_particles=S, seed=1, offset=10 + i)
    torch.cuda.synchronize()
    print(f"n_chunks={nc}: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms", flush=True)

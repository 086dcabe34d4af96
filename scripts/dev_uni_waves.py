"""How much of the c2 step is the partial last round of the persistent particle kernel?  Times whole numbers of
rounds (148 SMs x 16 warps x 32 persons = 75 776 persons) against the 1M-person workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
wave = 148 * 16 * 32
import os
SIZES = (15_360, wave, 3 * wave, 1_000_000) if os.environ.get('VIPSY_SIZES') == 'chunks' else (13 * wave, 1_000_000, 14 * wave, 2_000_000, 1_000_000)
for N in SIZES:
    w = workloads.uni_irt(N, 100, "irt_4pl", seed=1234, device=dev)
    off = [0]

    def run():
        off[0] += 1
        return ops.irt_elbo_grad(w["y"], w["mu"], sigma=w["sigma"], n_particles=32, a=w["a"], b=w["b"], c=w["c"],
                                 d=w["d"], model="irt_4pl", seed=1, offset=off[0])
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = sum(ms) / len(ms)
    print(f"N={N} ({N / wave:.3f} rounds): {t:.4f} ms, {t / (N / wave) * 1e3:.2f} us per round-equivalent", flush=True)

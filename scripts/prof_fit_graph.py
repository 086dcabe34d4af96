"""Kernel list of one captured VIRT / VaeIRT minibatch iteration (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer
which = sys.argv[1] if len(sys.argv) > 1 else "virt"
g = torch.Generator().manual_seed(0)
y = (torch.rand(100_000, 100, generator=g) < 0.6).float()
m = (vb.VIRT if which == "virt" else vb.VaeIRT)(data=y, model="irt_2pl", subsample_size=100)
svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(1), cuda_graph=False)
m._device_rng = True                      # the stream-ordered form that SVI(cuda_graph=True) captures
for _ in range(3):
    svi._step_device(m.data)
torch.cuda.synchronize()
torch.cuda.profiler.start()
svi._step_device(m.data)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

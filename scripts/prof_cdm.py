"""Two launches each of the c5 kernels (DINA, HO-DINA) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
for hod in (False, True):
    w = workloads.cdm(N, 60, 10, hodina=hod)
    qm = ops.qmask_from_q(w["q"])
    kw = dict(lam0=w["lam0"], lam1=w["lam1"], mu=w["mu"], sigma=w["sigma"], n_particles=1) if hod else {}
    for i in range(2):
        ops.cdm_enum_elbo_grad(w["y"], qm, w["g"], w["s"], n_attrs=10, seed=1, offset=i, **kw)
torch.cuda.synchronize()

"""Race detector for the c3 evaluation: every kernel on the path is deterministic by construction, so repeated
evaluations of the same inputs must agree bit for bit.  usage: python scripts/dev_race_c3.py [n_evals]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B, J, K = 200_000, 500, 100
w = workloads.mirt(B, J, K)
eps = ops.fill_normal(1, B, K, 3, 1)
junk = torch.empty(300 << 20, dtype=torch.uint8, device="cuda")
for inject in (True, False):
    ref = None
    bad = 0
    for i in range(n):
        junk.random_(0, 255)               # different L2 / timing state between evaluations
        r = ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True,
                              eps=eps if inject else None, n_particles=1, seed=5, offset=9,
                              a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"])
        torch.cuda.synchronize()
        print(f"  eval {i} done", flush=True)
        cur = {k: r[k].clone() for k in ("loss", "d_mu", "d_chol", "d_a", "d_b")}
        if ref is None:
            ref = cur
            continue
        diff = {k: int((ref[k] != cur[k]).sum()) for k in cur}
        if any(diff.values()):
            bad += 1
            print(f"inject={inject} eval {i}: differing elements {diff}", flush=True)
    print(f"NO_BULK={os.environ.get('VIPSY_NO_BULK', '0')} inject={inject}: {bad} of {n - 1} evaluations differ from the first", flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops, workloads
from vipsy_b200.clocks import ClockSampler
from scripts.dev_bench_uni import timeit

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    for hod in (False, True):
        w = workloads.cdm(N, 60, 10, hodina=hod)
        qm = ops.qmask_from_q(w["q"])
        kw = dict(lam0=w["lam0"], lam1=w["lam1"], mu=w["mu"], sigma=w["sigma"], n_particles=1) if hod else {}
        off = [0]
        def run():
            off[0] += 1
            return ops.cdm_enum_elbo_grad(w["y"], qm, w["g"], w["s"], n_attrs=10, seed=1, offset=off[0], **kw)
        with ClockSampler() as cs:
            ms = timeit(run, it=20)
        print(f"cdm hodina={hod} N={N} J=60 K=10: {ms:.3f} ms  {N*60*1024/ms*1e3:.3e} person.item.pattern/s  "
              f"{N/ms*1e3:.3e} persons/s clk={cs.summary()['sm_mhz']}", flush=True)
        print("loss", float(run()["loss"]))

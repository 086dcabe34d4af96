import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_umma_selftest import run, expect, CASES
for variant in (0, 1, 2):
    for swap in (0,):
        if variant == 1 and swap: continue
        g = torch.Generator(device="cuda").manual_seed(variant)
        s1, s2 = CASES[variant]
        t1 = torch.randn(s1, device="cuda", generator=g); t2 = torch.randn(s2, device="cuda", generator=g) * 1.5
        ref = expect(variant, t1, t2)
        for ns in (1, 3):
            got = run(variant, t1, t2, swap=swap, n_split=ns).double()
            rows = ref.shape[0]
            err = float((got[:rows] - ref).abs().max() / ref.abs().max())
            print(f"variant {variant} swap {swap} n_split {ns}: rel err {err:.3e}", flush=True)

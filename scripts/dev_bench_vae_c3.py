"""Full VaeIRT iteration at the c3 shape (200k x 500, 2PL, K = 100, per-person full covariance, S = 1):
MvnEncoder (torch: fc1, fc21, fc22 64 -> 5050) + MIRT kernels + Adam.  Shows what SURVEY 8f-2 (fc22 fusion) is worth."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vipsy_b200 as vb
from vipsy_b200 import infer

N = int(os.environ.get("N", 200_000))
y = (torch.rand(N, 500, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0)) < 0.6).float()
for fused, graph in ((False, False), (False, True), (True, False), (True, True)):
    vb.clear_param_store()
    torch.manual_seed(0)
    m = vb.VaeIRT(data=y, model="irt_2pl", x_feature=100)
    m.fused_encoder = fused
    svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-3}), loss=vb.Trace_ELBO(1), cuda_graph=graph)
    for _ in range(3):
        svi.step(m.data)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        svi.step_async(m.data)
    torch.cuda.synchronize()
    print(f"VaeIRT irt_2pl K=100 {N}x500 S=1 full batch, fc22 {'fused into the guide kernels' if fused else 'in torch'}: {'graph' if graph else 'eager'} "
          f"{(time.perf_counter() - t0) / 10 * 1e3:.3f} ms/iteration", flush=True)
    if graph is False and not fused:
        # encoder alone (forward + backward through fc22), same tensors
        enc = m.encoder
        x = m._encoder_input(None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            loc, tril = enc(x)
            torch.autograd.backward([loc, tril], [torch.ones_like(loc), torch.ones_like(tril)])
        torch.cuda.synchronize()
        print(f"  MvnEncoder forward + backward alone (torch): {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms", flush=True)
    del svi, m

"""Kernel-only timing of the fused amortised-MVN-guide evaluation at the c3 shape (CUDA events, L2 flushed)."""
import torch
from vipsy_b200 import ops, build, workloads
build.build()
dev = "cuda"
B, J, K, H = 200_000, 500, 100, 64
w = workloads.mirt(B, J, K)
P = K * (K + 1) // 2
torch.manual_seed(0)
hidden = torch.nn.functional.softplus(torch.randn(B, H, device=dev)); w22 = (torch.rand(P, H, device=dev) - 0.5) * 0.02
b22 = torch.zeros(P, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=10):
    for _ in range(3): fn()
    tot = 0.0
    for i in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
kw = dict(a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"], seed=1, offset=2)
t_f = timed(lambda: ops.irt_elbo_grad(w["y"], w["mu"], mvn_encoder=(hidden, w22, b22), **kw))
t_v = timed(lambda: ops.irt_elbo_grad(w["y"], w["mu"], mvn_encoder=(hidden, w22, b22), want_grad=False, **kw))
t_m = timed(lambda: ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True, **kw))
t_s = timed(lambda: ops.mvn_encoder_sample(w["mu"], hidden, w22, b22, seed=1, offset=2))
print(f"c3 200k x 500 x K=100, S=1: fused guide loss+grads {t_f:.3f} ms | value only {t_v:.3f} ms | sample site alone {t_s:.3f} ms | "
      f"three-kernel path on a materialised factor {t_m:.3f} ms (+ fc22 fwd/bwd in torch)")

"""NormEncoder kernels at minibatch size (B = 100 of 100k x 100, the reference's own usage): launch-latency regime."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
from vipsy_b200.models import NormEncoder
g = torch.Generator().manual_seed(0)
codes = (torch.rand(100_000, 100, generator=g) < 0.6).to(torch.uint8).cuda()
torch.manual_seed(0)
enc = NormEncoder(100, 1, 64).cuda()
idx = torch.randperm(100_000, generator=g)[:100].cuda()
d = torch.randn(100, generator=g).cuda()
mu, sigma, pre = ops.norm_encoder_forward(codes, enc, idx)
for name, fn in (("forward", lambda: ops.norm_encoder_forward(codes, enc, idx)),
                 ("backward (+ finalize)", lambda: ops.norm_encoder_backward(codes, enc, sigma, pre, d, d, idx))):
    gr = torch.cuda.CUDAGraph()
    fn(); torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for _ in range(20):
            fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr.replay(); e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"norm_encoder {name}, B=100 of 100k x 100: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (graph replay)")

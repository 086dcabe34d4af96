"""NormEncoder kernels at the c2 shape (1M x 100, H = 64): CUDA-event timing, or one launch each under ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
from vipsy_b200.models import NormEncoder

N, J, H = int(os.environ.get("N", 1_000_000)), int(os.environ.get("J", 100)), int(os.environ.get("H", 64))
g = torch.Generator().manual_seed(0)
codes = (torch.rand(N, J, generator=g) < 0.6).to(torch.uint8).cuda()
torch.manual_seed(0)
enc = NormEncoder(J, 1, H).cuda()
d = torch.randn(N, generator=g).cuda()
xt = ops.norm_encoder_pack_tiles(codes) if os.environ.get("RESIDENT", "1") == "1" else None
mu, sigma, pre = ops.norm_encoder_forward(codes, enc, x_tiles=xt)
ops.norm_encoder_backward(codes, enc, sigma, pre, d, d, x_tiles=xt)
torch.cuda.synchronize()
if os.environ.get("TIME", "1") == "1":
    for name, fn in (("forward", lambda: ops.norm_encoder_forward(codes, enc, x_tiles=xt)),
                     ("backward", lambda: ops.norm_encoder_backward(codes, enc, sigma, pre, d, d, x_tiles=xt))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"norm_encoder {name} {N}x{J} H={H} resident={xt is not None}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")

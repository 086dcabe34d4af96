import torch
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
a.normal_()
ms = t(lambda: a.fill_(1.0)); print(f"fill 4 GiB: {ms:.3f} ms {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: a.sum()); print(f"sum 4 GiB: {ms:.3f} ms {4.295/ms*1e3:.0f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"copy 4+4 GiB: {ms:.3f} ms {8.59/ms*1e3:.0f} GB/s")
ms = t(lambda: torch.mul(a, 2.0, out=b)); print(f"mul 4+4 GiB: {ms:.3f} ms {8.59/ms*1e3:.0f} GB/s")

"""torchrun --nproc-per-node N scripts/test_peer_allreduce.py : vipsy_peer_allreduce vs NCCL all_reduce
(exactness of the rank-ordered sum, agreement across ranks, latency)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from vipsy_b200 import build
from vipsy_b200.dist import PeerAllReduce

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
build.build()
n = 6994
par = PeerAllReduce(n_float=n, n_double=1, device=dev)
ok = True
for it in range(20):
    g, l = par.slot_views()
    torch.manual_seed(1000 * it + rank)
    x = torch.randn(n, device=dev)
    y = torch.randn(1, device=dev, dtype=torch.float64)
    g.copy_(x); l.copy_(y)
    out_f, out_d = par.reduce()
    # reference: gather everything, sum in rank order on every rank
    xs = [torch.empty_like(x) for _ in range(world)]; ys = [torch.empty_like(y) for _ in range(world)]
    dist.all_gather(xs, x); dist.all_gather(ys, y)
    rf = torch.zeros_like(x); rd = torch.zeros_like(y)
    for r in range(world):
        rf = rf + xs[r]; rd = rd + ys[r]
    ok &= bool(torch.equal(out_f, rf)) and bool(torch.equal(out_d, rd))
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
torch.cuda.synchronize()

def bench(fn, iters=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

buf = torch.randn(n, device=dev); lb = torch.randn(1, device=dev, dtype=torch.float64)
t_nccl = bench(lambda: (dist.all_reduce(buf), dist.all_reduce(lb)))
t_peer = bench(lambda: par.reduce())
par2 = PeerAllReduce(n_float=400, n_double=1, device=dev)
buf2 = torch.randn(400, device=dev)
t_nccl2 = bench(lambda: (dist.all_reduce(buf2), dist.all_reduce(lb)))
t_peer2 = bench(lambda: par2.reduce())
if rank == 0:
    print(f"world={world}: bit-exact and identical on all ranks: {bool(flag.item())}; n_float={n}: NCCL (2 all_reduce) "
          f"{t_nccl:.1f} us, peer kernel {t_peer:.1f} us; n_float=400: NCCL {t_nccl2:.1f} us, peer kernel {t_peer2:.1f} us",
          flush=True)
dist.destroy_process_group()

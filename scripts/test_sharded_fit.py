"""torchrun --nproc-per-node N scripts/test_sharded_fit.py : a person-sharded VIRT fit (each rank owns a contiguous
block of persons) with the gradient exchange through NCCL and through the NVLink peer-memory kernel must produce
the same item parameters on every rank and in both modes, and match the unsharded fit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vipsy_b200 as vb
from vipsy_b200 import build, infer
from vipsy_b200.dist import attach, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
build.build()
g = torch.Generator().manual_seed(5)
N, J = 4000, 24
y = (torch.rand(N, J, generator=g) < 0.6).float()
y[torch.rand(N, J, generator=g) < 0.1] = float("nan")
lo, hi = shard_bounds(N, world, rank)


def run(peer, sharded=True, steps=6):
    vb.clear_param_store()
    data = y[lo:hi] if sharded else y
    m = vb.VIRT(data=data, model="irt_3pl", seed=3, device=dev)
    if sharded:
        attach(m, N, lo, peer=peer)
    svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 2e-2}), loss=vb.Trace_ELBO(num_particles=2))
    losses = [svi.step(m.data) for _ in range(steps)]
    return losses, vb.param("a").clone(), vb.param("b").clone(), vb.param("c").clone()


ref = run(False, sharded=False)
nccl = run(False)
peer = run(True)
ok = True
for name, r in (("nccl", nccl), ("peer", peer)):
    ok &= all(abs(a - b) <= 2e-5 * abs(b) for a, b in zip(r[0], ref[0]))
    for t, tr in zip(r[1:], ref[1:]):
        ok &= bool(torch.allclose(t, tr, rtol=2e-4, atol=2e-5))
    # identical on every rank
    for t in r[1:]:
        t0 = t.clone()
        dist.broadcast(t0, 0)
        ok &= bool(torch.equal(t0, t))


# the amortised K > 1 guide (fused fc21 / fc22, csrc/enc_mvn.cu): encoder weights are replicated, their gradients -- among
# them d fc21 / d fc22 from the fused call -- are summed across the shards together with the item gradients
def run_vae(sharded, steps=5):
    vb.clear_param_store()
    torch.manual_seed(11)
    data = y[lo:hi] if sharded else y
    m = vb.VaeIRT(data=data, model="irt_2pl", x_feature=4, hidden_dim=16, seed=7, device=dev)
    if sharded:
        attach(m, N, lo, peer=False)
    svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(num_particles=1), cuda_graph=False)
    losses = [svi.step(m.data) for _ in range(steps)]
    return losses, vb.param("a").clone(), {n: p.detach().clone() for n, p in m.encoder.named_parameters()}


vref, vsh = run_vae(False), run_vae(True)
ok_vae = all(abs(a - b) <= 5e-5 * abs(b) for a, b in zip(vsh[0], vref[0])) and bool(torch.allclose(vsh[1], vref[1], rtol=5e-4, atol=5e-5))
for n in vref[2]:
    ok_vae &= bool(torch.allclose(vsh[2][n], vref[2][n], rtol=5e-4, atol=5e-5))
ok &= ok_vae
if rank == 0:
    print(f"world={world}: sharded VaeIRT K=4 (fused MVN guide) == unsharded: {ok_vae}; losses {vref[0][0]:.3f} -> {vref[0][-1]:.3f} "
          f"vs {vsh[0][0]:.3f} -> {vsh[0][-1]:.3f}", flush=True)
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world}: sharded fit == unsharded fit (NCCL and peer-memory exchange), identical on all ranks: "
          f"{bool(flag.item())}; losses {ref[0][0]:.4f} -> {ref[0][-1]:.4f}", flush=True)
dist.destroy_process_group()

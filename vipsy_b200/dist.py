"""Person-sharded data parallelism: one process per GPU, persons split into
contiguous blocks, per-person variational parameters and their optimiser state stay
local; the loss and the gradients of every replicated parameter (item parameters,
shared Cholesky factor, HO-DINA lambdas, encoder weights) are summed with ONE
all-reduce of a packed fp32 buffer per step -- NCCL over NVLink on the GPU box,
gloo in the CPU tests.  There is no other exchange on the data path.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous person block of `rank`: sizes differ by at most one, cover [0, n)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradReducer:
    """Callable(loss, leaves) -> loss: all-reduce(sum) of [loss_hi, loss_lo, grads...]."""

    def __init__(self, group=None):
        self.group = group
        self.bytes_last = 0

    def __call__(self, loss: torch.Tensor, leaves: Sequence[torch.Tensor]) -> torch.Tensor:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return loss
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in leaves]
        dev = loss.device
        hi = loss.detach().to(torch.float32)
        lo = (loss.detach().to(torch.float64) - hi.to(torch.float64)).to(torch.float32)
        flat = torch.cat([hi.reshape(1), lo.reshape(1)] + [g.reshape(-1).to(torch.float32) for g in grads]).to(dev)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        self.bytes_last = flat.numel() * 4
        o = 2
        for p, g in zip(leaves, grads):
            n = g.numel()
            p.grad = flat[o:o + n].view_as(p).to(p.dtype).clone()
            o += n
        return flat[0].to(torch.float64) + flat[1].to(torch.float64)


def attach(model, global_size: int, person_offset: int, group=None):
    """Mark an estimator (built on this rank's person block) as one shard of a larger job."""
    model.global_size = int(global_size)
    model.person_offset = int(person_offset)
    model.reducer = GradReducer(group)
    return model


def broadcast_replicated(leaves: List[torch.Tensor], src: int = 0, group=None):
    """Make replicated parameters identical on every rank before the first step."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in leaves:
            dist.broadcast(p.data, src, group=group)

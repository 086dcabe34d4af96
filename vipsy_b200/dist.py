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
    """Callable(loss, leaves) -> loss: all-reduce(sum) of [loss_hi, loss_lo, grads...].

    ``peer=True`` (CUDA, one node): the exchange is ``PeerAllReduce`` -- one kernel over NVLink peer memory, the
    loss travelling as fp64 -- instead of a collective-library call; same sums, taken in rank order."""

    def __init__(self, group=None, peer: bool = False):
        self.group = group
        self.bytes_last = 0
        self.peer = bool(peer)
        self._par = None

    def _reduce_peer(self, loss, leaves, grads):
        n = sum(g.numel() for g in grads)
        if self._par is None or self._par.n_float != n:
            self._par = PeerAllReduce(n_float=n, n_double=1, device=loss.device, group=self.group)
        f, d = self._par.slot_views()
        o = 0
        for g in grads:
            f[o:o + g.numel()].copy_(g.reshape(-1))
            o += g.numel()
        d.copy_(loss.detach().to(torch.float64).reshape(1))
        out_f, out_d = self._par.reduce()
        self.bytes_last = n * 4 + 8
        o = 0
        for p, g in zip(leaves, grads):
            k = g.numel()
            p.grad = out_f[o:o + k].view_as(p).to(p.dtype).clone()
            o += k
        return out_d[0].clone()

    def __call__(self, loss: torch.Tensor, leaves: Sequence[torch.Tensor]) -> torch.Tensor:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return loss
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in leaves]
        if self.peer and loss.is_cuda:
            return self._reduce_peer(loss, leaves, grads)
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


class PeerAllReduce:
    """The step's one exchange as ONE kernel over NVLink peer memory (``vipsy_peer_allreduce``): every rank's
    producer kernels write their packed item gradients + loss straight into a symmetric buffer
    (``torch.distributed._symmetric_memory``), the kernel publishes a flag to every peer, waits for theirs and
    sums all payloads in rank order -- no collective-library call, bit-identical results on all ranks.

        par = PeerAllReduce(n_float=ncol, n_double=1, device=dev)
        g, l = par.slot_views()                 # where this step's producers must write
        ops.irt_elbo_grad(..., item_grads_out=g, loss_out=l)
        grads, loss = par.reduce()              # summed over ranks (fp32 / fp64 device tensors)
    """

    FLAG_BYTES = 1024

    def __init__(self, n_float: int, n_double: int = 1, device=None, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _native as N
        self._N = N
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world * 4 > self.FLAG_BYTES:
            raise ValueError("PeerAllReduce: world size too large for the flag area")
        self.n_float, self.n_double = int(n_float), int(n_double)
        self.slot_bytes = ((self.n_float * 4 + 7) // 8) * 8 + self.n_double * 8
        self.slot_bytes = ((self.slot_bytes + 255) // 256) * 256
        total = self.FLAG_BYTES + 2 * self.slot_bytes
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=self.buf.device)
        self.out_f = torch.empty(max(self.n_float, 1), dtype=torch.float32, device=self.buf.device)
        self.out_d = torch.empty(max(self.n_double, 1), dtype=torch.float64, device=self.buf.device)
        self.epoch = 0
        torch.cuda.synchronize(self.buf.device)
        self.hdl.barrier()                       # flags are zero everywhere before the first epoch
        torch.cuda.synchronize(self.buf.device)

    def slot_views(self):
        """(fp32[n_float], fp64[n_double]) views of the slot the NEXT ``reduce()`` will read."""
        off = self.FLAG_BYTES + ((self.epoch + 1) & 1) * self.slot_bytes
        doff = off + ((self.n_float * 4 + 7) // 8) * 8
        f = self.buf[off:off + self.n_float * 4].view(torch.float32)
        d = self.buf[doff:doff + self.n_double * 8].view(torch.float64)
        return f, d

    def reduce(self):
        self.epoch += 1
        L = self._N.lib()
        self._N.check(L.vipsy_peer_allreduce(self.peer_ptrs.data_ptr(), self.rank, self.world, self.FLAG_BYTES,
                                             self.slot_bytes, self.n_float, self.n_double, self.out_f.data_ptr(),
                                             self.out_d.data_ptr(), self.epoch & 0xffffffff,
                                             torch.cuda.current_stream(self.buf.device).cuda_stream))
        return self.out_f[:self.n_float], self.out_d[:self.n_double]


def attach(model, global_size: int, person_offset: int, group=None, peer: bool = False):
    """Mark an estimator (built on this rank's person block) as one shard of a larger job."""
    model.global_size = int(global_size)
    model.person_offset = int(person_offset)
    model.reducer = GradReducer(group, peer=peer)
    return model


def broadcast_replicated(leaves: List[torch.Tensor], src: int = 0, group=None):
    """Make replicated parameters identical on every rank before the first step."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in leaves:
            dist.broadcast(p.data, src, group=group)

"""Host-buffer entry to the fused path: what a binding of
``loss_and_grads(model, guide, data)`` (vi.py:507) calls when the response
matrix and the variational parameters live in (pinned) host memory.

The response matrix may be handed over in the 2-bit transport format of
``ops.pack_responses_2bit_host`` (4 cells per byte; packed once when the data are
loaded): the chunk is expanded to byte codes on the device right after its copy.

The persons are cut into chunks; chunk i+1's host->device copy and chunk i-1's
device->host copy of the per-person gradients overlap chunk i's kernel on three
CUDA streams.  Item-level gradients and the loss are accumulated on the device
and read back once.  Every byte moved is counted (``h2d_bytes`` / ``d2h_bytes``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops


class IrtHostStep:
    def __init__(self, n_persons: int, n_items: int, model: str = "irt_4pl", n_chunks: int = 8, device="cuda"):
        self.N, self.J, self.model = n_persons, n_items, model
        self.dev = torch.device(device)
        self.n_chunks = max(1, min(n_chunks, n_persons))
        per = -(-n_persons // self.n_chunks)
        per = -(-per // 32) * 32                      # whole warps of persons per chunk
        self.bounds = [(s, min(s + per, n_persons)) for s in range(0, n_persons, per)]
        d = self.dev
        self.y = torch.empty((n_persons, n_items), dtype=torch.uint8, device=d)
        self.y2 = None                                # device staging of 2-bit rows, allocated on first use
        self.mu = torch.empty((n_persons, 1), dtype=torch.float32, device=d)
        self.sigma = torch.empty((n_persons, 1), dtype=torch.float32, device=d)
        self.d_mu = torch.empty((n_persons, 1), dtype=torch.float32, device=d)
        self.d_sigma = torch.empty((n_persons, 1), dtype=torch.float32, device=d)
        self.items = torch.empty((4, n_items), dtype=torch.float32, device=d)
        self.h_d_mu = torch.empty((n_persons, 1), dtype=torch.float32, pin_memory=True)
        self.h_d_sigma = torch.empty((n_persons, 1), dtype=torch.float32, pin_memory=True)
        self.h_items = torch.empty((4, n_items), dtype=torch.float32, pin_memory=True)
        self.h_item_grads = None
        self.h_loss = torch.empty(1, dtype=torch.float64, pin_memory=True)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(d) for _ in range(3))
        self.h2d_bytes = self.d2h_bytes = 0
        self.kernel_launches = 0

    def __call__(self, y: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, a, b, c=None, d=None, *,
                 n_particles: int = 1, D: float = 1.0, scale: float = 1.0, seed: int = 0, offset: int = 0,
                 person_offset: int = 0, estimator: str = "pathwise", baseline: float = 0.0) -> Dict[str, torch.Tensor]:
        """All tensor arguments are HOST tensors (pinned for async copies).  y: N x J uint8 codes, or
        N x ceil(J/4) uint8 rows in the 2-bit transport format."""
        J = self.J
        packed = y.shape[1] != J
        if packed:
            if y.shape[1] != (J + 3) // 4:
                raise ValueError("y must be N x J codes or N x ceil(J/4) 2-bit rows")
            if self.y2 is None:
                self.y2 = torch.empty((self.N, (J + 3) // 4), dtype=torch.uint8, device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        h = self.h_items
        h[0].copy_(b.reshape(-1)); h[1].copy_((a if a is not None else torch.ones(J)).reshape(-1))
        h[2].copy_((c if c is not None else torch.zeros(J)).reshape(-1))
        h[3].copy_((d if d is not None else torch.ones(J)).reshape(-1))
        h2d = d2h = 0
        keep = []
        item_acc = loss_acc = None
        with torch.cuda.stream(self.s_in):
            self.items.copy_(h, non_blocking=True)
            h2d += h.numel() * 4
        for (s, e) in self.bounds:
            ev_in, ev_run = torch.cuda.Event(), torch.cuda.Event()
            with torch.cuda.stream(self.s_in):
                if packed:
                    self.y2[s:e].copy_(y[s:e], non_blocking=True)
                else:
                    self.y[s:e].copy_(y[s:e], non_blocking=True)
                self.mu[s:e].copy_(mu[s:e], non_blocking=True)
                self.sigma[s:e].copy_(sigma[s:e], non_blocking=True)
                ev_in.record()
                h2d += (e - s) * (y.shape[1] + 8)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in)
                if packed:
                    ops.unpack_responses_2bit(self.y2[s:e], J, out=self.y[s:e])
                    self.kernel_launches += 1
                r = ops.irt_elbo_grad(self.y[s:e], self.mu[s:e], sigma=self.sigma[s:e], n_particles=n_particles,
                                      a=self.items[1], b=self.items[0], c=self.items[2], d=self.items[3],
                                      model=self.model, D=D, scale=scale, seed=seed, offset=offset,
                                      person_offset=person_offset + s, estimator=estimator, baseline=baseline,
                                      d_mu_out=self.d_mu[s:e], d_scale_out=self.d_sigma[s:e])
                self.kernel_launches += 3
                if item_acc is None:
                    item_acc, loss_acc = r["item_grads"], r["loss"]
                else:
                    item_acc = item_acc + r["item_grads"]
                    loss_acc = loss_acc + r["loss"]
                keep.append(r)
                ev_run.record()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_run)
                self.h_d_mu[s:e].copy_(self.d_mu[s:e], non_blocking=True)
                self.h_d_sigma[s:e].copy_(self.d_sigma[s:e], non_blocking=True)
                d2h += (e - s) * 8
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_stream(self.s_run)
            if self.h_item_grads is None or self.h_item_grads.numel() != item_acc.numel():
                self.h_item_grads = torch.empty(item_acc.numel(), dtype=torch.float32, pin_memory=True)
            self.h_item_grads.copy_(item_acc, non_blocking=True)
            self.h_loss.copy_(loss_acc, non_blocking=True)
            d2h += item_acc.numel() * 4 + 8
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)
        self.h2d_bytes, self.d2h_bytes = h2d, d2h
        return {"loss": self.h_loss, "d_mu": self.h_d_mu, "d_sigma": self.h_d_sigma,
                "item_grads": self.h_item_grads}

"""Host-buffer entry to the fused path: what a binding of
``loss_and_grads(model, guide, data)`` (vi.py:507) calls when the response
matrix and the variational parameters live in (pinned) host memory.

Thin wrapper over ``vipsy_irt_host_step`` (csrc/host_step.cu): the chunked
copy-in / expand + kernel / copy-out pipeline over three CUDA streams is
enqueued from C, no interpreter between launches.  The response matrix may be
handed over in the 2-bit transport format of ``ops.pack_responses_2bit_host``
(4 cells per byte; packed once when the data are loaded).  Every byte moved is
counted (``h2d_bytes`` / ``d2h_bytes``).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import torch

from . import _native as N


class IrtHostStep:
    def __init__(self, n_persons: int, n_items: int, model: str = "irt_4pl", n_chunks: int = 4, device="cuda"):
        self.N, self.J, self.model = n_persons, n_items, model
        self.dev = torch.device(device)
        self.n_chunks = max(1, min(n_chunks, max(n_persons // 32, 1)))
        self.m = N.IRT_MODEL[model]
        self.h_d_mu = torch.empty((n_persons, 1), dtype=torch.float32, pin_memory=True)
        self.h_d_sigma = torch.empty((n_persons, 1), dtype=torch.float32, pin_memory=True)
        self.h_item_grads = torch.empty(self.m * n_items, dtype=torch.float32, pin_memory=True)
        self.h_loss = torch.empty(1, dtype=torch.float64, pin_memory=True)
        self.h_items = torch.empty((4, n_items), dtype=torch.float32, pin_memory=True)
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.scratch = None
        self.h2d_bytes = self.d2h_bytes = 0
        self.kernel_launches = 0

    def __call__(self, y: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, a, b, c=None, d=None, *,
                 n_particles: int = 1, D: float = 1.0, scale: float = 1.0, seed: int = 0, offset: int = 0,
                 person_offset: int = 0, estimator: str = "pathwise", baseline: float = 0.0,
                 a_free=None, n_masked=None) -> Dict[str, torch.Tensor]:
        """All tensor arguments are HOST tensors (pinned for asynchronous copies).  y: N x J uint8 codes, or
        N x ceil(J/4) uint8 rows in the 2-bit transport format.  ``a_free``: optional J-element 0/1 host mask of
        site 'a' (vi.py:511-512); d_a of frozen entries comes back 0, as on the device entry.  ``n_masked``: the
        number of missing cells of the whole matrix if the caller knows it (it is constant between steps): saves the
        counting pass of every chunk.  2-bit rows with S > 1, J % 4 == 0 and the pathwise estimator are expanded by
        the particle kernel itself while it stages a tile (no byte matrix in HBM, no expansion launch).  Returns
        pinned host tensors."""
        J, m = self.J, self.m
        for t, nm in ((y, "y"), (mu, "mu"), (sigma, "sigma")):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError(f"{nm} must be a contiguous HOST tensor")
        packed = y.shape[1] != J
        if packed and y.shape[1] != (J + 3) // 4:
            raise ValueError("y must be N x J codes or N x ceil(J/4) 2-bit rows")
        if y.dtype != torch.uint8 or mu.dtype != torch.float32 or sigma.dtype != torch.float32:
            raise ValueError("y must be uint8, mu / sigma float32")
        h = self.h_items
        h[0].copy_(b.reshape(-1))
        if m >= 2:
            h[1].copy_(a.reshape(-1))
        if m >= 3:
            h[2].copy_(c.reshape(-1))
        if m >= 4:
            h[3].copy_(d.reshape(-1))
        args = N.IrtHostArgs()
        dv = args.dev
        dv.n_rows, dv.batch, dv.n_items, dv.n_dims, dv.n_particles = self.N, self.N, J, 1, int(n_particles)
        dv.model, dv.estimator = m, N.EST[estimator]
        dv.D, dv.scale, dv.baseline = float(D), float(scale), float(baseline)
        dv.seed, dv.offset, dv.person_offset = seed, offset, person_offset
        if n_masked is not None:
            dv.flags |= N.IRT_FLAG_MASKED_KNOWN
            dv.n_masked_cells = int(n_masked)
        args.y_host, args.y_packed2, args.n_chunks = y.data_ptr(), int(packed), self.n_chunks
        args.mu_host, args.sigma_host = mu.data_ptr(), sigma.data_ptr()
        args.b_host = h[0].data_ptr()
        args.a_host = h[1].data_ptr() if m >= 2 else None
        args.c_host = h[2].data_ptr() if m >= 3 else None
        args.d_host = h[3].data_ptr() if m >= 4 else None
        args.loss_host = self.h_loss.data_ptr()
        args.d_mu_host, args.d_sigma_host = self.h_d_mu.data_ptr(), self.h_d_sigma.data_ptr()
        args.item_grads_host = self.h_item_grads.data_ptr()
        if a_free is not None and m >= 2:
            self._a_free = a_free.reshape(-1).to(torch.uint8).contiguous()        # kept alive until the sync below
            if self._a_free.numel() != J or self._a_free.is_cuda:
                raise ValueError("a_free must be a host tensor with J elements")
            args.a_free_host = self._a_free.data_ptr()
        L = N.lib()
        need = L.vipsy_irt_host_scratch_bytes(C.byref(args))
        if self.scratch is None or self.scratch.numel() < need:
            self.scratch = torch.empty(need, dtype=torch.uint8, device=self.dev)
        args.dev_scratch, args.dev_scratch_bytes = self.scratch.data_ptr(), self.scratch.numel()
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        args.s_in, args.s_run, args.s_out = self.s_in.cuda_stream, self.s_run.cuda_stream, self.s_out.cuda_stream
        N.check(L.vipsy_irt_host_step(C.byref(args)))
        self.s_out.synchronize()
        cur.wait_stream(self.s_run)
        chunks = self.n_chunks
        direct2 = packed and n_particles > 1 and estimator == "pathwise" and J % 4 == 0     # as in csrc/host_step.cu
        self.kernel_launches += chunks * (3 + int(packed and not direct2)) + 1
        self.h2d_bytes = self.N * (y.shape[1] + 8) + m * J * 4 + (J if args.a_free_host else 0)
        self.d2h_bytes = self.N * 8 + m * J * 4 + 8
        return {"loss": self.h_loss, "d_mu": self.h_d_mu, "d_sigma": self.h_d_sigma, "item_grads": self.h_item_grads}

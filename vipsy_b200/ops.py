"""Tensor-level wrappers over the C ABI (include/vipsy_b200.h).

Each function takes CUDA tensors, launches on torch's current stream and returns
CUDA tensors without synchronising.  These are what the estimator classes in
``vipsy_b200.models`` call once per ``SVI.step`` (vi.py:503-516).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _native as N

_workspaces: Dict[tuple, torch.Tensor] = {}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32(t: Optional[torch.Tensor], name: str, numel: Optional[int] = None):
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if numel is not None and t.numel() != numel:
        raise ValueError(f"{name}: expected {numel} elements, got {tuple(t.shape)}")
    return t


def _workspace(kind: str, nbytes: int, device) -> torch.Tensor:
    key = (kind, device.index, torch.cuda.current_stream().cuda_stream)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def pack_responses(y: torch.Tensor) -> torch.Tensor:
    """float matrix with NaN = missing (the reference's ``data``) -> uint8 codes 0/1/255."""
    if y.dtype == torch.uint8:
        return y.contiguous()
    yf = _f32(y, "y")
    out = torch.empty(yf.shape, dtype=torch.uint8, device=yf.device)
    N.check(N.lib().vipsy_pack_responses(_ptr(yf), yf.numel(), _ptr(out), _stream()))
    return out


def pack_responses_2bit_host(codes: torch.Tensor, pin: bool = True) -> torch.Tensor:
    """HOST uint8 code matrix (0/1/255) -> HOST 2-bit transport format, n_rows x ceil(J/4) bytes (done once)."""
    if codes.is_cuda or codes.dtype != torch.uint8 or codes.dim() != 2 or codes.stride(1) != 1:
        raise ValueError("codes must be a host uint8 matrix with unit column stride")
    n_rows, J = codes.shape
    out = torch.empty((n_rows, (J + 3) // 4), dtype=torch.uint8, pin_memory=pin)
    N.check(N.lib().vipsy_pack_responses_2bit_host(codes.data_ptr(), n_rows, J, codes.stride(0), out.data_ptr()))
    return out


def unpack_responses_2bit(packed: torch.Tensor, n_items: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """DEVICE 2-bit rows -> uint8 code matrix (stream-ordered)."""
    if not packed.is_cuda or packed.dtype != torch.uint8 or packed.dim() != 2 or not packed.is_contiguous():
        raise ValueError("packed must be a contiguous CUDA uint8 matrix")
    n_rows = packed.shape[0]
    if packed.shape[1] != (n_items + 3) // 4:
        raise ValueError("packed rows must hold ceil(J/4) bytes")
    if out is None:
        out = torch.empty((n_rows, n_items), dtype=torch.uint8, device=packed.device)
    N.check(N.lib().vipsy_unpack_responses_2bit(_ptr(packed), n_rows, n_items, _ptr(out), out.stride(0), _stream()))
    return out


def observed_csr(y: torch.Tensor):
    """uint8 code matrix -> (row_ptr int64[n_rows+1], entries uint16[nnz], nnz): the observed cells only,
    entry = item | response << 15.  Built once per response matrix (the matrix is constant across SVI steps);
    ``irt_elbo_grad(..., obs=...)`` then streams 2 bytes per OBSERVED cell instead of 1 byte per cell."""
    if not y.is_cuda or y.dtype != torch.uint8 or y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be a CUDA uint8 code matrix with unit column stride")
    n_rows, J = y.shape
    counts = torch.empty(max(n_rows, 1), dtype=torch.int32, device=y.device)
    row_ptr = torch.empty(n_rows + 1, dtype=torch.int64, device=y.device)
    L = N.lib()
    N.check(L.vipsy_observed_row_ptr(_ptr(y), n_rows, J, y.stride(0), _ptr(counts), _ptr(row_ptr), _stream()))
    nnz = int(row_ptr[-1].item())
    entries = torch.empty(max(nnz, 1), dtype=torch.uint16, device=y.device)
    N.check(L.vipsy_observed_fill(_ptr(y), n_rows, J, y.stride(0), _ptr(row_ptr), _ptr(entries), _stream()))
    return row_ptr, entries, nnz


def fill_normal(S: int, B: int, K: int, seed: int, offset, person_offset: int = 0, device="cuda"):
    """S x B x K draws of the in-kernel N(0,1) generator.  ``offset``: the call counter, an int or a one-element
    int64 CUDA tensor read at execution time (CUDA-graph replays then draw fresh noise)."""
    out = torch.empty((S, B, K), dtype=torch.float32, device=device)
    if isinstance(offset, torch.Tensor):
        N.check(N.lib().vipsy_fill_normal_dev(_ptr(out), S, B, K, seed, _ptr(offset), person_offset, _stream()))
    else:
        N.check(N.lib().vipsy_fill_normal(_ptr(out), S, B, K, seed, offset, person_offset, _stream()))
    return out


def fill_uniform(S: int, B: int, K: int, seed: int, offset, person_offset: int = 0, device="cuda"):
    out = torch.empty((S, B, K), dtype=torch.float32, device=device)
    if isinstance(offset, torch.Tensor):
        N.check(N.lib().vipsy_fill_uniform_dev(_ptr(out), S, B, K, seed, _ptr(offset), person_offset, _stream()))
    else:
        N.check(N.lib().vipsy_fill_uniform(_ptr(out), S, B, K, seed, offset, person_offset, _stream()))
    return out


def adam_step(tensors, step_dev: torch.Tensor, zero_grad: bool = True):
    """One launch of the multi-tensor Adam (+ free mask, + zero grad) over ``tensors``: a list of dicts with
    param, grad, exp_avg, exp_avg_sq (fp32 CUDA, same numel), optional free_mask (uint8), lr_dev, and the
    scalars lr, beta1, beta2, eps, step_offset.  ``step_dev``: int64 CUDA scalar, the step count after increment."""
    n = len(tensors)
    if n == 0:
        return
    arr = (N.AdamTensor * n)()
    for i, t in enumerate(tensors):
        p = t["param"]
        for k in ("param", "grad", "exp_avg", "exp_avg_sq"):
            x = t[k]
            if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.numel() == p.numel()):
                raise ValueError(f"adam_step: {k} must be a contiguous fp32 CUDA tensor of the parameter's size")
        a = arr[i]
        a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), t["grad"].data_ptr(), t["exp_avg"].data_ptr(), t["exp_avg_sq"].data_ptr()
        fm = t.get("free_mask")
        if fm is not None and not (fm.is_cuda and fm.dtype == torch.uint8 and fm.is_contiguous() and fm.numel() == p.numel()):
            raise ValueError("adam_step: free_mask must be a contiguous uint8 CUDA tensor of the parameter's size")
        a.free_mask = _ptr(fm)
        a.lr_dev = _ptr(t.get("lr_dev"))
        a.n, a.step_offset = p.numel(), int(t.get("step_offset", 0))
        a.lr, a.beta1, a.beta2, a.eps = float(t["lr"]), float(t["beta1"]), float(t["beta2"]), float(t["eps"])
    N.check(N.lib().vipsy_adam_step(arr, n, _ptr(step_dev), int(zero_grad), _stream()))


def qmask_from_q(q: torch.Tensor) -> torch.Tensor:
    """K x J 0/1 Q-matrix (vi.py:741) -> J uint32-in-int32 bit masks, bit k = Q[k][j]."""
    K, J = q.shape
    w = (2 ** torch.arange(K, dtype=torch.int64, device=q.device)).unsqueeze(1)
    return ((q != 0).to(torch.int64) * w).sum(0).to(torch.int32).contiguous()


def cdm_eta_table(qmask: torch.Tensor, K: int, model: str) -> torch.Tensor:
    J = qmask.numel()
    out = torch.empty((2 ** K, J), dtype=torch.uint8, device=qmask.device)
    N.check(N.lib().vipsy_cdm_eta_table(_ptr(qmask), J, K, N.CDM_MODEL[model], _ptr(out), _stream()))
    return out


def irt_elbo_grad(
    y: torch.Tensor,                       # n_rows x J uint8 codes (pack_responses)
    mu: torch.Tensor,                      # B x K
    *,
    sigma: Optional[torch.Tensor] = None,  # K == 1: B (x1)
    chol: Optional[torch.Tensor] = None,   # K > 1: K x K | B x K x K | B x K(K+1)/2
    eps: Optional[torch.Tensor] = None,    # S x B x K, or None -> in-kernel Philox
    n_particles: Optional[int] = None,
    a: Optional[torch.Tensor] = None, b: torch.Tensor = None,
    c: Optional[torch.Tensor] = None, d: Optional[torch.Tensor] = None,
    model: str = "irt_2pl", D: float = 1.0, scale: float = 1.0,
    estimator: str = "pathwise", baseline: float = 0.0,
    idx: Optional[torch.Tensor] = None,
    a_free: Optional[torch.Tensor] = None,
    chol_packed: bool = False, chol_raw: bool = False,
    seed: int = 0, offset: int = 0, person_offset: int = 0,
    want_grad: bool = True, strict: bool = False, tensor_cores: bool = True, sparse: bool = False,
    row_kernel: bool = True, obs=None, n_masked: Optional[int] = None, fused: bool = True,
    d_mu_out: Optional[torch.Tensor] = None, d_scale_out: Optional[torch.Tensor] = None,
    item_grads_out: Optional[torch.Tensor] = None, loss_out: Optional[torch.Tensor] = None,
    mvn_encoder=None,
) -> Dict[str, Optional[torch.Tensor]]:
    """One fused ELBO + gradient evaluation (Trace_ELBO.loss_and_grads, vi.py:507).

    ``mvn_encoder=(hidden, fc22_weight, fc22_bias)`` (K > 1, instead of ``chol``): the amortised guide of
    VaeIRT (vi.py:685-693).  ``mu`` is fc21's output, ``hidden`` = softplus(fc1(x)) for the batch rows; fc22, the
    tril scatter and LowerCholeskyTransform are evaluated inside the sampler on the tensor cores and the B x K(K+1)/2
    factor matrix never exists in HBM.  Returns ``d_hidden`` (through fc22 only), ``d_w22``, ``d_b22`` instead of
    ``d_chol``.  With five entries ``(hidden, fc22.weight, fc22.bias, fc21.weight, fc21.bias)`` the location head is
    evaluated inside the call as well (``mu`` may be None): ``d_w21`` / ``d_b21`` are returned too and ``d_hidden``
    holds the gradient through both heads.

    ``item_grads_out`` (fp32, packed [d_b | d_a | d_c | d_d]) / ``loss_out`` (fp64, 1): optional destinations,
    e.g. views into a symmetric peer-memory buffer (``dist.PeerAllReduce``).

    ``d_mu_out`` / ``d_scale_out``: optional preallocated fp32 CUDA buffers for the
    per-person gradients (d_sigma for K == 1, d_chol otherwise).  ``tensor_cores=False``
    forces the SIMT contraction for K > 1 (default: tcgen05 for K >= 8).  ``sparse=True`` (K == 1)
    tells the kernel most cells are missing: it compacts each person's observed cells first.
    ``obs``: the (row_ptr, entries, nnz) triple of ``observed_csr(y)``; used by the K == 1 pathwise kernels at any S.
    ``n_masked``: number of missing cells among the batch rows, if the caller knows it (saves a counting pass).
    ``fused=False`` (2 <= K <= 8, per-person factor) forces the three-kernel path instead of the fused kernel.
    ``row_kernel=False`` (K == 1, S == 1) forces the lane-per-person kernels instead of the warp-per-person ones."""
    if not y.is_cuda or y.dtype != torch.uint8:
        raise ValueError("y must be a CUDA uint8 code matrix (see pack_responses)")
    if y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be 2-D with unit column stride")
    dev = y.device
    n_rows, J = y.shape
    if mu is None and mvn_encoder is not None and len(mvn_encoder) == 5:
        # fc21 fused as well: the location is computed inside the call; a placeholder carries the shape only
        mu_ = torch.empty((mvn_encoder[0].shape[0], mvn_encoder[3].shape[0]), dtype=torch.float32, device=y.device)
    else:
        mu_ = _f32(mu, "mu")
    if mu_.dim() == 1:
        mu_ = mu_.unsqueeze(1)
    B, K = mu_.shape
    if eps is not None:
        eps_ = _f32(eps, "eps")
        S = eps_.numel() // max(B * K, 1)
        if eps_.numel() != S * B * K:
            raise ValueError(f"eps must be S x B x K, got {tuple(eps.shape)}")
    else:
        eps_ = None
        S = int(n_particles or 1)
    m = N.IRT_MODEL[model]
    args = N.IrtArgs()
    args.n_rows, args.batch, args.n_items, args.n_dims, args.n_particles = n_rows, B, J, K, S
    args.model, args.estimator = m, N.EST[estimator]
    args.D, args.scale, args.baseline = float(D), float(scale), float(baseline)
    args.flags = (N.IRT_FLAG_STRICT if strict else 0) | (0 if tensor_cores else N.IRT_FLAG_NO_TENSOR) \
        | (N.IRT_FLAG_SPARSE if sparse else 0) | (0 if row_kernel else N.IRT_FLAG_NO_ROW) \
        | (0 if fused else N.IRT_FLAG_NO_FUSED)
    args.y, args.y_row_stride = _ptr(y), y.stride(0)
    if n_masked is not None:
        args.flags |= N.IRT_FLAG_MASKED_KNOWN
        args.n_masked_cells = int(n_masked)
    if obs is not None:
        rp, en, nnz = obs
        if rp.numel() != n_rows + 1 or rp.dtype != torch.int64 or en.dtype != torch.uint16:
            raise ValueError("obs must be the (row_ptr, entries, nnz) triple of observed_csr(y)")
        args.obs_row_ptr, args.obs_entries, args.obs_nnz = _ptr(rp), _ptr(en), int(nnz)
    idx_ = None
    if idx is not None:
        idx_ = idx.to(device=dev, dtype=torch.int64).contiguous()
        if idx_.numel() != B:
            raise ValueError("idx must have B entries")
    args.idx = _ptr(idx_)
    args.mu, args.eps = _ptr(mu_), _ptr(eps_)
    args.seed, args.offset, args.person_offset = seed, offset, person_offset
    a_ = _f32(a, "a", K * J) if m >= 2 else None
    b_ = _f32(b, "b", J)
    c_ = _f32(c, "c", J) if m >= 3 else None
    d_ = _f32(d, "d", J) if m >= 4 else None
    args.a, args.b, args.c, args.d = _ptr(a_), _ptr(b_), _ptr(c_), _ptr(d_)
    af_ = None
    if a_free is not None and m >= 2:
        af_ = a_free.to(device=dev, dtype=torch.uint8).contiguous()
    args.a_free = _ptr(af_)

    out: Dict[str, Optional[torch.Tensor]] = {}
    if loss_out is not None:
        if not (loss_out.is_cuda and loss_out.dtype == torch.float64 and loss_out.numel() == 1):
            raise ValueError("loss_out must be a one-element fp64 CUDA tensor")
        loss = loss_out
    else:
        loss = torch.empty(1, dtype=torch.float64, device=dev)
    args.loss = _ptr(loss)
    sg_ = ch_ = None
    enc_ = None
    if mvn_encoder is not None:
        if K < 2:
            raise ValueError("mvn_encoder is the K > 1 guide")
        tri = K * (K + 1) // 2
        enc_ = tuple(_f32(t, n) for t, n in zip(mvn_encoder, ("hidden", "fc22.weight", "fc22.bias", "fc21.weight", "fc21.bias")))
        H = enc_[0].shape[1]
        if enc_[0].shape[0] != B or tuple(enc_[1].shape) != (tri, H) or enc_[2].numel() != tri:
            raise ValueError("mvn_encoder: hidden B x H, fc22.weight K(K+1)/2 x H, fc22.bias K(K+1)/2")
        if len(enc_) not in (3, 5) or (len(enc_) == 5 and (tuple(enc_[3].shape) != (K, H) or enc_[4].numel() != K)):
            raise ValueError("mvn_encoder: (hidden, fc22.weight, fc22.bias[, fc21.weight K x H, fc21.bias K])")
        if H > MVN_ENCODER_MAX_HIDDEN:
            raise ValueError(f"mvn_encoder: hidden_dim {H} > {MVN_ENCODER_MAX_HIDDEN}")
        args.cov_mode = N.COV_PERSON
    elif K == 1:
        sg_ = _f32(sigma, "sigma", B)
        args.sigma, args.cov_mode = _ptr(sg_), N.COV_DIAG1
    else:
        ch_ = _f32(chol, "chol")
        tri = K * (K + 1) // 2
        if ch_.numel() == K * K and not chol_packed and ch_.dim() == 2:
            args.cov_mode, args.chol_layout = N.COV_SHARED, N.CHOL_DENSE
        elif chol_packed:
            if ch_.numel() == tri:
                args.cov_mode = N.COV_SHARED
            elif ch_.numel() == B * tri:
                args.cov_mode = N.COV_PERSON
            else:
                raise ValueError("packed chol must hold K(K+1)/2 or B*K(K+1)/2 values")
            args.chol_layout = N.CHOL_PACKED
        elif ch_.numel() == B * K * K:
            args.cov_mode, args.chol_layout = N.COV_PERSON, N.CHOL_DENSE
        else:
            raise ValueError(f"chol shape {tuple(chol.shape)} does not match K={K}, B={B}")
        args.chol, args.chol_raw = _ptr(ch_), int(bool(chol_raw))
    if want_grad:
        def _out(buf, shape, zero=False):
            if buf is None:
                return (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=dev)
            if not (buf.is_cuda and buf.dtype == torch.float32 and buf.is_contiguous()
                    and buf.numel() == math.prod(shape)):
                raise ValueError("preallocated gradient buffer has the wrong dtype/size/layout")
            return buf
        d_mu = _out(d_mu_out, (B, K))
        args.d_mu = _ptr(d_mu)
        out["d_mu"] = d_mu
        if K == 1:
            d_sigma = _out(d_scale_out, (B, 1))
            args.d_sigma = _ptr(d_sigma)
            out["d_sigma"] = d_sigma
        elif enc_ is not None:
            out["d_hidden"] = torch.empty_like(enc_[0])
            out["d_w22"] = torch.empty_like(enc_[1])
            if len(enc_) == 5:
                out["d_w21"] = torch.empty_like(enc_[3])
                out["d_b21"] = torch.empty_like(enc_[4])
            out["d_b22"] = torch.empty(enc_[2].numel(), dtype=torch.float32, device=dev)
        else:
            d_chol = _out(d_scale_out, tuple(ch_.shape), zero=d_scale_out is None)
            args.d_chol = _ptr(d_chol)
            out["d_chol"] = d_chol
        # item gradients live in one packed fp32 buffer [d_b | d_a | d_c | d_d] so a
        # person-sharded job all-reduces it with a single collective
        ncol = J + (K * J if m >= 2 else 0) + (J if m >= 3 else 0) + (J if m >= 4 else 0)
        if item_grads_out is not None:
            if not (item_grads_out.is_cuda and item_grads_out.dtype == torch.float32 and item_grads_out.is_contiguous()
                    and item_grads_out.numel() == ncol):
                raise ValueError(f"item_grads_out must be a contiguous fp32 CUDA tensor of {ncol} elements")
            packed = item_grads_out.view(-1)
        else:
            packed = torch.empty(ncol, dtype=torch.float32, device=dev)
        out["item_grads"] = packed
        o = 0
        out["d_b"] = packed[o:o + J].view(1, J); o += J
        args.d_b = out["d_b"].data_ptr()
        if m >= 2:
            out["d_a"] = packed[o:o + K * J].view(K, J); o += K * J
            args.d_a = out["d_a"].data_ptr()
        if m >= 3:
            out["d_c"] = packed[o:o + J].view(1, J); o += J
            args.d_c = out["d_c"].data_ptr()
        if m >= 4:
            out["d_d"] = packed[o:o + J].view(1, J); o += J
            args.d_d = out["d_d"].data_ptr()
    L = N.lib()
    if enc_ is not None:
        ea = N.MvnEncoderIrtArgs()
        ea.irt = args
        ea.n_hidden, ea.hidden, ea.w22, ea.b22 = enc_[0].shape[1], _ptr(enc_[0]), _ptr(enc_[1]), _ptr(enc_[2])
        if len(enc_) == 5:
            ea.w21, ea.b21 = _ptr(enc_[3]), _ptr(enc_[4])
        if want_grad:
            ea.d_hidden, ea.d_w22, ea.d_b22 = _ptr(out["d_hidden"]), _ptr(out["d_w22"]), _ptr(out["d_b22"])
            if len(enc_) == 5:
                ea.d_w21, ea.d_b21 = _ptr(out["d_w21"]), _ptr(out["d_b21"])
        nbytes = L.vipsy_irt_mvn_encoder_workspace_bytes(C.byref(ea))
        ws = _workspace("irt_mvn", nbytes, dev)
        ea.workspace, ea.workspace_bytes = _ptr(ws), ws.numel()
        N.check(L.vipsy_irt_mvn_encoder_elbo_grad(C.byref(ea), _stream()))
    else:
        nbytes = L.vipsy_irt_workspace_bytes(C.byref(args))
        ws = _workspace("irt", nbytes, dev)
        args.workspace, args.workspace_bytes = _ptr(ws), ws.numel()
        N.check(L.vipsy_irt_elbo_grad(C.byref(args), _stream()))
    out["loss"] = loss
    out["_keepalive"] = (mu_, sg_, ch_, eps_, a_, b_, c_, d_, af_, idx_, ws, obs, enc_)
    return out


MVN_ENCODER_MAX_HIDDEN = 64


def _mvn_workspace_bytes(K: int, H: int) -> int:
    """Bytes of the fc22 operand image (csrc/enc_mvn.cuh: mv_chunks x mv_chunk_bytes, rounded to 256)."""
    nb = (K + 7) // 8
    pp = 32 * nb * (nb + 1)
    kh = (H + 1 + 15) // 16 * 16
    return ((pp + 127) // 128 * (2 * 128 * kh * 2) + 255) & ~255


def mvn_encoder_sample(mu: Optional[torch.Tensor], hidden: torch.Tensor, w22: torch.Tensor, b22: torch.Tensor, *,
                       w21: Optional[torch.Tensor] = None, b21: Optional[torch.Tensor] = None,
                       eps: Optional[torch.Tensor] = None, n_particles: int = 1, seed: int = 0, offset: int = 0,
                       person_offset: int = 0) -> Dict[str, torch.Tensor]:
    """The sample site of VaeIRT's amortised MVN guide (vi.py:693) with MvnEncoder's fc22 fused in:
    x [S, B, K] = mu + L eps with L = tril(raw, -1) + diag(exp(raw_ii)), raw = fc22(hidden) evaluated on the
    tensor cores and consumed on chip.  Also returns the draws, diag L and sum_i log L_ii.  With ``w21`` / ``b21``
    (fc21) the location is computed inside the call too and ``mu`` may be None."""
    h_, w_, b_ = _f32(hidden, "hidden"), _f32(w22, "w22"), _f32(b22, "b22")
    w21_, b21_ = _f32(w21, "w21", None), _f32(b21, "b21", None)
    if mu is None:
        if w21_ is None or b21_ is None:
            raise ValueError("mvn_encoder_sample: mu or (w21, b21) is required")
        mu_ = torch.empty((h_.shape[0], w21_.shape[0]), dtype=torch.float32, device=h_.device)   # shape only
    else:
        mu_ = _f32(mu, "mu")
    B, K = mu_.shape
    H = h_.shape[1]
    if w21_ is not None and (tuple(w21_.shape) != (K, H) or b21_ is None or b21_.numel() != K):
        raise ValueError("w21 must be K x H with b21 of K entries")
    if h_.shape[0] != B or w_.shape != (K * (K + 1) // 2, H) or b_.numel() != K * (K + 1) // 2:
        raise ValueError("hidden must be B x H, w22 K(K+1)/2 x H, b22 K(K+1)/2")
    eps_ = _f32(eps, "eps", None)
    S = int(n_particles) if eps_ is None else eps_.numel() // max(B * K, 1)
    dev = mu_.device
    a = N.MvnEncoderIrtArgs()
    a.irt.batch = a.irt.n_rows = B
    a.irt.n_dims, a.irt.n_particles, a.irt.n_items = K, S, 1
    a.irt.mu, a.irt.eps = _ptr(mu_), _ptr(eps_)
    a.irt.seed, a.irt.offset, a.irt.person_offset = seed, offset, person_offset
    a.n_hidden, a.hidden, a.w22, a.b22 = H, _ptr(h_), _ptr(w_), _ptr(b_)
    if w21_ is not None:
        a.w21, a.b21 = _ptr(w21_), _ptr(b21_)
    ws = _workspace("mvn_sample", _mvn_workspace_bytes(K, H), dev)
    a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
    out = {"x": torch.empty((S, B, K), dtype=torch.float32, device=dev),
           "eps": torch.empty((S, B, K), dtype=torch.float32, device=dev),
           "diag": torch.empty((B, K), dtype=torch.float32, device=dev),
           "logdet": torch.empty(B, dtype=torch.float32, device=dev)}
    N.check(N.lib().vipsy_mvn_encoder_sample(C.byref(a), _ptr(out["x"]), _ptr(out["eps"]), _ptr(out["diag"]),
                                             _ptr(out["logdet"]), _stream()))
    out["_keepalive"] = (mu_, h_, w_, b_, eps_, ws, w21_, b21_)
    return out


def cdm_enum_elbo_grad(
    y: torch.Tensor,                     # n_rows x J uint8 codes
    q: torch.Tensor,                     # K x J 0/1 Q-matrix (or J int32 masks from qmask_from_q)
    g: torch.Tensor, s: torch.Tensor,    # J guess / slip (constrained values)
    *,
    n_attrs: Optional[int] = None,
    model: str = "dina", scale: float = 1.0,
    batch: Optional[int] = None, idx: Optional[torch.Tensor] = None,
    lam0: Optional[torch.Tensor] = None, lam1: Optional[torch.Tensor] = None,     # HO-DINA (vi.py:901-902)
    mu: Optional[torch.Tensor] = None, sigma: Optional[torch.Tensor] = None,      # theta guide (vi.py:928-929)
    eps: Optional[torch.Tensor] = None, n_particles: int = 1,
    seed: int = 0, offset: int = 0, person_offset: int = 0,
    want_grad: bool = True, want_posterior: bool = False,
    log_prior: Optional[torch.Tensor] = None, score_missing: bool = False,
) -> Dict[str, Optional[torch.Tensor]]:
    """TraceEnum_ELBO.loss_and_grads for VCCDM.model (vi.py:839-857) or, when lam0/lam1/mu/sigma
    are given, VCHoDina (vi.py:897-934): exact marginal over the 2^K attribute patterns.
    ``log_prior`` (B x 2^K) + ``score_missing`` give VaeCCDM.model (vi.py:872-891): per-person pattern prior
    from the encoder, missing cells scored as y = -1; d loss / d log_prior = -scale * out["posterior"]."""
    if not y.is_cuda or y.dtype != torch.uint8 or y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be a CUDA uint8 code matrix with unit column stride")
    dev = y.device
    n_rows, J = y.shape
    if q.dim() == 2:
        K = q.shape[0]
        qm = qmask_from_q(q.to(dev))
    else:
        if n_attrs is None:
            raise ValueError("n_attrs is required when q is given as bit masks")
        K, qm = int(n_attrs), q.to(device=dev, dtype=torch.int32).contiguous()
    hod = lam0 is not None
    args = N.CdmEnumArgs()
    idx_ = None
    if idx is not None:
        idx_ = idx.to(device=dev, dtype=torch.int64).contiguous()
        B = idx_.numel()
    else:
        B = n_rows if batch is None else int(batch)
    g_, s_ = _f32(g, "g", J), _f32(s, "s", J)
    eps_ = None
    S = 1
    if hod:
        l0_, l1_ = _f32(lam0, "lam0", K), _f32(lam1, "lam1", K)
        mu_, sg_ = _f32(mu, "mu", B), _f32(sigma, "sigma", B)
        if eps is not None:
            eps_ = _f32(eps, "eps")
            S = eps_.numel() // max(B, 1)
            if eps_.numel() != S * B:
                raise ValueError("eps must be S x B (x 1)")
        else:
            S = int(n_particles)
        args.lam0, args.lam1, args.mu, args.sigma, args.eps = _ptr(l0_), _ptr(l1_), _ptr(mu_), _ptr(sg_), _ptr(eps_)
    args.n_rows, args.batch, args.n_items, args.n_attrs, args.n_particles = n_rows, B, J, K, S
    args.model, args.hodina, args.scale = N.CDM_MODEL[model], int(hod), float(scale)
    args.y, args.y_row_stride, args.idx, args.qmask = _ptr(y), y.stride(0), _ptr(idx_), _ptr(qm)
    args.g, args.s = _ptr(g_), _ptr(s_)
    args.seed, args.offset, args.person_offset = seed, offset, person_offset
    out: Dict[str, Optional[torch.Tensor]] = {}
    loss = torch.empty(1, dtype=torch.float64, device=dev)
    args.loss = _ptr(loss)
    if want_grad:
        ncol = 2 * J + (2 * K if hod else 0)
        packed = torch.zeros(ncol, dtype=torch.float32, device=dev)
        out["item_grads"] = packed
        out["d_g"], out["d_s"] = packed[:J].view(1, J), packed[J:2 * J].view(1, J)
        args.d_g, args.d_s = out["d_g"].data_ptr(), out["d_s"].data_ptr()
        if hod:
            out["d_lam0"], out["d_lam1"] = packed[2 * J:2 * J + K].view(1, K), packed[2 * J + K:].view(1, K)
            args.d_lam0, args.d_lam1 = out["d_lam0"].data_ptr(), out["d_lam1"].data_ptr()
            out["d_mu"] = torch.empty((B, 1), dtype=torch.float32, device=dev)
            out["d_sigma"] = torch.empty((B, 1), dtype=torch.float32, device=dev)
            args.d_mu, args.d_sigma = _ptr(out["d_mu"]), _ptr(out["d_sigma"])
    if want_posterior:
        out["posterior"] = torch.zeros((B, 2 ** K), dtype=torch.float32, device=dev)
        args.posterior = _ptr(out["posterior"])
    lp_ = None
    if log_prior is not None:
        if hod:
            raise ValueError("log_prior is for the VCCDM / VaeCCDM model, not HO-DINA")
        lp_ = _f32(log_prior, "log_prior", B * 2 ** K)
        args.log_prior = _ptr(lp_)
    args.score_missing = int(bool(score_missing))
    L = N.lib()
    ws = _workspace("cdm_enum", L.vipsy_cdm_enum_workspace_bytes(C.byref(args)), dev)
    args.workspace, args.workspace_bytes = _ptr(ws), ws.numel()
    N.check(L.vipsy_cdm_enum_elbo_grad(C.byref(args), _stream()))
    out["loss"] = loss
    out["_keepalive"] = (qm, g_, s_, idx_, eps_, ws, args, lp_)
    return out


def cdm_score_elbo_grad(
    y: torch.Tensor, q: torch.Tensor, g: torch.Tensor, s: torch.Tensor, attr_p: torch.Tensor, *,
    n_attrs: Optional[int] = None, u: Optional[torch.Tensor] = None, n_particles: int = 1,
    model: str = "dina", scale: float = 1.0, baseline: float = 0.0, idx: Optional[torch.Tensor] = None,
    seed: int = 0, offset: int = 0, person_offset: int = 0, want_grad: bool = True,
) -> Dict[str, Optional[torch.Tensor]]:
    """Trace_ELBO.loss_and_grads for BaseCDM.model + the Bernoulli guide of VCDM / VaeCDM
    (vi.py:745-756, 796-816): REINFORCE.  attr = [u < attr_p] with u injected or Philox."""
    if not y.is_cuda or y.dtype != torch.uint8 or y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be a CUDA uint8 code matrix with unit column stride")
    dev = y.device
    n_rows, J = y.shape
    if q.dim() == 2:
        K, qm = q.shape[0], qmask_from_q(q.to(dev))
    else:
        K, qm = int(n_attrs), q.to(device=dev, dtype=torch.int32).contiguous()
    ap_ = _f32(attr_p, "attr_p")
    B = ap_.numel() // K
    u_ = None
    S = int(n_particles)
    if u is not None:
        u_ = _f32(u, "u")
        S = u_.numel() // max(B * K, 1)
    idx_ = None if idx is None else idx.to(device=dev, dtype=torch.int64).contiguous()
    g_, s_ = _f32(g, "g", J), _f32(s, "s", J)
    args = N.CdmScoreArgs()
    args.n_rows, args.batch, args.n_items, args.n_attrs, args.n_particles = n_rows, B, J, K, S
    args.model, args.scale, args.baseline = N.CDM_MODEL[model], float(scale), float(baseline)
    args.y, args.y_row_stride, args.idx, args.qmask = _ptr(y), y.stride(0), _ptr(idx_), _ptr(qm)
    args.g, args.s, args.attr_p, args.u = _ptr(g_), _ptr(s_), _ptr(ap_), _ptr(u_)
    args.seed, args.offset, args.person_offset = seed, offset, person_offset
    out: Dict[str, Optional[torch.Tensor]] = {}
    loss = torch.empty(1, dtype=torch.float64, device=dev)
    args.loss = _ptr(loss)
    if want_grad:
        packed = torch.zeros(2 * J, dtype=torch.float32, device=dev)
        out["item_grads"] = packed
        out["d_g"], out["d_s"] = packed[:J].view(1, J), packed[J:].view(1, J)
        args.d_g, args.d_s = out["d_g"].data_ptr(), out["d_s"].data_ptr()
        out["d_attr_p"] = torch.empty((B, K), dtype=torch.float32, device=dev)
        args.d_attr_p = _ptr(out["d_attr_p"])
    L = N.lib()
    ws = _workspace("cdm_score", L.vipsy_cdm_score_workspace_bytes(C.byref(args)), dev)
    args.workspace, args.workspace_bytes = _ptr(ws), ws.numel()
    N.check(L.vipsy_cdm_score_elbo_grad(C.byref(args), _stream()))
    out["loss"] = loss
    out["_keepalive"] = (qm, g_, s_, ap_, u_, idx_, ws)
    return out


SAMPLE_MAX_BATCH = 2048


def sample_indices(n_rows: int, batch: int, seed: int, counter_dev: torch.Tensor) -> torch.Tensor:
    """``randperm(n_rows)[:batch]`` in one stream-ordered launch: ``batch`` distinct row indices, uniform over ordered
    subsets, keyed by (seed, *counter_dev); the kernel increments the device counter (graph-safe)."""
    if not counter_dev.is_cuda or counter_dev.dtype != torch.int64 or counter_dev.numel() != 1:
        raise ValueError("counter_dev must be a one-element CUDA int64 tensor")
    out = torch.empty(batch, dtype=torch.int64, device=counter_dev.device)
    N.check(N.lib().vipsy_sample_indices_dev(_ptr(out), batch, n_rows, seed & (2 ** 64 - 1), _ptr(counter_dev), _stream()))
    return out


# --------------------------------------------------------------------------- NormEncoder (vi.py:417-435, 677-684)
NORM_ENCODER_MAX_HIDDEN = 64


def norm_encoder_pack_tiles(y: torch.Tensor) -> Optional[torch.Tensor]:
    """Resident bf16 operand tiles of the (constant) response matrix for the tensor-core encoder kernels, or None
    when the shape has no tensor-core path (J > 255).  32 KB per 128 persons (64 KB when J > 127); build once."""
    if not y.is_cuda or y.dtype != torch.uint8 or y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be a CUDA uint8 code matrix (see pack_responses)")
    L = N.lib()
    nbytes = L.vipsy_norm_encoder_tiles_bytes(y.shape[0], y.shape[1])
    if nbytes == 0:
        return None
    tiles = torch.empty(nbytes, dtype=torch.uint8, device=y.device)
    N.check(L.vipsy_norm_encoder_pack_tiles(_ptr(y), y.shape[0], y.shape[1], y.stride(0), _ptr(tiles), _stream()))
    return tiles


def _norm_encoder_args(y, idx, enc, tensor_cores=True, x_tiles=None):
    """enc: object with fc1 / fc21 / fc22 ``nn.Linear`` members (fc21, fc22 have one output)."""
    if not y.is_cuda or y.dtype != torch.uint8 or y.dim() != 2 or y.stride(1) != 1:
        raise ValueError("y must be a CUDA uint8 code matrix (see pack_responses)")
    H, J = enc.fc1.weight.shape
    if J != y.shape[1] or enc.fc21.weight.shape != (1, H) or enc.fc22.weight.shape != (1, H):
        raise ValueError("encoder shapes do not match the response matrix / K = 1")
    a = N.NormEncoderArgs()
    a.n_rows, a.n_items, a.n_hidden = y.shape[0], J, H
    a.y, a.y_row_stride = _ptr(y), y.stride(0)
    a.flags = 0 if tensor_cores else 1
    a.x_tiles = _ptr(x_tiles) if (x_tiles is not None and idx is None) else None
    if idx is not None:
        if idx.dtype != torch.int64 or not idx.is_cuda:
            raise ValueError("idx must be a CUDA int64 tensor")
        idx = idx.contiguous()
        a.idx, a.batch = _ptr(idx), idx.numel()
    else:
        a.batch = y.shape[0]
    keep = [idx]
    for name, t in (("w1", enc.fc1.weight), ("b1", enc.fc1.bias), ("w21", enc.fc21.weight), ("b21", enc.fc21.bias),
                    ("w22", enc.fc22.weight), ("b22", enc.fc22.bias)):
        t = _f32(t, name)
        keep.append(t)
        setattr(a, name, _ptr(t))
    return a, keep


def norm_encoder_forward(y: torch.Tensor, enc, idx: Optional[torch.Tensor] = None, tensor_cores: bool = True,
                         x_tiles: Optional[torch.Tensor] = None):
    """mu [B], sigma [B], pre [B, H] of the amortised K = 1 guide, straight from the response codes (or, full
    batch, from the resident operand tiles of ``norm_encoder_pack_tiles``)."""
    a, keep = _norm_encoder_args(y, idx, enc, tensor_cores, x_tiles)
    dev, B, H = y.device, int(a.batch), int(a.n_hidden)
    mu = torch.empty(B, dtype=torch.float32, device=dev)
    sigma = torch.empty(B, dtype=torch.float32, device=dev)
    pre = torch.empty((B, H), dtype=torch.float32, device=dev)
    a.mu, a.sigma, a.pre = _ptr(mu), _ptr(sigma), _ptr(pre)
    N.check(N.lib().vipsy_norm_encoder_forward(C.byref(a), _stream()))
    return mu, sigma, pre


def norm_encoder_backward(y: torch.Tensor, enc, sigma: torch.Tensor, pre: torch.Tensor, d_mu: torch.Tensor,
                          d_sigma: torch.Tensor, idx: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                          tensor_cores: bool = True, x_tiles: Optional[torch.Tensor] = None):
    """Gradients of the six encoder tensors given d loss / d mu and d loss / d sigma; returns a dict keyed by the
    ``named_parameters`` names (fc1.weight, fc1.bias, fc21.weight, ...), shaped like the parameters."""
    a, keep = _norm_encoder_args(y, idx, enc, tensor_cores, x_tiles)
    dev, B, H, J = y.device, int(a.batch), int(a.n_hidden), int(a.n_items)
    sigma, pre = _f32(sigma, "sigma", B), _f32(pre, "pre", B * H)
    d_mu, d_sigma = _f32(d_mu, "d_mu", B), _f32(d_sigma, "d_sigma", B)
    a.sigma, a.pre, a.d_mu, a.d_sigma, a.grad_scale = _ptr(sigma), _ptr(pre), _ptr(d_mu), _ptr(d_sigma), grad_scale
    out = {"fc1.weight": torch.empty((H, J), dtype=torch.float32, device=dev),
           "fc1.bias": torch.empty(H, dtype=torch.float32, device=dev),
           "fc21.weight": torch.empty((1, H), dtype=torch.float32, device=dev),
           "fc21.bias": torch.empty(1, dtype=torch.float32, device=dev),
           "fc22.weight": torch.empty((1, H), dtype=torch.float32, device=dev),
           "fc22.bias": torch.empty(1, dtype=torch.float32, device=dev)}
    a.d_w1, a.d_b1 = _ptr(out["fc1.weight"]), _ptr(out["fc1.bias"])
    a.d_w21, a.d_b21 = _ptr(out["fc21.weight"]), _ptr(out["fc21.bias"])
    a.d_w22, a.d_b22 = _ptr(out["fc22.weight"]), _ptr(out["fc22.bias"])
    L = N.lib()
    ws = _workspace("norm_encoder", L.vipsy_norm_encoder_workspace_bytes(C.byref(a)), dev)
    a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
    N.check(L.vipsy_norm_encoder_backward(C.byref(a), _stream()))
    return out

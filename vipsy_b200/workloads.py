"""Seeded synthetic workloads for the BASELINE.json configurations.

The distributions follow the reference's own generators (``RandomIrt1PL..4PL``
vi.py:202-304, ``RandomDina/RandomHoDina`` vi.py:134-199, ``RandomMilIrt2PL``
vi.py:307-380) but are written for the device: 1e8-cell matrices are drawn on
the GPU in seconds.  Everything returned is a plain dict of tensors.
"""
from __future__ import annotations

import math

import torch


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def uni_irt(n_persons, n_items, model="irt_4pl", missing=0.0, seed=1234, device="cuda", chunk=1 << 18):
    """c2 / c4: unidimensional 1-4PL; responses as uint8 codes (255 = missing)."""
    g = _gen(device, seed)
    J = n_items
    a = torch.rand(1, J, device=device, generator=g) * 2.5 + 0.5          # U(0.5, 3)   vi.py:244-256
    b = torch.randn(1, J, device=device, generator=g)                      # N(0, 1)     vi.py:209-229
    c = torch.rand(1, J, device=device, generator=g) * 0.15 + 0.05         # U(.05, .2)  vi.py:271-279
    d = torch.rand(1, J, device=device, generator=g) * 0.15 + 0.8          # U(.8, .95)  vi.py:291-299
    y = torch.empty((n_persons, J), dtype=torch.uint8, device=device)
    theta = torch.randn(n_persons, 1, device=device, generator=g)
    for s in range(0, n_persons, chunk):
        e = min(s + chunk, n_persons)
        sg = torch.sigmoid(theta[s:e] * (a if model != "irt_1pl" else 1.0) + b)
        lo = c if model in ("irt_3pl", "irt_4pl") else 0.0
        hi = d if model == "irt_4pl" else 1.0
        p = lo + (hi - lo) * sg
        yy = (torch.rand(p.shape, device=device, generator=g) < p).to(torch.uint8)
        if missing > 0:
            yy[torch.rand(p.shape, device=device, generator=g) < missing] = 255
        y[s:e] = yy
    mu = torch.randn(n_persons, 1, device=device, generator=g) * 0.5       # stand-in for encoder output
    sigma = torch.exp(torch.randn(n_persons, 1, device=device, generator=g) * 0.1)
    return dict(y=y, mu=mu, sigma=sigma, a=a, b=b, c=c, d=d, theta=theta, model=model)


def cdm(n_persons, n_items=60, n_attrs=10, seed=4321, device="cuda", hodina=False, missing=0.0, chunk=1 << 18):
    """c5: DINA / HO-DINA, Q ~ Bern(.5) with empty columns repaired (vi.py:148-153)."""
    g = _gen(device, seed)
    K, J = n_attrs, n_items
    q = (torch.rand(K, J, device=device, generator=g) < 0.5).float()
    empty = q.sum(0) == 0
    if empty.any():
        k = torch.randint(0, K, (int(empty.sum()),), device=device, generator=g)
        q[k, empty.nonzero().squeeze(1)] = 1.0
    gs = torch.rand(2, J, device=device, generator=g) * 0.3                # g, s ~ U(0, .3)
    theta = torch.randn(n_persons, 1, device=device, generator=g)
    lam0 = torch.randn(1, K, device=device, generator=g)
    lam1 = torch.rand(1, K, device=device, generator=g) * 2.5 + 0.5
    y = torch.empty((n_persons, J), dtype=torch.uint8, device=device)
    need = q.sum(0)
    for s in range(0, n_persons, chunk):
        e = min(s + chunk, n_persons)
        if hodina:
            pa = torch.sigmoid(theta[s:e] * lam1 + lam0)
        else:
            pa = torch.full((e - s, K), 0.5, device=device)
        attr = (torch.rand(pa.shape, device=device, generator=g) < pa).float()
        eta = (attr @ q == need).float()
        p = (1 - gs[1]) * eta + gs[0] * (1 - eta)
        yy = (torch.rand(p.shape, device=device, generator=g) < p).to(torch.uint8)
        if missing > 0:
            yy[torch.rand(p.shape, device=device, generator=g) < missing] = 255
        y[s:e] = yy
    return dict(y=y, q=q, g_true=gs[0:1], s_true=gs[1:2], lam0_true=lam0, lam1_true=lam1, theta=theta,
                g=torch.full((1, J), 0.1, device=device), s=torch.full((1, J), 0.1, device=device),
                lam0=torch.zeros(1, K, device=device), lam1=torch.ones(1, K, device=device),
                mu=torch.zeros(n_persons, 1, device=device), sigma=torch.ones(n_persons, 1, device=device))


def mirt(n_persons, n_items=500, n_dims=100, seed=2468, device="cuda", chunk=1 << 16):
    """c3: 2PL MIRT, mdisc ~ LogN(0,.5), mdiff ~ N(.5,1), staircase-zeroed loadings
    (vi.py:311-380, simplified direction cosines), per-person packed Cholesky input."""
    g = _gen(device, seed)
    K, J = n_dims, n_items
    mdisc = torch.exp(torch.randn(J, device=device, generator=g) * 0.5)
    mdiff = torch.randn(J, device=device, generator=g) + 0.5
    dirs = torch.rand(K, J, device=device, generator=g)
    dirs = dirs / dirs.norm(dim=0, keepdim=True)
    a = (mdisc * dirs).clamp_min(0.01)
    free = torch.ones(K, J, dtype=torch.bool, device=device)
    for i in range(K):
        free[i, J - i:] = False
    a = a * free
    b = (-mdiff * mdisc).view(1, J)
    theta = torch.randn(n_persons, K, device=device, generator=g)
    y = torch.empty((n_persons, J), dtype=torch.uint8, device=device)
    for s in range(0, n_persons, chunk):
        e = min(s + chunk, n_persons)
        p = torch.sigmoid(theta[s:e] @ a + b)
        y[s:e] = (torch.rand(p.shape, device=device, generator=g) < p).to(torch.uint8)
    # trained-like stand-in for the encoder: mu near theta, small off-diagonals
    mu = theta * 0.8 + 0.2 * torch.randn(n_persons, K, device=device, generator=g)
    tri = K * (K + 1) // 2
    packed = torch.randn(n_persons, tri, device=device, generator=g) * (0.05 / math.sqrt(K))
    rows = torch.arange(K, device=device)
    diag_pos = rows * (rows + 1) // 2 + rows
    packed[:, diag_pos] = -0.7 + 0.1 * torch.randn(n_persons, K, device=device, generator=g)   # raw log-diag
    return dict(y=y, mu=mu, chol_packed_raw=packed, a=a * 0.9 + 0.05 * free, b=b, a_free=free, theta=theta)

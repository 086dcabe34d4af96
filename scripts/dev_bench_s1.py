"""Development timing of the S == 1 row kernels against the lane-per-person kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
from scripts.dev_bench_uni import timeit


def case(N, J, model, missing, sparse, row, csr=False):
    g = torch.Generator(device="cuda").manual_seed(1)
    y = (torch.rand(N, J, device="cuda", generator=g) < 0.6).to(torch.uint8)
    if missing:
        y[torch.rand(N, J, device="cuda", generator=g) < missing] = 255
    mu = torch.randn(N, 1, device="cuda", generator=g) * 0.5
    sg = torch.exp(torch.randn(N, 1, device="cuda", generator=g) * 0.1)
    a = torch.rand(1, J, device="cuda") * 2.5 + 0.5
    b = torch.randn(1, J, device="cuda")
    c = torch.rand(1, J, device="cuda") * 0.15 + 0.05
    d = torch.rand(1, J, device="cuda") * 0.15 + 0.8
    off = [0]
    obs = ops.observed_csr(y) if csr else None

    def run():
        off[0] += 1
        return ops.irt_elbo_grad(y, mu, sigma=sg, n_particles=1, a=a, b=b, c=c, d=d, model=model, seed=1,
                                 offset=off[0], sparse=sparse, row_kernel=row, obs=obs)
    ms = timeit(run, it=30)
    gb = (N * (J + 16) if not csr else 2 * obs[2] + 8 * N + 16 * N) / 1e9
    print(f"N={N} J={J} S=1 {model} missing={missing} sparse={sparse} row={row} csr={csr}: {ms*1e3:.1f} us  "
          f"{N*J/ms*1e3:.3e} cells/s  {gb/ms*1e3:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    case(2_000_000, 200, "irt_2pl", 0.9, True, True, csr=True)
    case(2_000_000, 200, "irt_2pl", 0.5, True, True, csr=True)
    case(2_000_000, 200, "irt_2pl", 0.0, True, True, csr=True)
    case(2_000_000, 200, "irt_4pl", 0.9, True, True, csr=True)
    for row in (True,):
        case(2_000_000, 200, "irt_2pl", 0.9, True, row)
        case(2_000_000, 200, "irt_2pl", 0.9, False, row)
        case(2_000_000, 200, "irt_2pl", 0.5, True, row)
        case(2_000_000, 200, "irt_2pl", 0.0, False, row)
        case(1_000_000, 100, "irt_4pl", 0.0, False, row)
        case(1_000_000, 100, "irt_2pl", 0.0, False, row)
        case(500_000, 500, "irt_2pl", 0.0, False, row)
    for row in (False, True):
        case(2_000_000, 50, "irt_2pl", 0.0, False, row)
        case(2_000_000, 20, "irt_4pl", 0.0, False, row)
        case(4_000_000, 6, "irt_2pl", 0.0, False, row)
        case(1_000_000, 101, "irt_2pl", 0.0, False, row)

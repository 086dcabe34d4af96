"""Development timing of the K=1 kernel (not the contract bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vipsy_b200 import ops
from vipsy_b200.clocks import ClockSampler


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def case(N, J, S, model, missing=0.0, strict=False, inject=False, sparse=False):
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
    eps = torch.randn(S, N, 1, device="cuda") if inject else None
    off = [0]
    n_masked = int((y == 255).sum().item()) if os.environ.get("VIPSY_KNOWN_MASK") else None

    def run():
        off[0] += 1
        return ops.irt_elbo_grad(y, mu, sigma=sg, eps=eps, n_particles=S, a=a, b=b, c=c, d=d, model=model,
                                 seed=1, offset=off[0], strict=strict, sparse=sparse, n_masked=n_masked)
    with ClockSampler() as cs:
        ms = timeit(run, it=30)
    ev = N * J * S
    ck = cs.summary()
    print(f"N={N} J={J} S={S} {model} missing={missing} strict={strict} inject={inject} sparse={sparse}: "
          f"{ms:.3f} ms  {ev / ms * 1e3:.3e} evals/s  clk={ck['sm_mhz']} {ck['reasons']} {ck['power_w_max']}W", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    case(1_000_000, 100, 32, "irt_4pl")
    case(1_000_000, 100, 32, "irt_4pl", strict=True)
    case(1_000_000, 100, 32, "irt_4pl", inject=True)
    case(1_000_000, 100, 32, "irt_2pl")
    case(1_000_000, 100, 32, "irt_3pl")
    case(1_000_000, 100, 32, "irt_1pl")
    case(1_000_000, 100, 8, "irt_4pl")
    case(1_000_000, 100, 1, "irt_4pl")
    case(2_000_000, 200, 1, "irt_2pl", missing=0.9)
    case(2_000_000, 200, 32, "irt_2pl", missing=0.9)
    case(2_000_000, 200, 1, "irt_2pl", missing=0.9, sparse=True)
    case(2_000_000, 200, 32, "irt_2pl", missing=0.9, sparse=True)
    case(2_000_000, 200, 1, "irt_2pl", missing=0.5, sparse=True)
    case(2_000_000, 200, 1, "irt_2pl", missing=0.5)
    case(1000, 5, 1, "irt_2pl")

"""A/B of the K=1 particle kernel: pipelined item loop (default) vs the branchy one (VIPSY_UNI_NOPIPE=1).
Runs itself twice as subprocesses, compares every output bit for bit, prints both timings."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [(1_000_000, 100, 32, "irt_4pl", 0.0, False), (1_000_000, 100, 32, "irt_4pl", 0.0, True),
         (1_000_000, 100, 32, "irt_2pl", 0.0, False), (1_000_000, 100, 32, "irt_3pl", 0.0, False),
         (1_000_000, 100, 32, "irt_1pl", 0.0, False), (1_000_000, 100, 8, "irt_4pl", 0.0, False),
         (1_000_000, 100, 4, "irt_4pl", 0.3, False), (300_001, 37, 2, "irt_4pl", 0.2, False)]


def child(tag):
    import torch
    from vipsy_b200 import ops
    from scripts.dev_bench_uni import timeit
    out = {}
    for ci, (N, J, S, model, missing, strict) in enumerate(CASES):
        g = torch.Generator(device="cuda").manual_seed(1)
        y = (torch.rand(N, J, device="cuda", generator=g) < 0.6).to(torch.uint8)
        if missing:
            y[torch.rand(N, J, device="cuda", generator=g) < missing] = 255
        mu = torch.randn(N, 1, device="cuda", generator=g) * 0.5
        sg = torch.exp(torch.randn(N, 1, device="cuda", generator=g) * 0.1)
        a = torch.rand(1, J, device="cuda", generator=g) * 2.5 + 0.5
        b = torch.randn(1, J, device="cuda", generator=g)
        c = torch.rand(1, J, device="cuda", generator=g) * 0.15 + 0.05
        d = torch.rand(1, J, device="cuda", generator=g) * 0.15 + 0.8
        if ci == 6:                       # a few persons far out: the exponent bound must fail for their tiles
            mu[::1000] = 40.0
        run = lambda: ops.irt_elbo_grad(y, mu, sigma=sg, n_particles=S, a=a, b=b, c=c, d=d, model=model, seed=1,
                                        offset=7, strict=strict)
        r = run()
        for k, v in r.items() if isinstance(r, dict) else enumerate(r):
            if torch.is_tensor(v):
                out[f"{ci}/{k}"] = v.detach().cpu()
        ms = timeit(run, it=20)
        print(f"[{tag}] N={N} J={J} S={S} {model} missing={missing} strict={strict}: {ms:.3f} ms "
              f"{N * J * S / ms * 1e3:.3e} evals/s", flush=True)
    torch.save(out, f"/tmp/ab_{tag}.pt")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    import torch
    for tag, env, unr in (("old", "1", "1"), ("pipe", "0", "1"), ("pipe2", "0", "2"), ("defer", "0", "0")):
        e = dict(os.environ, VIPSY_UNI_NOPIPE=env, VIPSY_UNI_UNROLL=unr)
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=e, check=True)
    for other in ("pipe", "pipe2", "defer"):
        a, b = torch.load("/tmp/ab_old.pt"), torch.load(f"/tmp/ab_{other}.pt")
        bad = 0
        for k in a:
            same = torch.equal(torch.nan_to_num(a[k]), torch.nan_to_num(b[k]))
            if not same:
                bad += 1
                print("DIFF", other, k, (a[k].double() - b[k].double()).abs().max().item(), a[k].abs().max().item())
        print(f"{other}: compared {len(a)} tensors, {bad} differ")

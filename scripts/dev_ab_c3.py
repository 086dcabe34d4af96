"""A/B of the K = 100 guide-side kernels: bulk-copy fed (default) vs register-fed (VIPSY_NO_BULK=1)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(tag):
    import torch
    from vipsy_b200 import ops, workloads
    from scripts.dev_bench_uni import timeit
    out = {}
    for ci, (N, K, J, S) in enumerate(((200_000, 100, 500, 1), (100_001, 100, 60, 2), (200_000, 64, 40, 1), (200_003, 50, 40, 1), (100_000, 128, 40, 1), (100_000, 96, 40, 1), (200_000, 40, 40, 1))):
        w = workloads.mirt(N, J, K)
        dchol = torch.zeros_like(w["chol_packed_raw"])
        dmu = torch.empty_like(w["mu"])
        run = lambda: ops.irt_elbo_grad(w["y"], w["mu"], chol=w["chol_packed_raw"], chol_packed=True, chol_raw=True,
                                        n_particles=S, a=w["a"], b=w["b"], model="irt_2pl", a_free=w["a_free"],
                                        seed=1, offset=3, d_mu_out=dmu, d_scale_out=dchol)
        r = run()
        torch.cuda.synchronize()
        out[f"{ci}/loss"] = r["loss"].detach().cpu().clone()
        out[f"{ci}/d_mu"] = dmu.cpu().clone()
        out[f"{ci}/d_chol"] = dchol.cpu().clone()
        out[f"{ci}/d_a"] = r["d_a"].cpu().clone()
        ms = timeit(run, it=10)
        print(f"[{tag}] N={N} K={K} J={J} S={S}: {ms:.3f} ms", flush=True)
    torch.save(out, f"/tmp/abc3_{tag}.pt")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    import torch
    for tag, env in (("old", "1"), ("bulk", "0")):
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=dict(os.environ, VIPSY_NO_BULK=env), check=True)
    a, b = torch.load("/tmp/abc3_old.pt"), torch.load("/tmp/abc3_bulk.pt")
    for k in a:
        d = (a[k].double() - b[k].double()).abs()
        sc = a[k].double().abs().max().item()
        print(k, "max abs diff", d.max().item(), "scale", sc, "rel", d.max().item() / max(sc, 1e-30),
              "n_bad", int((d > 1e-4 * sc).sum()), "first_bad", (d.flatten() > 1e-4 * sc).nonzero()[:4].flatten().tolist())

"""Generate tests/golden/golden_models_v1.npz: ELBO-level golden vectors produced by the reference's OWN
``model`` / ``guide`` methods (``/root/reference/vi.py``, imported unmodified) executed under
``tests/golden/minipyro.py``, the stand-in for the absent ``pyro-ppl==1.4.0``.

Run ONLY in the builder container (needs /root/reference):   python tests/golden/make_golden_models.py

For every estimator class one ``loss_and_grads`` call is made exactly the way the patched ``SVI.step``
(vi.py:503-516) makes it; recorded are the data, the parameter values, the plate subsample, the base noise of
every draw, the loss and the gradient of every unconstrained parameter (``a`` after the ``free`` mask of
vi.py:511-512).  ``tests/test_oracle_golden.py`` feeds the same inputs to the oracle and must reproduce loss
and gradients; the CUDA kernels are then checked against the oracle (``tests/test_parity_*.py``).
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minipyro  # noqa: E402

REF = "/root/reference"


def load_reference():
    minipyro.install()
    spec = importlib.util.spec_from_file_location("vi_ref_models", os.path.join(REF, "vi.py"))
    vi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vi)
    return vi


def responses(g, N, J, missing):
    y = (torch.rand(N, J, generator=g) < 0.55).float()
    if missing:
        y[torch.rand(N, J, generator=g) < missing] = float("nan")
    return y


def perturb(g, scale=0.3):
    """Move every stored parameter off its init so that the golden case is not a symmetric point."""
    with torch.no_grad():
        for name, (u, _) in minipyro.get_param_store().items():
            u.add_(torch.randn(u.shape, generator=g) * scale)
            if name == "a":
                pass


def run_case(out, key, obj, elbo, data, g, n_warm_perturb=True):
    """One loss_and_grads on fresh parameters: a first call creates the parameter sites (as the first SVI.step
    would), they are perturbed, gradients are cleared, and the recorded call follows."""
    minipyro.clear_param_store()
    torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
    elbo.loss_and_grads(obj.model, obj.guide, data)
    if n_warm_perturb:
        perturb(g)
    for u, _ in minipyro.get_param_store().values():
        u.grad = None
    loss = elbo.loss_and_grads(obj.model, obj.guide, data)
    out[f"{key}/loss"] = np.float64(loss)
    out[f"{key}/data"] = data.numpy()
    out[f"{key}/S"] = np.int64(elbo.num_particles)
    for name, (u, c) in minipyro.get_param_store().items():
        out[f"{key}/param_u/{name}"] = u.detach().numpy().copy()
        gr = u.grad.detach().clone() if u.grad is not None else torch.zeros_like(u)
        if name == "a" and getattr(obj, "a_free", None) is not None:
            gr = gr * obj.a_free                                        # vi.py:511-512
            out[f"{key}/a_free"] = obj.a_free.numpy().astype(np.uint8)
        out[f"{key}/grad_u/{name}"] = gr.numpy()
    for s, rec in enumerate(elbo.records):
        for pn, idx in rec["idx"].items():
            out[f"{key}/idx/{s}"] = idx.numpy()
        for sn, noise in rec["noise"].items():
            out[f"{key}/noise/{sn}/{s}"] = noise.numpy()
    print(f"{key}: loss {loss:.6f}  params {list(minipyro.get_param_store())}")


def main():
    vi = load_reference()
    E, EE = minipyro.Trace_ELBO, minipyro.TraceEnum_ELBO
    g = torch.Generator().manual_seed(20261017)
    out = {}
    # ---- IRT, K = 1 (vi.py:588-595, 701-705) ---------------------------------------------------------------
    for model in ("irt_1pl", "irt_2pl", "irt_3pl", "irt_4pl"):
        y = responses(g, 40, 7, 0.12)
        run_case(out, f"virt_{model}_k1_S3", vi.VIRT(data=y, model=model, D=1.702 if model == "irt_3pl" else 1), E(3), y, g)
    y = responses(g, 50, 6, 0.1)
    run_case(out, "virt_irt_2pl_k1_sub15", vi.VIRT(data=y, model="irt_2pl", subsample_size=15), E(1), y, g)
    # ---- IRT, K = 3 (vi.py:596-615, 706-723): shared and per-person covariance, staircase free mask -----------
    y = responses(g, 30, 9, 0.1)
    run_case(out, "virt_irt_2pl_k3_person_S2", vi.VIRT(data=y, model="irt_2pl", x_feature=3), E(2), y, g)
    y = responses(g, 30, 9, 0.0)
    run_case(out, "virt_irt_4pl_k3_shared_S2", vi.VIRT(data=y, model="irt_4pl", x_feature=3, share_cov=True), E(2), y, g)
    # ---- amortised IRT (vi.py:673-693) ------------------------------------------------------------------------
    y = responses(g, 35, 8, 0.15)
    torch.manual_seed(11)
    run_case(out, "vaeirt_irt_3pl_k1_S2", vi.VaeIRT(data=y, model="irt_3pl", hidden_dim=5), E(2), y, g)
    y = responses(g, 24, 10, 0.1)
    torch.manual_seed(12)
    run_case(out, "vaeirt_irt_2pl_k3_sub12", vi.VaeIRT(data=y, model="irt_2pl", x_feature=3, hidden_dim=5, subsample_size=12), E(1), y, g)
    # ---- REINFORCE CDMs (vi.py:745-756, 796-816) ----------------------------------------------------------------
    K, J = 3, 8
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    q[0, q.sum(0) == 0] = 1.0
    for model in ("dina", "dino"):
        y = responses(g, 30, J, 0.0)
        run_case(out, f"vcdm_{model}_S2", vi.VCDM(data=y, q=q, model=model), E(2), y, g)
    out["cdm/q"] = q.numpy()
    y = responses(g, 30, J, 0.0)
    torch.manual_seed(13)
    run_case(out, "vaecdm_dina_sub10", vi.VaeCDM(data=y, q=q, model="dina", hidden_dim=5, subsample_size=10), E(1), y, g)
    # ---- enumerated CDMs (vi.py:839-857, 872-891, 897-934, 968-981) ---------------------------------------------
    for model in ("dina", "dino"):
        y = responses(g, 25, J, 0.15)
        run_case(out, f"vccdm_{model}", vi.VCCDM(data=y, q=q, model=model), EE(1), y, g)
    y = responses(g, 40, J, 0.1)
    run_case(out, "vccdm_dina_sub16", vi.VCCDM(data=y, q=q, model="dina", subsample_size=16), EE(1), y, g)
    y = responses(g, 20, J, 0.15)
    torch.manual_seed(14)
    run_case(out, "vaeccdm_dina", vi.VaeCCDM(data=y, q=q, model="dina", hidden_dim=5), EE(1), y, g)
    y = responses(g, 25, J, 0.15)
    run_case(out, "vchodina_S2", vi.VCHoDina(data=y, q=q), EE(2), y, g)
    y = responses(g, 25, J, 0.15)
    torch.manual_seed(15)
    run_case(out, "vaechodina_sub9", vi.VaeCHoDina(data=y, q=q, hidden_dim=5, subsample_size=9), EE(1), y, g)
    path = os.path.join(HERE, "golden_models_v1.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()

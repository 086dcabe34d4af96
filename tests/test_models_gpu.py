"""GPU: the estimator classes (reference constructor/fit API) end to end."""
import numpy as np
import pytest
import torch

import vipsy_b200 as vb
from oracle import vipsy_oracle as o
from tests.util import assert_close
from vipsy_b200 import infer

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fresh_store():
    from vipsy_b200 import build
    build.build()
    vb.clear_param_store()
    yield
    vb.clear_param_store()


def lsat(golden):
    return torch.from_numpy(golden["lsat_y"]).float()


def test_virt_step_matches_oracle_through_the_class_api(golden):
    """VIRT(data=y, model='irt_2pl') first loss_and_grads == oracle on the kernel's own noise."""
    from vipsy_b200 import ops
    y = lsat(golden)
    m = vb.VIRT(data=y, model="irt_2pl", seed=5)
    cfg = vb.Trace_ELBO(num_particles=4)
    loss, sites = m.loss_and_grads(cfg, y)
    eps = ops.fill_normal(4, 1000, 1, seed=5, offset=1).cpu()
    ref = o.irt_elbo_grad(y, torch.zeros(1000, 1), torch.ones(1000, 1), eps, torch.ones(1, 5), torch.zeros(1, 5),
                          model="irt_2pl")
    g = {n: p.grad for n, p, _ in sites}
    assert_close(loss, ref["loss"], 1e-5, 1e-5, "loss")
    assert_close(g["a"], ref["d_a"], 1e-5, 1e-5, "a")
    assert_close(g["b"], ref["d_b"], 1e-5, 1e-5, "b")
    assert_close(g["x_local"], ref["d_mu"], 1e-5, 1e-5, "x_local")
    assert_close(g["x_scale"], ref["d_scale"] * 1.0, 1e-5, 1e-5, "x_scale (log-sigma, sigma=1)")


def test_lsat_2pl_fit_recovers_literature_estimates(golden):
    """BASELINE config 1 end to end (SURVEY 8c anchor (i)): b -> ~[2.7,.95,.2,1.2,2.0]."""
    y = lsat(golden)
    torch.manual_seed(0)
    m = vb.VIRT(data=y, model="irt_2pl")
    m.fit(optim=vb.Adam({"lr": 5e-2}), loss=vb.Trace_ELBO(num_particles=1), max_iter=1500)
    b = vb.param("b").cpu().reshape(-1)
    a = vb.param("a").cpu().reshape(-1)
    assert torch.allclose(b, torch.tensor([2.72, 0.95, 0.21, 1.24, 2.01]), atol=0.35), b
    assert (a > 0.3).all() and (a < 1.3).all(), a


def test_vaeirt_encoder_receives_kernel_gradients():
    from vipsy_b200 import ops
    torch.manual_seed(1)
    N, J = 400, 12
    y = (torch.rand(N, J) < 0.6).float()
    y[torch.rand(N, J) < 0.1] = float("nan")
    m = vb.VaeIRT(data=y, model="irt_3pl", hidden_dim=8, seed=3)
    loss, sites = m.loss_and_grads(vb.Trace_ELBO(num_particles=2), y)
    enc = m.encoder
    # CPU restatement: same encoder weights, same noise
    import copy
    enc_cpu = copy.deepcopy(enc).cpu()
    for p in enc_cpu.parameters():
        p.grad = None
    yin = torch.where(torch.isnan(y), torch.full_like(y, -1.0), y)
    loc, sc = enc_cpu(yin)
    eps = ops.fill_normal(2, N, 1, seed=3, offset=1).cpu()
    ref = o.irt_elbo_grad(y, loc.detach(), sc.detach(), eps, torch.ones(1, J), torch.zeros(1, J),
                          torch.full((1, J), 0.1), None, model="irt_3pl")
    torch.autograd.backward([loc, sc], [ref["d_mu"].float(), ref["d_scale"].float()])
    assert_close(loss, ref["loss"], 2e-5, 2e-5, "loss")
    for (n, p), pc in zip(enc.named_parameters(), enc_cpu.parameters()):
        assert_close(p.grad, pc.grad, 1e-3, 2e-4, n)      # fp32 GEMM on device vs CPU
    names = [n for n, _, _ in sites]
    assert "encoder$$$fc1.weight" in names and "c" in names and "a" in names


@pytest.mark.parametrize("share", [False, True])
def test_virt_mirt_unconstrained_gradients(share):
    from vipsy_b200 import ops
    torch.manual_seed(2)
    N, J, K = 90, 14, 3
    y = (torch.rand(N, J) < 0.5).float()
    m = vb.VIRT(data=y, model="irt_2pl", x_feature=K, share_cov=share, seed=9)
    st = infer.get_param_store()
    loss, sites = m.loss_and_grads(vb.Trace_ELBO(num_particles=2), y)
    eps = ops.fill_normal(2, N, K, seed=9, offset=1).cpu()
    raw = (torch.zeros(K, K) if share else torch.zeros(N, K, K))       # init: L = I  -> raw = 0
    ref = o.irt_elbo_grad(y, torch.zeros(N, K), raw, eps, m.a0, torch.zeros(1, J), model="irt_2pl",
                          a_free=m.a_free, chol_raw=True)
    g = {n: p.grad for n, p, _ in sites}
    assert_close(loss, ref["loss"], 2e-5, 2e-5)
    assert_close(g["x_local"], ref["d_mu"], 2e-5, 2e-5)
    assert_close(g["x_scale"].cpu().tril(), ref["d_scale"].tril(), 2e-5, 2e-5)
    assert_close(g["a"], ref["d_a"], 2e-5, 2e-5)
    assert st.constraint("x_scale") == infer.LOWER_CHOLESKY


def test_subsampling_scales_and_scatters():
    torch.manual_seed(3)
    N, J = 500, 8
    y = (torch.rand(N, J) < 0.5).float()
    m = vb.VIRT(data=y, model="irt_2pl", subsample_size=100)
    loss, sites = m.loss_and_grads(vb.Trace_ELBO(num_particles=3), y)
    g = {n: p.grad for n, p, _ in sites}
    touched = int((g["x_local"].abs().sum(1) != 0).sum())
    assert 100 <= touched <= 300                         # 3 particles x 100 persons, fresh idx each
    assert 0.3 * N * J < float(loss) < 3.0 * N * J       # scaled by N/B: same order as full batch


def test_vccdm_fit_recovers_guess_and_slip():
    g0 = torch.Generator().manual_seed(4)
    N, J, K = 3000, 30, 4
    q = (torch.rand(K, J, generator=g0) < 0.5).float()
    q[:, q.sum(0) == 0] = torch.eye(K)[:, :1]
    attr = (torch.rand(N, K, generator=g0) < 0.5).float()
    gt, st_ = torch.rand(1, J, generator=g0) * 0.25 + 0.02, torch.rand(1, J, generator=g0) * 0.25 + 0.02
    p = o.cdm_prob("dina", attr, q, gt, st_)
    y = (torch.rand(N, J, generator=g0) < p).float()
    m = vb.VCCDM(data=y, q=q, model="dina")
    m.fit(optim=vb.Adam({"lr": 2e-2}), max_iter=600)
    assert float((vb.param("g").cpu() - gt).abs().mean()) < 0.04
    assert float((vb.param("s").cpu() - st_).abs().mean()) < 0.06
    post = m.posterior()
    assert post.shape == (N, 2 ** K) and torch.allclose(post.sum(1), torch.ones(N, device=post.device), atol=1e-4)
    # MAP pattern recovers most attribute vectors
    map_c = post.argmax(1).cpu()
    true_c = (attr * (2 ** torch.arange(K))).sum(1).long()
    assert float((map_c == true_c).float().mean()) > 0.5


def test_vchodina_and_vcdm_steps_run_and_match():
    from vipsy_b200 import ops
    torch.manual_seed(5)
    N, J, K = 200, 20, 5
    q = (torch.rand(K, J) < 0.5).float()
    y = (torch.rand(N, J) < 0.5).float()
    m = vb.VCHoDina(data=y, q=q, seed=2)
    loss, sites = m.loss_and_grads(vb.TraceEnum_ELBO(num_particles=2), y)
    eps = ops.fill_normal(2, N, 1, seed=2, offset=1).cpu()
    ref = o.hodina_elbo_grad(y, q, torch.full((1, J), .1), torch.full((1, J), .1), torch.zeros(1, K),
                             torch.ones(1, K), torch.zeros(N, 1), torch.ones(N, 1), eps)
    g = {n: p.grad for n, p, _ in sites}
    assert_close(loss, ref["loss"], 2e-5, 2e-5)
    assert_close(g["lam0"], ref["d_lam0"], 2e-5, 2e-5)
    assert_close(g["lam1"], ref["d_lam1"] * 1.0, 2e-5, 2e-5)          # exp chain at lam1 = 1
    assert_close(g["g"], ref["d_g"] * 0.1 * 0.9, 2e-5, 2e-5)          # sigmoid chain at 0.1
    assert_close(g["theta_local"], ref["d_mu"], 2e-5, 2e-5)
    vb.clear_param_store()
    m2 = vb.VCDM(data=y, q=q, model="dino")
    l0 = m2.fit(max_iter=3)
    assert np.isfinite(l0)



def test_host_step_two_bit_transport_matches_device_call():
    """IrtHostStep with the response matrix in the 2-bit host transport format == the same step on codes
    already resident in HBM (same Philox stream): bit-identical outputs, 4x fewer response bytes over PCIe."""
    from vipsy_b200 import ops, workloads
    from vipsy_b200.host_api import IrtHostStep
    N, J = 5000, 100
    w = workloads.uni_irt(N, J, "irt_4pl", missing=0.1, seed=5)
    ref = ops.irt_elbo_grad(w["y"], w["mu"], sigma=w["sigma"], n_particles=8, a=w["a"], b=w["b"], c=w["c"], d=w["d"],
                            model="irt_4pl", seed=9, offset=3)
    host = IrtHostStep(N, J, "irt_4pl", n_chunks=1)
    hp = {k: w[k].cpu() for k in ("a", "b", "c", "d", "mu", "sigma")}
    hy = w["y"].cpu()
    for y_in in (hy, ops.pack_responses_2bit_host(hy)):
        got = host(y_in, hp["mu"], hp["sigma"], hp["a"], hp["b"], hp["c"], hp["d"], n_particles=8, seed=9, offset=3)
        got = {k: v.clone() for k, v in got.items()}
        assert torch.equal(got["d_mu"], ref["d_mu"].cpu()) and torch.equal(got["d_sigma"], ref["d_sigma"].cpu())
        assert torch.equal(got["item_grads"], ref["item_grads"].cpu())
        assert float(got["loss"]) == float(ref["loss"])
    assert host.h2d_bytes == N * (25 + 8) + 4 * J * 4
    host4 = IrtHostStep(N, J, "irt_4pl", n_chunks=4)          # chunked: per-person outputs identical, item sums to rounding
    got4 = host4(ops.pack_responses_2bit_host(hy), hp["mu"], hp["sigma"], hp["a"], hp["b"], hp["c"], hp["d"], n_particles=8, seed=9, offset=3)
    assert torch.equal(got4["d_mu"], ref["d_mu"].cpu())
    assert_close(got4["item_grads"], ref["item_grads"], 1e-5, 1e-5, "item grads over 4 chunks")
    assert_close(got4["loss"], ref["loss"], 1e-9, 1e-9, "loss over 4 chunks")
    assert torch.equal(ops.unpack_responses_2bit(ops.pack_responses_2bit_host(hy).cuda(), J).cpu(), hy)
    # the `free` mask of site 'a' (vi.py:511-512) travels with the host call: frozen entries come back exactly 0
    free = torch.rand(J) < 0.5
    gotf = host(hy, hp["mu"], hp["sigma"], hp["a"], hp["b"], hp["c"], hp["d"], n_particles=8, seed=9, offset=3, a_free=free)
    ig, igf = got["item_grads"].clone().reshape(4, J), gotf["item_grads"].reshape(4, J)
    assert torch.equal(igf[1], ig[1] * free) and torch.equal(igf[[0, 2, 3]], ig[[0, 2, 3]])


# ---- device-resident iteration: fused Adam, device RNG counters, CUDA graph ----------------
def test_fused_adam_matches_torch_adam():
    from vipsy_b200 import ops
    torch.manual_seed(0)
    shapes = [(7, 13), (1, 5), (1000, 1)]
    ps = [torch.randn(s, device="cuda") for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opts = [torch.optim.Adam([r], lr=lr) for r, lr in zip(ref, (1e-2, 5e-2, 1e-3))]
    free = (torch.rand(shapes[0], device="cuda") < 0.7)
    st = [dict(param=p, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p), lr=lr, beta1=0.9, beta2=0.999,
               eps=1e-8, step_offset=0) for p, lr in zip(ps, (1e-2, 5e-2, 1e-3))]
    st[0]["free_mask"] = free.to(torch.uint8)
    st[1]["lr_dev"] = torch.full((1,), 5e-2, device="cuda")          # device learning rate overrides lr
    st[1]["lr"] = 123.0
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for it in range(6):
        gs = [torch.randn(s, device="cuda") * (it + 1) for s in shapes]
        for r, g, o_ in zip(ref, gs, opts):
            r.grad = (g * free if r.shape == free.shape else g).clone()
            o_.step()
        for t, g in zip(st, gs):
            t["grad"] = g.clone()
        step.add_(1)
        ops.adam_step(st, step)
    for p, r in zip(ps, ref):
        assert_close(p, r.detach(), 2e-6, 2e-6, "param")


def test_fill_normal_device_counter_is_the_same_stream():
    from vipsy_b200 import ops
    ctr = torch.tensor([41], dtype=torch.int64, device="cuda")
    a = ops.fill_normal(3, 50, 4, seed=9, offset=41)
    b = ops.fill_normal(3, 50, 4, seed=9, offset=ctr)
    assert torch.equal(a, b)
    ctr.add_(1)
    assert torch.equal(ops.fill_uniform(2, 10, 3, seed=9, offset=42), ops.fill_uniform(2, 10, 3, seed=9, offset=ctr))


def _fit_pair(make, steps, **fit_kw):
    """The same fit twice -- eager torch.optim objects vs the captured device-resident iteration."""
    out = []
    for graph in (False, True):
        vb.clear_param_store()
        torch.manual_seed(3)
        m = make()
        m._device_rng = True            # both runs draw noise through the device counter: identical streams
        losses = []
        svi = infer.SVI(m.model, m.guide, optim=fit_kw["optim"](), loss=fit_kw["loss"], cuda_graph=graph)
        for _ in range(steps):
            losses.append(svi.step(m.data))
        out.append((losses, {n: vb.get_param_store().get(n).clone() for n in vb.get_param_store().names()}))
    return out


def test_cuda_graph_step_equals_eager_virt_4pl():
    g = torch.Generator().manual_seed(0)
    y = (torch.rand(600, 24, generator=g) < 0.6).float()
    y[torch.rand(600, 24, generator=g) < 0.1] = float("nan")
    (le, pe), (lg, pg) = _fit_pair(lambda: vb.VIRT(data=y, model="irt_4pl", seed=4), 12,
                                   optim=lambda: vb.Adam({"lr": 2e-2}), loss=vb.Trace_ELBO(num_particles=2))
    assert np.allclose(le, lg, rtol=2e-5), (le, lg)
    for n in pe:
        assert_close(pg[n], pe[n], 1e-4, 1e-4, n)
    assert lg[-1] < lg[0]


def test_cuda_graph_step_equals_eager_minibatch_vaeirt():
    """Subsampled fit: both modes draw the minibatch with the same kernel and device counter (and the same noise
    streams), so the captured iteration reproduces the eager one loss by loss."""
    g = torch.Generator().manual_seed(1)
    y = (torch.rand(900, 20, generator=g) < 0.6).float()
    y[torch.rand(900, 20, generator=g) < 0.1] = float("nan")
    (le, pe), (lg, pg) = _fit_pair(lambda: vb.VaeIRT(data=y, model="irt_2pl", subsample_size=64, hidden_dim=16, seed=2), 10,
                                   optim=lambda: vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(num_particles=2))
    assert np.allclose(le, lg, rtol=2e-5), (le, lg)
    for n in pe:
        assert_close(pg[n], pe[n], 1e-4, 1e-4, n)


def test_cuda_graph_step_equals_eager_vaeirt_mirt_with_scheduler():
    """Encoder forward + backward inside the captured graph, staircase free mask, MultiStepLR through lr_dev."""
    g = torch.Generator().manual_seed(1)
    y = (torch.rand(300, 40, generator=g) < 0.5).float()
    sched = lambda: vb.MultiStepLR({"optimizer": torch.optim.Adam, "optim_args": {"lr": 1e-2}, "milestones": [4, 8],
                                    "gamma": 0.5})
    out = []
    for graph in (False, True):
        vb.clear_param_store()
        torch.manual_seed(3)
        m = vb.VaeIRT(data=y, model="irt_2pl", x_feature=3, hidden_dim=8, seed=4)
        m._device_rng = True
        opt = sched()
        svi = infer.SVI(m.model, m.guide, optim=opt, loss=vb.Trace_ELBO(num_particles=1), cuda_graph=graph)
        losses = []
        for _ in range(12):
            losses.append(svi.step(m.data))
            opt.step()
        out.append((losses, vb.param("a").clone(), [p.detach().clone() for p in m.encoder.parameters()]))
    (le, ae, ee), (lg, ag, eg) = out
    assert np.allclose(le, lg, rtol=1e-4), (le, lg)
    assert_close(ag, ae, 1e-3, 1e-3, "a")
    assert float(ag[2, -1]) == 0.0 and float(ag[1, -1]) == 0.0          # frozen staircase entries never move
    for x, r in zip(eg, ee):
        assert_close(x, r, 2e-3, 2e-3, "encoder")


def test_cuda_graph_minibatch_fit_lsat(golden):
    """The reference's default usage shape: subsample_size < N, many cheap iterations, one graph launch each;
    fresh minibatch and noise on every replay (the fit converges to the literature estimates)."""
    y = lsat(golden)
    torch.manual_seed(0)
    m = vb.VIRT(data=y, model="irt_2pl", subsample_size=250)
    m.fit(optim=vb.Adam({"lr": 5e-2}), loss=vb.Trace_ELBO(num_particles=1), max_iter=3000, cuda_graph=True)
    b = vb.param("b").cpu().reshape(-1)
    assert torch.allclose(b, torch.tensor([2.72, 0.95, 0.21, 1.24, 2.01]), atol=0.45), b


@pytest.mark.parametrize("cls", ["VCDM", "VCCDM", "VCHoDina", "VaeCHoDina"])
def test_cuda_graph_cdm_families(cls):
    g = torch.Generator().manual_seed(2)
    K, J, N = 3, 20, 400
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    q[0, q.sum(0) == 0] = 1.0
    y = (torch.rand(N, J, generator=g) < 0.5).float()
    res = []
    for graph in (False, True):
        vb.clear_param_store()
        torch.manual_seed(3)
        m = getattr(vb, cls)(data=y, q=q, seed=2, **({"hidden_dim": 8} if cls.startswith("Vae") else {}))
        m._device_rng = True
        svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-2}), loss=m._default_loss() if hasattr(m, "_default_loss") else vb.TraceEnum_ELBO(), cuda_graph=graph)
        res.append([svi.step(m.data) for _ in range(8)])
    assert np.allclose(res[0], res[1], rtol=1e-4), res


def test_vaeccdm_class_matches_oracle_and_trains():
    """VaeCCDM end to end: the encoder (softmax over the batch axis, sic) receives -scale * posterior through
    log Categorical(attr_p).logits; first-step loss and g/s gradients equal the oracle's."""
    import copy
    g = torch.Generator().manual_seed(5)
    K, J, N = 3, 18, 120
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    q[0, q.sum(0) == 0] = 1.0
    y = (torch.rand(N, J, generator=g) < 0.5).float()
    y[torch.rand(N, J, generator=g) < 0.1] = float("nan")
    torch.manual_seed(1)
    m = vb.VaeCCDM(data=y, q=q, hidden_dim=8)
    enc_cpu = copy.deepcopy(m.encoder).cpu()
    loss, sites = m.loss_and_grads(vb.TraceEnum_ELBO(num_particles=1), y)
    yin = torch.where(torch.isnan(y), torch.full_like(y, -1.0), y)
    attr_p = enc_cpu(yin)
    ref = o.cdm_enum_elbo_grad(y, q, torch.full((1, J), 0.1), torch.full((1, J), 0.1), model="dina",
                               prior_probs=attr_p.detach(), mask_missing=False)
    assert_close(loss, ref["loss"], 2e-5, 2e-5, "loss")
    gr = {n: p.grad for n, p, _ in sites}
    assert_close(gr["g"], ref["d_g"] * 0.1 * 0.9, 1e-4, 1e-4, "g (unconstrained)")
    attr_p.backward(ref["d_prior"].float())
    # fc2.bias has an exactly-zero gradient (a per-pattern shift cancels in the softmax over the batch axis):
    # compare every tensor on the scale of the largest encoder gradient
    gmax = max(float(pc.grad.abs().max()) for pc in enc_cpu.parameters())
    for (n, p), pc in zip(m.encoder.named_parameters(), enc_cpu.parameters()):
        assert float((p.grad.cpu() - pc.grad).abs().max()) <= 2e-3 * gmax, n
    vb.clear_param_store()
    torch.manual_seed(1)
    m = vb.VaeCCDM(data=y, q=q, hidden_dim=8)
    first = m.fit(optim=vb.Adam({"lr": 1e-2}), max_iter=1)
    last = m.fit(optim=vb.Adam({"lr": 1e-2}), max_iter=150)
    assert last < first


def test_cuda_graph_minibatch_multi_particle_and_evaluate_loss():
    """Captured iteration with a fresh minibatch per particle (several kernel calls per step) and the reference's
    evaluate_loss-after-every-step option (vi.py:770) running eagerly beside the graph."""
    g = torch.Generator().manual_seed(4)
    K, J, N = 3, 16, 300
    q = (torch.rand(K, J, generator=g) < 0.5).float()
    q[0, q.sum(0) == 0] = 1.0
    y = (torch.rand(N, J, generator=g) < 0.5).float()
    torch.manual_seed(0)
    m = vb.VIRT(data=y, model="irt_2pl", subsample_size=64)
    l = m.fit(optim=vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(num_particles=3), max_iter=40, cuda_graph=True)
    assert np.isfinite(l)
    vb.clear_param_store()
    m2 = vb.VCCDM(data=y, q=q, subsample_size=100)
    l2 = m2.fit(optim=vb.Adam({"lr": 1e-2}), max_iter=30, evaluate_loss_each_iter=True, cuda_graph=True)
    assert np.isfinite(l2)
    vb.clear_param_store()
    m3 = vb.VaeIRT(data=y, model="irt_2pl", x_feature=2, subsample_size=50, hidden_dim=8)      # fused K = 2 kernel in the graph
    l3 = m3.fit(optim=vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(num_particles=2), max_iter=30)
    assert np.isfinite(l3)


def test_default_fit_falls_back_to_eager_for_adam_options_the_device_step_does_not_cover():
    """fit() captures a CUDA graph by default; a valid Adam configuration outside the device-resident step's
    coverage (weight decay from a per-parameter callable, gradient clipping) must run through the eager
    torch.optim path and give the same result as an explicit cuda_graph=False, not fail on the first step."""
    g = torch.Generator().manual_seed(1)
    y = (torch.rand(400, 12, generator=g) < 0.6).float()

    def per_param(module_name, param_name):
        return {"lr": 2e-2, "weight_decay": 1e-2} if param_name == "a" else {"lr": 2e-2}
    res = []
    for graph in (None, False):
        vb.clear_param_store()
        m = vb.VIRT(data=y, model="irt_2pl", seed=2)
        m._device_rng = True
        loss = m.fit(optim=vb.Adam(per_param, clip_args={"clip_norm": 50.0}), max_iter=6, cuda_graph=graph)
        res.append((loss, vb.get_param_store().get("a").clone(), vb.get_param_store().get("b").clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    vb.clear_param_store()
    with pytest.raises(NotImplementedError):
        vb.VIRT(data=y, model="irt_2pl").fit(optim=vb.Adam(per_param), max_iter=2, cuda_graph=True)


def test_noise_counters_stay_in_step_across_eager_captured_and_readout_calls():
    """The host and device call counters advance together in every mode, so no (seed, offset) Philox stream is
    used twice: eager evaluations before a captured fit, the replays, and read-outs after it."""
    torch.manual_seed(0)
    q = (torch.rand(3, 10) < 0.5).float()
    q[:, q.sum(0) == 0] = 1.0
    y = (torch.rand(300, 10) < 0.5).float()
    m = vb.VCHoDina(data=y, q=q, seed=5)
    cfg = vb.TraceEnum_ELBO(num_particles=1)
    m.loss(cfg)
    m.loss(cfg)                                             # two eager calls first: host count 2, no device counter
    assert m._calls == 2 and m._calls_dev is None
    m.fit(max_iter=7, cuda_graph=True)
    assert int(m._calls_dev) == m._calls == 2 + 7
    m.posterior()                                           # kind=None call: still advances both
    assert int(m._calls_dev) == m._calls == 10


def test_vchodina_posterior_uses_the_model_the_fit_uses():
    """VCHoDina ignores `model=` (always DINA, vi.py:905): the read-out must score with the same link."""
    torch.manual_seed(1)
    q = (torch.rand(3, 12) < 0.5).float()
    q[:, q.sum(0) == 0] = 1.0
    y = (torch.rand(200, 12) < 0.5).float()
    post = []
    for model in ("dina", "dino"):
        vb.clear_param_store()
        m = vb.VCHoDina(data=y, q=q, model=model, seed=3)
        post.append(m.posterior())
    assert torch.equal(post[0], post[1])


def test_vaeirt_mirt_fused_encoder_tail_equals_torch_autograd():
    """VaeIRT, K > 1: the default path (fc22 + sampler fused, csrc/enc_mvn.cu) against fused_encoder=False (MvnEncoder
    and its backward in torch, the factor matrix in HBM): same noise stream, same losses and parameters."""
    g = torch.Generator().manual_seed(4)
    y = (torch.rand(500, 30, generator=g) < 0.55).float()
    y[torch.rand(500, 30, generator=g) < 0.05] = float("nan")
    out = []
    for fused in (True, False):
        vb.clear_param_store()
        torch.manual_seed(7)
        m = vb.VaeIRT(data=y, model="irt_2pl", x_feature=4, hidden_dim=16, seed=9)
        m.fused_encoder = fused
        svi = infer.SVI(m.model, m.guide, optim=vb.Adam({"lr": 1e-2}), loss=vb.Trace_ELBO(num_particles=2), cuda_graph=False)
        losses = [svi.step(m.data) for _ in range(6)]
        out.append((losses, {n: vb.get_param_store().get(n).clone() for n in vb.get_param_store().names()}))
    (lf, pf), (lt, pt) = out
    assert np.allclose(lf, lt, rtol=2e-5), (lf, lt)
    for n in pf:
        assert_close(pf[n], pt[n], 2e-4, 2e-4, n)
    assert lf[-1] < lf[0]


def test_host_step_direct_two_bit_rows_edge_shapes():
    """2-bit rows expanded by the particle kernel itself (VIPSY_IRT_FLAG_Y_PACKED2, csrc/irt_uni.cu): ragged last tile,
    several chunks, the caller's count of missing cells, and the fall-back through the expansion kernel when J % 4 != 0
    or S == 1 -- all bit-identical per person to the device call on byte codes."""
    from vipsy_b200 import ops, workloads
    from vipsy_b200.host_api import IrtHostStep
    for (N, J, S, model, chunks) in ((4099, 100, 32, "irt_4pl", 3), (777, 8, 4, "irt_2pl", 1), (2050, 36, 2, "irt_3pl", 2),
                                     (1500, 50, 8, "irt_4pl", 2), (1500, 100, 1, "irt_2pl", 2)):
        w = workloads.uni_irt(N, J, model, missing=0.2, seed=7)
        names = ("a", "b", "c", "d")[:int(model[4])]                    # irt_2pl -> a, b; irt_3pl -> a, b, c; ...
        kw = {k: w[k] for k in names}
        ref = ops.irt_elbo_grad(w["y"], w["mu"], sigma=w["sigma"], n_particles=S, model=model, seed=9, offset=3, **kw)
        hy = ops.pack_responses_2bit_host(w["y"].cpu())
        host = IrtHostStep(N, J, model, n_chunks=chunks)
        hp = {k: (w[k].cpu() if k in names + ("mu", "sigma") else None) for k in ("a", "b", "c", "d", "mu", "sigma")}
        for n_masked in (None, int((w["y"] == 255).sum())):
            got = host(hy, hp["mu"], hp["sigma"], hp["a"], hp["b"], hp["c"], hp["d"], n_particles=S, seed=9, offset=3,
                       n_masked=n_masked)
            assert torch.equal(got["d_mu"], ref["d_mu"].cpu()) and torch.equal(got["d_sigma"], ref["d_sigma"].cpu()), (N, J, S)
            assert_close(got["item_grads"], ref["item_grads"], 1e-5, 1e-5, f"item grads {N}x{J} S={S}")
            assert_close(got["loss"], ref["loss"], 1e-9, 1e-9, f"loss {N}x{J} S={S}")

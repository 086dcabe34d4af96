"""CPU tests of the host-side mirror of the reference interface: param store + constraint
transforms, per-parameter Adam / MultiStepLR, SVI.step order, person sharding over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vipsy_oracle as o
from vipsy_b200 import infer
from vipsy_b200.dist import GradReducer, shard_bounds


def test_constraint_transforms_round_trip():
    torch.manual_seed(0)
    for c, v in [(infer.REAL, torch.randn(3, 4)), (infer.POSITIVE, torch.rand(5) + 0.1),
                 (infer.UNIT_INTERVAL, torch.rand(2, 6) * 0.9 + 0.05),
                 (infer.LOWER_CHOLESKY, o.lower_cholesky(torch.randn(4, 3, 3) * 0.3))]:
        u = infer.inverse_transform(c, v)
        assert torch.allclose(infer.transform(c, u), v, atol=1e-6)
    # lower_cholesky transform == oracle's LowerCholeskyTransform restatement
    raw = torch.randn(2, 5, 5)
    assert torch.equal(infer.transform(infer.LOWER_CHOLESKY, raw), o.lower_cholesky(raw))


def test_param_store_semantics():
    infer.clear_param_store()
    c = infer.param("c", torch.zeros(1, 4) + 0.1, constraint=infer.constraints.unit_interval)
    assert torch.allclose(c, torch.full((1, 4), 0.1))
    again = infer.param("c", torch.zeros(1, 4) + 0.7, constraint=infer.constraints.unit_interval)
    assert torch.allclose(again, c)                       # later calls ignore init
    st = infer.get_param_store()
    assert st.unconstrained("c").requires_grad and st.constraint("c") == infer.UNIT_INTERVAL
    assert infer.constraints.interval(0, 1) == infer.UNIT_INTERVAL
    infer.clear_param_store()
    assert "c" not in st


def test_adam_per_parameter_args_and_scheduler():
    seen = []

    def args(module_name, param_name):                    # test.py:321-327 style callable
        seen.append((module_name, param_name))
        return {"lr": 1e-2} if param_name == "a" else {"lr": 1e-3}
    a = torch.zeros(3, requires_grad=True)
    w = torch.zeros(3, requires_grad=True)
    opt = infer.Adam(args)
    a.grad, w.grad = torch.ones(3), torch.ones(3)
    opt([("a", a), ("encoder$$$fc1.weight", w)])
    assert ("a", "a") in seen and ("encoder", "fc1.weight") in seen
    assert torch.allclose(a.detach(), torch.full((3,), -1e-2), atol=1e-6)
    assert torch.allclose(w.detach(), torch.full((3,), -1e-3), atol=1e-6)
    sched = infer.MultiStepLR({"optimizer": torch.optim.Adam, "optim_args": {"lr": 0.1}, "milestones": [2], "gamma": 0.1})
    p = torch.zeros(1, requires_grad=True)
    lrs = []
    for _ in range(4):
        p.grad = torch.ones(1)
        sched([("p", p)])
        sched.step()                                      # vi.py:639-640: once per iteration
        lrs.append(sched.optim_objs[p].param_groups[0]["lr"])
    assert lrs[0] == pytest.approx(0.1) and lrs[-1] == pytest.approx(0.01)
    assert isinstance(sched, infer.PyroLRScheduler)


def test_svi_step_order_free_mask_and_zero_grads():
    class Owner:
        def __init__(self):
            self.a = torch.ones(2, 3, requires_grad=True)
            self.free = torch.tensor([[1, 1, 1], [1, 1, 0]], dtype=torch.bool)

        def model(self, data):
            pass
        guide = model

        def loss_and_grads(self, cfg, data):
            self.a.grad = torch.ones(2, 3)
            return torch.tensor(3.5), [("a", self.a, self.free)]

        def loss(self, cfg, data):
            return torch.tensor(1.25)
    ow = Owner()
    svi = infer.SVI(ow.model, ow.guide, infer.Adam({"lr": 0.5}), infer.Trace_ELBO(num_particles=2))
    assert svi.step(None) == 3.5
    assert ow.a.grad is None                              # zero_grads after the optimiser step
    assert float(ow.a[1, 2]) == 1.0                       # frozen entry did not move (vi.py:511-512)
    assert float(ow.a[0, 0]) == pytest.approx(0.5)
    assert svi.evaluate_loss(None) == 1.25


def test_supports_fused_checks_the_resolved_optimiser_arguments():
    """The device-resident Adam covers torch.optim.Adam's default options only: anything else must be reported by
    supports_fused (so fit() can fall back to the eager torch.optim path) instead of failing on the first step."""
    assert infer.Adam({"lr": 1e-2}).supports_fused()
    assert infer.Adam({"lr": 1e-2, "betas": (0.8, 0.99)}).supports_fused()
    assert not infer.Adam({"lr": 1e-2, "weight_decay": 1e-4}).supports_fused()
    assert not infer.Adam({"lr": 1e-2, "amsgrad": True}).supports_fused()
    assert not infer.Adam({"lr": 1e-2, "maximize": True}).supports_fused()
    assert not infer.Adam({"lr": 1e-2}, clip_args={"clip_norm": 1.0}).supports_fused()
    assert not infer.PyroOptim(torch.optim.SGD, {"lr": 1e-2}).supports_fused()

    def per_param(module_name, param_name):
        return {"lr": 1e-2, "weight_decay": 1e-3} if param_name == "b" else {"lr": 1e-2}
    opt = infer.Adam(per_param)
    assert opt.supports_fused(["a", "encoder$$$fc1.weight"]) and not opt.supports_fused(["a", "b"])
    sched = infer.MultiStepLR({"optimizer": torch.optim.Adam, "optim_args": {"lr": 0.1, "amsgrad": True},
                               "milestones": [2], "gamma": 0.1})
    assert not sched.supports_fused()


def test_svi_auto_mode_falls_back_to_eager_and_strict_mode_raises():
    class Owner:
        _calls = 0

        def __init__(self):
            self.b = torch.zeros(3, requires_grad=True)

        def model(self, data):
            pass
        guide = model

        def loss_and_grads(self, cfg, data):
            self.b.grad = torch.ones(3)
            return torch.tensor(2.0), [("b", self.b, None)]

    def per_param(module_name, param_name):
        return {"lr": 0.5, "weight_decay": 0.1}
    ow = Owner()
    svi = infer.SVI(ow.model, ow.guide, infer.Adam(per_param), infer.Trace_ELBO(), cuda_graph=None)
    assert svi.cuda_graph                                   # callable arguments: undecided until the names are known
    assert svi.step(None) == 2.0 and not svi.cuda_graph     # decided on the first step: eager torch.optim.Adam
    assert float(ow.b[0]) == pytest.approx(-0.5) and ow.b.grad is None
    assert svi.step(None) == 2.0
    with pytest.raises(NotImplementedError):
        infer.SVI(ow.model, ow.guide, infer.Adam({"lr": 0.5, "amsgrad": True}), infer.Trace_ELBO(), cuda_graph=True)
    strict = infer.SVI(ow.model, ow.guide, infer.Adam(per_param), infer.Trace_ELBO(), cuda_graph=True)
    with pytest.raises(NotImplementedError):
        strict.step(None)
    auto = infer.SVI(ow.model, ow.guide, infer.Adam({"lr": 0.5}, clip_args={"clip_value": 0.1}), infer.Trace_ELBO(),
                     cuda_graph=None)
    assert not auto.cuda_graph


def test_clip_args_are_applied_per_parameter_before_the_step():
    """pyro.optim clip_args: {'clip_norm': x} / {'clip_value': y}, dict or callable, per parameter tensor."""
    p = torch.zeros(4, requires_grad=True)
    q = torch.zeros(4, requires_grad=True)
    opt = infer.PyroOptim(torch.optim.SGD, {"lr": 1.0}, clip_args={"clip_norm": 1.0})
    p.grad, q.grad = torch.full((4,), 3.0), torch.full((4,), 0.1)
    opt([("p", p), ("q", q)])
    assert torch.allclose(p.detach(), torch.full((4,), -0.5), atol=1e-5)      # norm 6 -> 1
    assert torch.allclose(q.detach(), torch.full((4,), -0.1), atol=1e-6)      # norm 0.2: untouched
    r = torch.zeros(2, requires_grad=True)
    optv = infer.PyroOptim(torch.optim.SGD, {"lr": 1.0},
                           clip_args=lambda m, n: {"clip_value": 0.25} if n == "r" else {})
    r.grad = torch.tensor([2.0, -0.1])
    optv([("r", r)])
    assert torch.allclose(r.detach(), torch.tensor([-0.25, 0.1]))
    with pytest.raises(ValueError):
        infer.PyroOptim(torch.optim.SGD, {"lr": 1.0}, clip_args={"clip": 1.0})([("p", p)])


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        N, J, S = 101, 9, 3
        y = (torch.rand(N, J, generator=g) < 0.5).float()
        y[torch.rand(N, J, generator=g) < 0.2] = float("nan")
        mu, sg = torch.randn(N, 1, generator=g) * 0.5, torch.exp(torch.randn(N, 1, generator=g) * 0.2)
        a, b = torch.rand(1, J, generator=g) + 0.5, torch.randn(1, J, generator=g)
        eps = torch.randn(S, N, 1, generator=g)
        lo, hi = shard_bounds(N, world, rank)
        r = o.irt_elbo_grad(y[lo:hi], mu[lo:hi], sg[lo:hi], eps[:, lo:hi], a, b, model="irt_2pl", scale=1.0)
        la, lb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        la.grad, lb.grad = r["d_a"].float(), r["d_b"].float()
        red = GradReducer()
        loss = red(torch.tensor(r["loss"], dtype=torch.float64), [la, lb])
        full = o.irt_elbo_grad(y, mu, sg, eps, a, b, model="irt_2pl")
        ok = (abs(float(loss) - full["loss"]) < 1e-4 * abs(full["loss"])
              and torch.allclose(la.grad.double(), full["d_a"], rtol=1e-5, atol=1e-4)
              and torch.allclose(lb.grad.double(), full["d_b"], rtol=1e-5, atol=1e-4)
              and torch.allclose(r["d_mu"], full["d_mu"][lo:hi])          # per-person grads stay local
              and red.bytes_last == 4 * (2 + 2 * J))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_person_sharding_world2_gloo():
    """Sum of per-shard item gradients == full-batch gradients; one packed all-reduce per step."""
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_two_bit_transport_packer_host():
    """vipsy_pack_responses_2bit_host: 4 cells per byte, codes 0/1/255 -> 0/1/3, ragged last byte padded with 3."""
    from vipsy_b200 import build, ops
    build.build()
    g = torch.Generator().manual_seed(3)
    for J in (1, 4, 5, 100, 203):
        codes = (torch.rand(37, J, generator=g) < 0.5).to(torch.uint8)
        codes[torch.rand(37, J, generator=g) < 0.3] = 255
        packed = ops.pack_responses_2bit_host(codes, pin=False)
        assert packed.shape == (37, (J + 3) // 4)
        pad = torch.full((37, 4 * packed.shape[1]), 255, dtype=torch.uint8)
        pad[:, :J] = codes
        two = (pad & 3).view(37, -1, 4).to(torch.int32)
        want = two[..., 0] | (two[..., 1] << 2) | (two[..., 2] << 4) | (two[..., 3] << 6)
        assert torch.equal(packed.to(torch.int32), want)
    bad = torch.full((2, 8), 7, dtype=torch.uint8)
    with pytest.raises(Exception):
        ops.pack_responses_2bit_host(bad, pin=False)


def test_oracle_norm_encoder_is_the_reference_module():
    """oracle.norm_encoder restates NormEncoder.forward + the guide's NaN -> -1 rule (vi.py:417-435, 679-682):
    same numbers as the nn.Module the estimator classes hold, and differentiable in the weights."""
    import torch
    from oracle import vipsy_oracle as o
    from vipsy_b200.models import NormEncoder
    torch.manual_seed(0)
    enc = NormEncoder(9, 1, 6).double()
    y = (torch.rand(20, 9) < 0.5).double()
    y[torch.rand(20, 9) < 0.2] = float("nan")
    x = torch.where(torch.isnan(y), torch.full_like(y, -1.0), y)
    loc, sc = enc(x)
    p = dict(enc.named_parameters())
    loc2, sc2 = o.norm_encoder(y, p["fc1.weight"], p["fc1.bias"], p["fc21.weight"], p["fc21.bias"], p["fc22.weight"],
                               p["fc22.bias"])
    assert torch.allclose(loc, loc2, atol=1e-12) and torch.allclose(sc, sc2, atol=1e-12)
    (loc2.sum() + sc2.sum()).backward()
    assert p["fc1.weight"].grad is not None and float(p["fc1.weight"].grad.abs().sum()) > 0

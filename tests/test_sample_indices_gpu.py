"""GPU: the single-launch minibatch draw of the `data` plate (vipsy_sample_indices_dev) -- what
pyro.plate(subsample_size=B) draws as randperm(N)[:B]: distinct, in range, uniform, deterministic in
(seed, counter), counter advanced on the device (CUDA-graph safe)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from vipsy_b200 import build
    build.build()
    from vipsy_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("N,B", [(1000, 100), (100_000, 100), (5, 5), (2048, 2048), (3_000_000, 2000), (7, 1)])
def test_distinct_in_range_and_counter(ops, N, B):
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    a = ops.sample_indices(N, B, 123, c)
    b = ops.sample_indices(N, B, 123, c)
    torch.cuda.synchronize()
    assert int(c) == 2
    for x in (a, b):
        assert x.shape == (B,) and int(x.min()) >= 0 and int(x.max()) < N
        assert x.unique().numel() == B
    if B > 1 and N > 4 * B:
        assert not torch.equal(a, b)                         # a fresh draw per call
    c.zero_()
    a2 = ops.sample_indices(N, B, 123, c)
    assert torch.equal(a, a2)                                # deterministic in (seed, counter)
    c.zero_()
    a3 = ops.sample_indices(N, B, 124, c)
    if N > 4 * B:
        assert not torch.equal(a, a3)


def test_uniform_over_values_and_positions(ops):
    N, B, R = 200, 50, 4000
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    draws = torch.stack([ops.sample_indices(N, B, 7, c) for _ in range(R)])          # R x B
    counts = torch.bincount(draws.reshape(-1), minlength=N).double()
    exp = R * B / N                                                                  # 1000, sd ~ 27
    assert float((counts - exp).abs().max()) < 6 * (exp * (1 - B / N)) ** 0.5
    first = torch.bincount(draws[:, 0], minlength=N).double()                        # position 0 alone: R / N = 20
    assert float(first.max()) < 20 + 6 * 20 ** 0.5 and float(first.min()) > 20 - 5 * 20 ** 0.5 - 1
    # a full draw (B == N) is a permutation, uniform over positions
    perms = torch.stack([ops.sample_indices(6, 6, 9, c) for _ in range(3000)])
    assert bool((perms.sort(dim=1).values == torch.arange(6, device="cuda")).all())
    pos = torch.stack([torch.bincount(perms[:, k], minlength=6) for k in range(6)]).double()
    assert float((pos - 500).abs().max()) < 6 * (500 * 5 / 6) ** 0.5


def test_graph_replay_draws_fresh_minibatches(ops):
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.sample_indices(10_000, 64, 1, c)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = ops.sample_indices(10_000, 64, 1, c)
    seen = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        seen.append(out.clone())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
    assert int(c) == 4


def test_rejects_large_batch(ops):
    from vipsy_b200 import _native
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    with pytest.raises(_native.VipsyError):
        ops.sample_indices(10_000, 2049, 1, c)

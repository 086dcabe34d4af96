"""tcgen05 building blocks of the tensor-core MIRT kernel: each operand orientation
(K-major / MN-major shared-memory descriptors, fp16 hi/lo split, TMEM read-out) against
a float64 matmul."""
import pytest
import torch

from vipsy_b200 import _native as N

pytestmark = pytest.mark.gpu


def run(variant, t1, t2, swap=0, n_split=3):
    L = N.lib()
    n = t2.shape[0] if variant in (1, 4) else t2.shape[1]
    out = torch.full((128, n), float("nan"), device="cuda")
    N.check(L.vipsy_selftest_umma(t1.data_ptr(), t1.shape[0], t1.shape[1], t2.data_ptr(), t2.shape[0], t2.shape[1],
                                  out.data_ptr(), variant, swap, n_split, None))
    torch.cuda.synchronize()
    return out


CASES = {0: ((128, 112), (112, 128)), 1: ((128, 128), (112, 128)), 2: ((128, 112), (128, 128)),
         3: ((128, 112), (112, 64)), 4: ((128, 80), (128, 80))}       # 3, 4: A in tensor memory


def expect(variant, t1, t2):
    a, b = t1.double(), t2.double()
    return a @ b if variant in (0, 3) else (a @ b.T if variant in (1, 4) else a.T @ b)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_tile_product_matches_float64(variant):
    g = torch.Generator(device="cuda").manual_seed(variant)
    s1, s2 = CASES[variant]
    t1 = torch.randn(s1, device="cuda", generator=g)
    t2 = torch.randn(s2, device="cuda", generator=g) * 1.5
    ref = expect(variant, t1, t2)
    got = run(variant, t1, t2).double()
    rows = ref.shape[0]
    err = float((got[:rows] - ref).abs().max() / ref.abs().max())
    assert err < 2e-6, err
    # one fp16 term alone is ~1e-3 accurate: the split is what buys fp32-class accuracy
    got1 = run(variant, t1, t2, n_split=1).double()
    err1 = float((got1[:rows] - ref).abs().max() / ref.abs().max())
    assert 1e-5 < err1 < 1e-2, err1

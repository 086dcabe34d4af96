import torch


def assert_close(got, ref, rtol=1e-5, atol_frac=1e-5, what=""):
    """|got - ref| <= rtol*|ref| + atol_frac*max|ref|  (ref is the fp64 oracle)."""
    g = got.detach().to("cpu", torch.float64).reshape(-1) if isinstance(got, torch.Tensor) else torch.tensor([float(got)], dtype=torch.float64)
    r = ref.detach().to("cpu", torch.float64).reshape(-1) if isinstance(ref, torch.Tensor) else torch.tensor([float(ref)], dtype=torch.float64)
    assert g.shape == r.shape, (what, g.shape, r.shape)
    assert torch.isfinite(g).all(), f"{what}: non-finite values"
    tol = rtol * r.abs() + atol_frac * max(float(r.abs().max()), 1e-30)
    err = (g - r).abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{g.numel()} off; max err {float(err.max()):.3e} "
                           f"(ref scale {float(r.abs().max()):.3e})")


def codes_from_float(y):
    """NaN-coded float matrix -> uint8 codes on CPU (independent of the CUDA packer)."""
    c = (y != 0).to(torch.uint8)
    c[torch.isnan(y)] = 255
    return c

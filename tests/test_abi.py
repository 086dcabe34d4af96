"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and
exports every symbol include/vipsy_b200.h declares; ctypes mirrors the structs."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from vipsy_b200 import _native as N
from vipsy_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vipsy_b200.h")


@pytest.fixture(scope="module")
def lib():
    B.build()
    return N.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vipsy_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert n in N.SYMBOLS, f"{n} declared in the header but not bound"
        assert getattr(lib, n) is not None
    assert lib.vipsy_abi_version() == 3


def test_struct_layout_matches_header():
    structs = {"VipsyIrtArgs": N.IrtArgs, "VipsyCdmEnumArgs": N.CdmEnumArgs, "VipsyCdmScoreArgs": N.CdmScoreArgs,
               "VipsyIrtHostArgs": N.IrtHostArgs, "VipsyAdamTensor": N.AdamTensor,
               "VipsyNormEncoderArgs": N.NormEncoderArgs, "VipsyMvnEncoderIrtArgs": N.MvnEncoderIrtArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, st in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in st._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-std=c99", "-o", exe, src], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for cname, st in structs.items():
        assert int(got[cname]) == C.sizeof(st), cname
        for f, _ in st._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(st, f).offset, f"{cname}.{f}"


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    a = N.IrtArgs()
    assert lib.vipsy_irt_elbo_grad(C.byref(a), None) == N.ERR_ARG
    assert b"model" in lib.vipsy_last_error()
    assert lib.vipsy_irt_elbo_grad(None, None) == N.ERR_ARG
    assert lib.vipsy_irt_workspace_bytes(C.byref(a)) == 0
    with pytest.raises(N.VipsyError):
        N.check(lib.vipsy_pack_responses(None, 4, None, None))


def test_peer_allreduce_rejects_bad_layouts(lib):
    """vipsy_peer_allreduce argument validation (the kernel itself needs >= 2 GPUs: scripts/test_peer_allreduce.py)."""
    f = lib.vipsy_peer_allreduce
    assert f(None, 0, 2, 1024, 512, 4, 1, None, None, 1, None) == N.ERR_ARG            # no pointer table
    dummy = C.c_uint64(0)
    assert f(C.byref(dummy), 2, 2, 1024, 512, 4, 1, None, None, 1, None) == N.ERR_ARG  # rank >= world
    assert f(C.byref(dummy), 0, 2, 4, 512, 4, 1, None, None, 1, None) == N.ERR_ARG     # flag area too small
    assert f(C.byref(dummy), 0, 2, 1024, 16, 400, 1, None, None, 1, None) == N.ERR_ARG  # slot too small
    assert b"vipsy_peer_allreduce" in lib.vipsy_last_error()


def test_host_step_rejects_half_a_gradient_pair(lib):
    """d_mu_host without d_sigma_host (or the reverse) is an argument error, not a write through NULL."""
    import torch
    a = N.IrtHostArgs()
    a.dev.n_rows = a.dev.batch = 64
    a.dev.n_items, a.dev.n_dims, a.dev.n_particles, a.dev.model = 4, 1, 1, 2
    keep = [torch.zeros(64 * 4, dtype=torch.uint8), torch.zeros(64), torch.ones(64), torch.ones(4), torch.zeros(4),
            torch.zeros(1, dtype=torch.float64), torch.zeros(64)]
    a.y_host, a.mu_host, a.sigma_host, a.a_host, a.b_host, a.loss_host, a.d_mu_host = (t.data_ptr() for t in keep)
    a.n_chunks = 1
    assert N.lib().vipsy_irt_host_step(C.byref(a)) == N.ERR_ARG
    assert b"go together" in N.lib().vipsy_last_error()


def test_two_bit_rows_flag_is_validated_without_a_gpu(lib):
    """VIPSY_IRT_FLAG_Y_PACKED2 (the particle kernel expands 2-bit transport rows itself) is only defined for
    K == 1, S > 1, the pathwise estimator, J % 4 == 0, y_row_stride == J / 4: anything else is refused before a launch."""
    import torch
    keep = [torch.zeros(64 * 2, dtype=torch.uint8), torch.zeros(64), torch.ones(64), torch.ones(6), torch.zeros(6),
            torch.zeros(1, dtype=torch.float64)]

    def args(J, S, stride):
        a = N.IrtArgs()
        a.n_rows = a.batch = 64
        a.n_items, a.n_dims, a.n_particles, a.model, a.estimator = J, 1, S, 2, N.EST["pathwise"]
        a.y, a.mu, a.sigma, a.a, a.b, a.loss = (t.data_ptr() for t in keep)
        a.y_row_stride, a.flags = stride, N.IRT_FLAG_Y_PACKED2
        return a
    for J, S, stride in ((6, 8, 2), (8, 1, 2), (8, 8, 8)):          # J % 4 != 0; S == 1; stride of a byte matrix
        assert lib.vipsy_irt_elbo_grad(C.byref(args(J, S, stride)), None) == N.ERR_UNSUPPORTED
        assert b"Y_PACKED2" in lib.vipsy_last_error()

"""ctypes binding of include/vipsy_b200.h (the C-ABI shared library).

There is no CPU fallback: if the library is missing or an entry point fails the
caller gets an exception, never a silently different code path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libvipsy_b200.so")

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3, -4
Y_MISSING = 255
IRT_MODEL = {"irt_1pl": 1, "irt_2pl": 2, "irt_3pl": 3, "irt_4pl": 4}
CDM_MODEL = {"dina": 0, "dino": 1}
EST = {"pathwise": 0, "score": 1}
COV_DIAG1, COV_SHARED, COV_PERSON = 0, 1, 2
CHOL_DENSE, CHOL_PACKED = 0, 1
IRT_FLAG_STRICT = 1
IRT_FLAG_NO_TENSOR = 2
IRT_FLAG_SPARSE = 4
IRT_FLAG_NO_ROW = 8
IRT_FLAG_MASKED_KNOWN = 16
IRT_FLAG_NO_FUSED = 32
IRT_FLAG_Y_PACKED2 = 64

_p = C.c_void_p


class VipsyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vipsy_b200 error {code}: {msg}")
        self.code = code


class IrtArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("batch", C.c_int64),
        ("n_items", C.c_int32), ("n_dims", C.c_int32), ("n_particles", C.c_int32), ("model", C.c_int32),
        ("estimator", C.c_int32), ("cov_mode", C.c_int32), ("chol_layout", C.c_int32), ("chol_raw", C.c_int32),
        ("D", C.c_float), ("scale", C.c_float), ("baseline", C.c_float), ("flags", C.c_int32),
        ("y", _p), ("y_row_stride", C.c_int64), ("idx", _p), ("mu", _p), ("sigma", _p), ("chol", _p),
        ("eps", _p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("person_offset", C.c_int64),
        ("a", _p), ("b", _p), ("c", _p), ("d", _p), ("a_free", _p),
        ("loss", _p), ("d_mu", _p), ("d_sigma", _p), ("d_chol", _p),
        ("d_a", _p), ("d_b", _p), ("d_c", _p), ("d_d", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t),
        ("obs_row_ptr", _p), ("obs_entries", _p), ("obs_nnz", C.c_int64), ("n_masked_cells", C.c_int64),
    ]


class IrtHostArgs(C.Structure):
    _fields_ = [
        ("dev", IrtArgs), ("y_host", _p), ("y_packed2", C.c_int32), ("n_chunks", C.c_int32),
        ("mu_host", _p), ("sigma_host", _p), ("a_host", _p), ("b_host", _p), ("c_host", _p), ("d_host", _p),
        ("loss_host", _p), ("d_mu_host", _p), ("d_sigma_host", _p), ("item_grads_host", _p),
        ("dev_scratch", _p), ("dev_scratch_bytes", C.c_size_t), ("s_in", _p), ("s_run", _p), ("s_out", _p),
        ("a_free_host", _p),
    ]


class MvnEncoderIrtArgs(C.Structure):
    _fields_ = [
        ("irt", IrtArgs), ("n_hidden", C.c_int32), ("reserved", C.c_int32),
        ("hidden", _p), ("w22", _p), ("b22", _p), ("d_hidden", _p), ("d_w22", _p), ("d_b22", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t),
        ("w21", _p), ("b21", _p), ("d_w21", _p), ("d_b21", _p),
    ]


class CdmEnumArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("batch", C.c_int64),
        ("n_items", C.c_int32), ("n_attrs", C.c_int32), ("n_particles", C.c_int32), ("model", C.c_int32),
        ("hodina", C.c_int32), ("reserved0", C.c_int32), ("scale", C.c_float), ("reserved1", C.c_float),
        ("y", _p), ("y_row_stride", C.c_int64), ("idx", _p), ("qmask", _p), ("g", _p), ("s", _p),
        ("lam0", _p), ("lam1", _p), ("mu", _p), ("sigma", _p), ("eps", _p),
        ("seed", C.c_uint64), ("offset", C.c_uint64), ("person_offset", C.c_int64),
        ("loss", _p), ("d_g", _p), ("d_s", _p), ("d_lam0", _p), ("d_lam1", _p), ("d_mu", _p), ("d_sigma", _p),
        ("posterior", _p), ("workspace", _p), ("workspace_bytes", C.c_size_t),
        ("log_prior", _p), ("score_missing", C.c_int32), ("reserved2", C.c_int32),
    ]


class CdmScoreArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("batch", C.c_int64),
        ("n_items", C.c_int32), ("n_attrs", C.c_int32), ("n_particles", C.c_int32), ("model", C.c_int32),
        ("scale", C.c_float), ("baseline", C.c_float),
        ("y", _p), ("y_row_stride", C.c_int64), ("idx", _p), ("qmask", _p), ("g", _p), ("s", _p),
        ("attr_p", _p), ("u", _p), ("seed", C.c_uint64), ("offset", C.c_uint64), ("person_offset", C.c_int64),
        ("loss", _p), ("d_g", _p), ("d_s", _p), ("d_attr_p", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t),
    ]


class AdamTensor(C.Structure):
    _fields_ = [
        ("param", _p), ("grad", _p), ("exp_avg", _p), ("exp_avg_sq", _p), ("free_mask", _p), ("lr_dev", _p),
        ("n", C.c_int64), ("step_offset", C.c_int64),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
    ]


class NormEncoderArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int64), ("batch", C.c_int64), ("n_items", C.c_int32), ("n_hidden", C.c_int32),
        ("y", _p), ("y_row_stride", C.c_int64), ("idx", _p),
        ("w1", _p), ("b1", _p), ("w21", _p), ("b21", _p), ("w22", _p), ("b22", _p),
        ("mu", _p), ("sigma", _p), ("pre", _p), ("d_mu", _p), ("d_sigma", _p),
        ("grad_scale", C.c_float), ("flags", C.c_int32),
        ("d_w1", _p), ("d_b1", _p), ("d_w21", _p), ("d_b21", _p), ("d_w22", _p), ("d_b22", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t), ("x_tiles", _p),
    ]


# every symbol include/vipsy_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vipsy_abi_version": (C.c_int, []),
    "vipsy_last_error": (C.c_char_p, []),
    "vipsy_irt_workspace_bytes": (C.c_size_t, [C.POINTER(IrtArgs)]),
    "vipsy_irt_elbo_grad": (C.c_int, [C.POINTER(IrtArgs), _p]),
    "vipsy_irt_host_scratch_bytes": (C.c_size_t, [C.POINTER(IrtHostArgs)]),
    "vipsy_irt_host_step": (C.c_int, [C.POINTER(IrtHostArgs)]),
    "vipsy_irt_mvn_encoder_workspace_bytes": (C.c_size_t, [C.POINTER(MvnEncoderIrtArgs)]),
    "vipsy_irt_mvn_encoder_elbo_grad": (C.c_int, [C.POINTER(MvnEncoderIrtArgs), _p]),
    "vipsy_mvn_encoder_sample": (C.c_int, [C.POINTER(MvnEncoderIrtArgs), _p, _p, _p, _p, _p]),
    "vipsy_cdm_enum_workspace_bytes": (C.c_size_t, [C.POINTER(CdmEnumArgs)]),
    "vipsy_cdm_enum_elbo_grad": (C.c_int, [C.POINTER(CdmEnumArgs), _p]),
    "vipsy_cdm_score_workspace_bytes": (C.c_size_t, [C.POINTER(CdmScoreArgs)]),
    "vipsy_cdm_score_elbo_grad": (C.c_int, [C.POINTER(CdmScoreArgs), _p]),
    "vipsy_pack_responses": (C.c_int, [_p, C.c_int64, _p, _p]),
    "vipsy_pack_responses_2bit_host": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, _p]),
    "vipsy_unpack_responses_2bit": (C.c_int, [_p, C.c_int64, C.c_int32, _p, C.c_int64, _p]),
    "vipsy_observed_row_ptr": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, _p, _p, _p]),
    "vipsy_observed_fill": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, _p, _p, _p]),
    "vipsy_cdm_eta_table": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32, _p, _p]),
    "vipsy_fill_normal": (C.c_int, [_p, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, C.c_int64, _p]),
    "vipsy_fill_normal_dev": (C.c_int, [_p, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, _p, C.c_int64, _p]),
    "vipsy_fill_uniform_dev": (C.c_int, [_p, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, _p, C.c_int64, _p]),
    "vipsy_peer_allreduce": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _p, _p,
                                       C.c_uint32, _p]),
    "vipsy_sample_indices_dev": (C.c_int, [_p, C.c_int32, C.c_int64, C.c_uint64, _p, _p]),
    "vipsy_norm_encoder_workspace_bytes": (C.c_size_t, [C.POINTER(NormEncoderArgs)]),
    "vipsy_norm_encoder_tiles_bytes": (C.c_size_t, [C.c_int64, C.c_int32]),
    "vipsy_norm_encoder_pack_tiles": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, _p, _p]),
    "vipsy_norm_encoder_forward": (C.c_int, [C.POINTER(NormEncoderArgs), _p]),
    "vipsy_norm_encoder_backward": (C.c_int, [C.POINTER(NormEncoderArgs), _p]),
    "vipsy_adam_step": (C.c_int, [C.POINTER(AdamTensor), C.c_int32, _p, C.c_int32, _p]),
    "vipsy_selftest_umma": (C.c_int, [_p, C.c_int32, C.c_int32, _p, C.c_int32, C.c_int32, _p, C.c_int32, C.c_int32,
                                      C.c_int32, _p]),
    "vipsy_fill_uniform": (C.c_int, [_p, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64, C.c_int64, _p]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VipsyError(ERR_CUDA, f"{LIB_PATH} not built; run `python -m vipsy_b200.build` "
                                       "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.vipsy_abi_version() != 3:
            raise VipsyError(ERR_ARG, "ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise VipsyError(rc, lib().vipsy_last_error().decode())

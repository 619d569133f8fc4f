"""ctypes binding of libaasist_b200.so (the C ABI declared in include/aasist_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call fails, an
exception is raised -- the product path never computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AASIST_B200_LIB: developer override (kernel experiments build variant libraries next to the tree)
LIB_PATH = os.environ.get("AASIST_B200_LIB") or os.path.join(_HERE, "csrc", "libaasist_b200.so")

KIND_AASIST, KIND_RAWGAT_ST, KIND_ROBUST = 0, 1, 2
ENC_RESIDUAL23, ENC_RES2NET, ENC_RESIDUAL33 = 0, 1, 2
ABI_VERSION = 2
PREC_FP32, PREC_F16X3, PREC_F16X2 = 0, 1, 2
PRECISIONS = {"fp32": PREC_FP32, "f16x3": PREC_F16X3, "f16x2": PREC_F16X2}


class AasistConfig(C.Structure):
    """Mirror of ``struct aasist_config`` (include/aasist_b200.h)."""
    _fields_ = [
        ("kind", C.c_int32), ("precision", C.c_int32), ("first_conv", C.c_int32),
        ("n_filters", C.c_int32), ("enc_channels", (C.c_int32 * 2) * 6),
        ("gat_dims", C.c_int32 * 2), ("pool_ratios", C.c_double * 4),
        ("temperatures", C.c_double * 4), ("sample_rate", C.c_int32), ("encoder", C.c_int32),
        ("res2net_width", C.c_int32), ("res2net_scale", C.c_int32), ("spk_emb_dim", C.c_int32),
        ("spk_level", C.c_int32), ("spk_use_attention", C.c_int32), ("reserved", C.c_int32 * 1),
    ]


class ForwardOpts(C.Structure):
    """Mirror of ``struct aasist_forward_opts``."""
    _fields_ = [("speaker_embedding", C.c_void_p), ("freq_mask_start", C.c_int32), ("freq_mask_count", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class AasistError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libaasist_b200 error {code}: {message}")
        self.code = code


_lib = None

# name -> (restype, argtypes); every symbol include/aasist_b200.h declares
SIGNATURES = {
    "aasist_abi_version": (C.c_int, []),
    "aasist_last_error": (C.c_char_p, []),
    "aasist_create": (C.c_int, [C.POINTER(AasistConfig), C.POINTER(C.c_void_p)]),
    "aasist_destroy": (C.c_int, [C.c_void_p]),
    "aasist_set_param": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "aasist_num_params": (C.c_int, [C.c_void_p]),
    "aasist_param_name": (C.c_char_p, [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]),
    "aasist_finalize": (C.c_int, [C.c_void_p]),
    "aasist_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32]),
    "aasist_hidden_dim": (C.c_int, [C.c_void_p]),
    "aasist_topk_layout": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "aasist_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "aasist_forward_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ForwardOpts), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "aasist_stage_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_void_p]),
    "aasist_pad_sequence_length": (C.c_int32, [C.POINTER(C.c_int32), C.c_int32]),
    "aasist_score_begin": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "aasist_score_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "aasist_score_finish": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "aasist_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "aasist_pad_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p]),
    "aasist_det_workspace_bytes": (C.c_int64, [C.c_int64]),
    "aasist_det_metrics": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                     C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "aasist_get_filterbank": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "aasist_frontend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_void_p]),
    "aasist_encoder_block": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "aasist_graph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "aasist_launch_count": (C.c_int64, [C.c_void_p]),
    "aasist_input_range_exceeded": (C.c_int, [C.c_void_p, C.c_int32]),
    "aasist_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "aasist_profile_report": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.c_int32]),
}


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m aasist_b200.build` (needs nvcc); "
            "aasist_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.aasist_abi_version() != ABI_VERSION:
        raise ImportError("libaasist_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(code: int) -> int:
    if code < 0:
        msg = load().aasist_last_error()
        raise AasistError(int(code), msg.decode() if msg else "")
    return code

"""ctypes binding of libopnet_b200.so (the C ABI declared in include/opnet_b200.h).

The library is the only compute path of this package.  If it is missing the import of any
op fails loudly -- there is deliberately no PyTorch or CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_uint64, c_ulonglong, c_void_p, POINTER

_PKG = os.path.dirname(os.path.abspath(__file__))
# OPN_B200_LIB selects a development variant of the library (e.g. the phase-counter build); never a fallback
LIB_PATH = os.environ.get("OPN_B200_LIB") or os.path.join(_PKG, "lib", "libopnet_b200.so")

OPN_OK = 0
OPN_ERR_TIMEOUT = -4

# name -> (restype, argtypes); mirrors include/opnet_b200.h one to one
_P = c_void_p


class WgradJob(ctypes.Structure):
    """opn_wgrad_job of include/opnet_b200.h."""
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("out", c_void_p), ("lda", c_int64), ("ldb", c_int64), ("ldc", c_int64),
                ("rows", c_int64), ("T", c_int64), ("M", c_int64), ("N", c_int64), ("shift", c_int32), ("trans_out", c_int32)]


SIGNATURES = {
    "opn_version": (c_int, []),
    "opn_last_error": (c_char_p, []),
    "opn_launch_count": (c_ulonglong, []),
    "opn_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "opn_sgemm_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "opn_sgemm": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_float, _P, c_int64, _P, c_int64, c_float, _P,
                          c_int64, _P, c_int, _P, c_int64, _P]),
    "opn_lstm_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "opn_lstm_fwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "opn_lstm_bwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "opn_lstm_batchwide": (c_int, [c_int64, c_int64]),
    "opn_lstm_status": (c_int, [_P, POINTER(c_uint32)]),
    "opn_set_status_page": (c_int, [_P]),
    "opn_set_precision": (c_int, [c_int]),
    "opn_get_precision": (c_int, []),
    "opn_opnet_fwd_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "opn_opnet_fwd": (c_int, [c_int64, c_int64, c_int64, c_int64] + [_P] * 16 + [c_int64, _P]),
    "opn_opnet_bwd_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "opn_opnet_bwd": (c_int, [c_int64, c_int64, c_int64, c_int64] + [_P] * 15 + [c_int64, _P]),
    "opn_opnet_bwd_begin": (c_int, [c_int64, c_int64, c_int64, c_int64] + [_P] * 15 + [c_int64, _P]),
    "opn_opnet_bwd_join": (c_int, [_P]),
    "opn_wgrad_workspace_bytes": (c_int64, [c_int32, POINTER(WgradJob)]),
    "opn_wgrad": (c_int, [c_int32, POINTER(WgradJob), _P, c_int64, _P]),
    "opn_attention_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "opn_attention_fwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, _P]),
    "opn_attention_bwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int64, c_float, c_uint64, c_uint64, _P]),
    "opn_wtt_fwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "opn_wtt_bwd": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "opn_relu_bwd": (c_int, [c_int64, _P, _P, _P]),
    "opn_colsum": (c_int, [c_int64, c_int64, _P, c_int64, _P, c_int, _P]),
    "opn_softmax_rows": (c_int, [c_int64, c_int64, _P, c_int64, c_float, _P]),
    "opn_softmax_rows_bwd": (c_int, [c_int64, c_int64, _P, _P, c_int64, c_float, _P]),
    "opn_layernorm_fwd": (c_int, [c_int64, c_int64, _P, _P, _P, _P, c_float, _P, _P, _P, _P]),
    "opn_layernorm_bwd": (c_int, [c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "opn_add": (c_int, [c_int64, _P, _P, _P, _P]),
    "opn_dropout": (c_int, [c_int64, _P, _P, c_float, c_uint64, c_uint64, _P]),
    "opn_loss_fwd_bwd": (c_int, [c_int64, c_int64, _P, _P, _P, c_int, _P, _P, _P]),
    "opn_head_loss_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "opn_head_loss": (c_int, [c_int64, c_int64, c_int64, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, c_int64, _P]),
    "opn_adam_step": (c_int, [c_int64, _P, _P, _P, _P, c_float, c_float, c_float, c_float, c_float, c_int64, _P]),
    "opn_iou_eval": (c_int, [c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "opn_to_pixels": (c_int, [c_int64, _P, _P, _P]),
}

_lib = None


class OpnError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OpnError(
            f"{LIB_PATH} is missing.  objectpermanence_b200 has no fallback path: build the CUDA library with "
            "`python -m objectpermanence_b200.build` (needs nvcc) or call __graft_entry__.build().")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OPN_OK:
        msg = load().opn_last_error().decode("utf-8", "replace")
        raise OpnError(f"libopnet_b200 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().opn_launch_count())

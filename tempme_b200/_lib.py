"""ctypes binding of libtempme_b200.so (the C ABI declared in include/tempme_b200.h).

The product path has no fallback: if the CUDA library is missing and cannot be built, importing
this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

TM_EIDX_NONE = -(2 ** 31)
TM_ERR_UNSUPPORTED = -6

_p = C.c_void_p
_i64 = C.c_int64
_u64 = C.c_uint64


class EncoderDesc(C.Structure):
    _fields_ = [("node_dim", C.c_int32), ("edge_dim", C.c_int32), ("hid_dim", C.c_int32),
                ("use_temporal", C.c_int32), ("if_cat", C.c_int32), ("edge_projected", C.c_int32), ("walk_fanout", C.c_int32), ("edge_identity_u8", C.c_int32)]


PARAM_FIELDS = ["lin_event_w", "lin_event_b", "gcn0_w", "gcn0_b", "gcn2_w", "gcn2_b", "att_w1_w", "att_w1_b",
                "att_w2_w", "att_w2_b", "att_mlp0_w", "att_mlp0_b", "att_mlp3_w", "att_mlp3_b", "mlp0_w", "mlp0_b",
                "mlp3_w", "mlp3_b", "mlp5_w", "mlp5_b", "basis_freq", "phase"]


class EncoderParams(C.Structure):
    _fields_ = [(f, _p) for f in PARAM_FIELDS]


class GateDesc(C.Structure):
    _fields_ = [("edge_dim", C.c_int32), ("time_dim", C.c_int32), ("hid_dim", C.c_int32)]


class GateParams(C.Structure):
    _fields_ = [(f, _p) for f in ("w0", "b0", "w3", "b3", "w6", "b6", "basis_freq", "phase")]


SIGNATURES = {
    "tm_version": (C.c_int, []),
    "tm_last_error": (C.c_char_p, []),
    "tm_launch_count": (_u64, []),
    "tm_graph_create": (C.c_int, [_i64, _i64, _p, _p, _p, _p, C.c_int, C.POINTER(_p)]),
    "tm_graph_create_from_events": (C.c_int, [_i64, _i64, _p, _p, _p, _p, C.c_int, C.POINTER(_p)]),
    "tm_graph_destroy": (None, [_p]),
    "tm_graph_sizes": (C.c_int, [_p, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "tm_graph_export": (C.c_int, [_p, _p, _p, _p, _p]),
    "tm_graph_export_edge_table": (C.c_int, [_p, _p]),
    "tm_graph_export_skey": (C.c_int, [_p, _p]),
    "tm_find_before_batch": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "tm_sample_hop": (C.c_int, [_p, _i64, _p, _p, _p, C.c_int, _u64, C.c_uint32, _u64, _p, _p, _p, _p, _p, _p]),
    "tm_sample_khop": (C.c_int, [_p, _i64, C.c_int, C.c_int, _p, _p, _p, _u64, _u64, _p, _p, _p, _p, _p]),
    "tm_sample_walks": (C.c_int, [_p, _i64, C.c_int, C.c_int, _p, _p, _p, _p, _u64, _u64, _p, _p,
                                  _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "tm_walk_final_step": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _u64, _u64, _p, _p, _p, _p, _p, _p]),
    "tm_walk_next_step_time": (C.c_int, [_p, _i64, C.c_int, _p, _p, _p, _u64, _u64, _p, _p, _p, _p, _p, _p, _p]),
    "tm_class_hist": (C.c_int, [_i64, _p, _p, _p, _p, _p, _p]),
    "tm_edge_identity": (C.c_int, [_i64, _i64, _p, _p, _p]),
    "tm_edge_identity_u8": (C.c_int, [_i64, _i64, _p, _p, _p]),
    "tm_encoder_blob_floats": (_i64, [C.POINTER(EncoderDesc)]),
    "tm_encoder_pack": (C.c_int, [C.POINTER(EncoderDesc), C.POINTER(EncoderParams), _p]),
    "tm_encoder_project_edges": (C.c_int, [C.POINTER(EncoderDesc), _p, _p, _i64, _p, C.c_int, _p]),
    "tm_encoder_workspace_floats": (_i64, [C.POINTER(EncoderDesc), _i64, _i64, _i64]),
    "tm_encoder_profile": (C.c_int, [C.c_int]),
    "tm_encoder_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tm_gemm_tf32x3": (C.c_int, [_i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, C.c_int, _p]),
    "tm_selftest_gemm": (C.c_int, [_p, _p, _p, C.c_int, C.c_int, C.c_int, _p]),
    "tm_selftest_cos": (C.c_int, [_p, _p, _i64, _p]),
    "tm_selftest_gather4": (C.c_int, [_p, _i64, C.c_int, _p, C.c_int, C.c_int, _p, _p]),
    "tm_encode_attention": (C.c_int, [C.POINTER(EncoderDesc), _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p,
                                      _p, _i64, _p, _i64, _p, _p, _p, C.c_int, _p]),
    "tm_walk_importance": (C.c_int, [_i64, _i64, _i64, _p, _p, _p, _p, _i64, _p, _p]),
    "tm_enhance_reduce": (C.c_int, [_i64, _i64, C.c_int, _p, _p, _p, _p, _p, _p, _p]),
    "tm_selftest_mma_rate": (C.c_int, [C.c_int, C.c_int, _p, _p]),
    "tm_kl_loss": (C.c_int, [_i64, _i64, _p, _p, _p, C.c_int, C.c_float, C.c_int, _p, _p, _p]),
    "tm_gate_blob_floats": (_i64, [C.POINTER(GateDesc)]),
    "tm_gate_pack": (C.c_int, [C.POINTER(GateDesc), C.POINTER(GateParams), _p]),
    "tm_edge_importance": (C.c_int, [C.POINTER(GateDesc), _p, _i64, _i64, _p, _p, _p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p, _p,
                                     C.c_int, _u64, C.c_int, _p]),
    "tm_beta_sample": (C.c_int, [_i64, _p, _p, _u64, _u64, _p, _p, _p, _p]),
    "tm_kl_loss_backward": (C.c_int, [_i64, _i64, _p, _p, _p, C.c_int, C.c_float, C.c_int, _p, _p, _p]),
    "tm_encode_score_gather": (C.c_int, [C.POINTER(EncoderDesc), _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p,
                                         _p, _i64, _p, _i64, _p, _p, C.POINTER(C.c_uint64), C.c_int, C.c_int, _p]),
    "tm_encode_score": (C.c_int, [C.POINTER(EncoderDesc), _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p,
                                  _p, _i64, _p, _i64, _p, _p, C.c_int, _p]),
}

_lib = None


class TempMEError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if _build.stale():
            if _build.nvcc_path() is None and not os.path.exists(path):
                raise ImportError("tempme_b200: csrc/libtempme_b200.so is missing and nvcc is not available; "
                                  "there is no CPU fallback. Run `python -m tempme_b200.build`.")
            if _build.nvcc_path() is not None:
                _build.build()
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = the .so does not match include/tempme_b200.h
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().tm_last_error().decode(errors="replace")
        if rc == TM_ERR_UNSUPPORTED:
            raise NotImplementedError(f"{what}: {msg}")
        raise TempMEError(f"{what} failed ({rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return t.ctypes.data_as(C.c_void_p)


def launch_count() -> int:
    return int(lib().tm_launch_count())

"""ctypes binding of libarx.so (include/arx.h).  No CPU fallback: if the CUDA
library is missing or fails to load, importing the scoring path raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libarx.so")
ARX_MAX_TRANSFORMERS = 4
ABI_VERSION = 1

EXPORTS = [
    "arx_create", "arx_destroy", "arx_last_error", "arx_abi_version", "arx_load_weights", "arx_tuple_count",
    "arx_tuple_table", "arx_embed", "arx_set_support_poses", "arx_set_support_features",
    "arx_get_support_features", "arx_support_way", "arx_support_blob_bytes", "arx_export_support",
    "arx_import_support", "arx_score", "arx_score_features", "arx_debug_attention", "arx_score_host",
    "arx_score_episodes", "arx_score_frames", "arx_stream_push", "arx_stream_reset", "arx_score_host_submit", "arx_score_host_wait", "arx_score_host_f16", "arx_score_host_submit_f16", "arx_decode_heatmaps", "arx_decode_heatmaps_cams", "arx_heads_load", "arx_heads_forward", "arx_launch_count", "arx_last_path", "arx_profile_enable", "arx_profile_read", "arx_debug_set", "arx_debug_read_trace",
]


class ArxConfig(C.Structure):
    _fields_ = [("seq_len", C.c_int32), ("n_joints", C.c_int32), ("feat_dim", C.c_int32), ("out_dim", C.c_int32),
                ("n_transformers", C.c_int32), ("cardinality", C.c_int32 * ARX_MAX_TRANSFORMERS),
                ("has_discriminator", C.c_int32), ("max_chunk", C.c_int32), ("force_path", C.c_int32)]


_FP = C.c_void_p


class ArxWeights(C.Structure):
    _fields_ = [("on_device", C.c_int32),
                ("fc1_w", _FP), ("fc1_b", _FP), ("fc2_w", _FP), ("fc2_b", _FP),
                ("pe", _FP * ARX_MAX_TRANSFORMERS), ("k_w", _FP * ARX_MAX_TRANSFORMERS),
                ("k_b", _FP * ARX_MAX_TRANSFORMERS), ("v_w", _FP * ARX_MAX_TRANSFORMERS),
                ("v_b", _FP * ARX_MAX_TRANSFORMERS), ("ln_g", _FP * ARX_MAX_TRANSFORMERS),
                ("ln_b", _FP * ARX_MAX_TRANSFORMERS),
                ("dr_w", _FP), ("dr_b", _FP), ("d1_w", _FP), ("d1_b", _FP), ("d2_w", _FP), ("d2_b", _FP),
                ("d3_w", _FP), ("d3_b", _FP)]


class ArxError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libarx.so from the package directory; raise loudly if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m isbfsar_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback for the scoring path.")
    lib = C.CDLL(LIB_PATH)
    H, I32, I64, VP = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    sig = {
        "arx_create": (C.c_int, [C.POINTER(ArxConfig), C.POINTER(H)]),
        "arx_destroy": (None, [H]),
        "arx_last_error": (C.c_char_p, [H]),
        "arx_abi_version": (C.c_int, []),
        "arx_load_weights": (C.c_int, [H, C.POINTER(ArxWeights), VP]),
        "arx_tuple_count": (C.c_int, [H, I32]),
        "arx_tuple_table": (C.c_int, [H, I32, VP, VP]),
        "arx_embed": (C.c_int, [H, VP, I64, VP, VP]),
        "arx_set_support_poses": (C.c_int, [H, VP, I32, VP]),
        "arx_set_support_features": (C.c_int, [H, VP, I32, VP]),
        "arx_get_support_features": (C.c_int, [H, VP, VP]),
        "arx_support_way": (C.c_int, [H]),
        "arx_support_blob_bytes": (I64, [H, I32]),
        "arx_export_support": (C.c_int, [H, VP, VP]),
        "arx_import_support": (C.c_int, [H, VP, I32, VP]),
        "arx_score": (C.c_int, [H, VP, I64, VP, VP, VP, VP]),
        "arx_score_features": (C.c_int, [H, I32, VP, I64, VP, VP]),
        "arx_score_frames": (C.c_int, [H, VP, I64, VP, VP, VP, VP]),
        "arx_score_episodes": (C.c_int, [H, VP, I32, I32, VP, I64, VP, VP, VP, VP]),
        "arx_debug_attention": (C.c_int, [H, VP, I64, VP, VP, VP]),
        "arx_score_host": (C.c_int, [H, VP, I64, VP, VP, VP]),
        "arx_stream_push": (C.c_int, [H, VP, VP, C.POINTER(I32)]),
        "arx_stream_reset": (C.c_int, [H]),
        "arx_score_host_submit": (C.c_int, [H, VP, I64, VP, VP, VP, C.POINTER(I64)]),
        "arx_score_host_wait": (C.c_int, [H, I64]),
        "arx_score_host_f16": (C.c_int, [H, VP, I64, VP, VP, VP]),
        "arx_score_host_submit_f16": (C.c_int, [H, VP, I64, VP, VP, VP, C.POINTER(I64)]),
        "arx_decode_heatmaps": (C.c_int, [H, VP, I64, VP, I32, VP, VP, VP, VP, VP]),
        "arx_decode_heatmaps_cams": (C.c_int, [H, VP, I64, VP, I32, VP, VP, VP, VP, VP]),
        "arx_heads_load": (C.c_int, [H, VP, VP, I32, VP]),
        "arx_heads_forward": (C.c_int, [H, VP, I64, VP, VP]),
        "arx_launch_count": (I64, [H]),
        "arx_last_path": (C.c_int, [H]),
        "arx_debug_set": (C.c_int, [H, I32, I32]),
        "arx_debug_read_trace": (C.c_int, [H, C.POINTER(C.c_longlong)]),
        "arx_profile_enable": (C.c_int, [H, I32]),
        "arx_profile_read": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(I64), I32]),
    }
    for name in EXPORTS:
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = sig[name]
    if lib.arx_abi_version() != ABI_VERSION:
        raise ImportError(f"libarx ABI {lib.arx_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, handle=None, what: str = "") -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.arx_last_error(handle)
    msg = msg.decode() if msg else ""
    names = {-1: "ARX_ERR_INVALID", -2: "ARX_ERR_CUDA", -3: "ARX_ERR_STATE", -4: "ARX_ERR_NOMEM"}
    err = f"{what}: {names.get(rc, rc)}: {msg}"
    if rc == -1:
        raise ValueError(err)
    raise ArxError(err)

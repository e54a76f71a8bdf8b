"""ctypes binding of libdfs_b200.so (include/dfs_b200.h).

The library is the product: there is no Python / CPU fallback.  If the shared object is missing
or a symbol cannot be resolved, importing callers fail with an explicit error.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdfs_b200.so")

T_FRAMES, N_FEATS = 321, 180


class NativeLibraryError(RuntimeError):
    pass


class ConvBn(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_float)) for k in ("weight", "bias", "bn_weight", "bn_bias", "bn_mean", "bn_var")]


class Cnn2dWeights(C.Structure):
    _fields_ = [("in_features", C.c_int), ("base_channels", C.c_int), ("conv", ConvBn * 3),
                ("fc_weight", C.POINTER(C.c_float)), ("fc_bias", C.POINTER(C.c_float))]


class Cnn1dWeights(C.Structure):
    _fields_ = [("in_features", C.c_int), ("base_channels", C.c_int), ("conv", ConvBn * 3),
                ("fc_weight", C.POINTER(C.c_float)), ("fc_bias", C.POINTER(C.c_float))]


class CaeWeights(C.Structure):
    _fields_ = [("base_channels", C.c_int), ("enc", ConvBn * 4), ("dec", ConvBn * 4),
                ("norm_mean", C.POINTER(C.c_float)), ("norm_std", C.POINTER(C.c_float))]


class DlqWeights(C.Structure):
    _fields_ = [("in_ch", C.c_int), ("hidden", C.c_int), ("conv", ConvBn * 3),
                ("fc1_weight", C.POINTER(C.c_float)), ("fc1_bias", C.POINTER(C.c_float)),
                ("fc2_weight", C.POINTER(C.c_float)), ("fc2_bias", C.POINTER(C.c_float))]


class Features(C.Structure):
    _fields_ = [("x", C.c_void_p), ("n", C.c_int64), ("stride_n", C.c_int64), ("stride_t", C.c_int64), ("stride_f", C.c_int64)]


class EerResult(C.Structure):
    _fields_ = [("eer", C.c_double), ("threshold", C.c_double), ("eer_idx", C.c_int64),
                ("n_bonafide", C.c_int64), ("n_spoof", C.c_int64)]


# name -> (restype, argtypes); every symbol include/dfs_b200.h declares
SIGNATURES = {
    "dfs_version": (C.c_int, []),
    "dfs_last_error": (C.c_char_p, []),
    "dfs_launch_count": (C.c_int64, []),
    "dfs_cnn2d_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Cnn2dWeights), C.c_int]),
    "dfs_cnn1d_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Cnn1dWeights), C.c_int]),
    "dfs_cae_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(CaeWeights), C.c_int]),
    "dfs_dlq_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(DlqWeights), C.c_int]),
    "dfs_dlq_score": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dfs_model_destroy": (C.c_int, [C.c_void_p]),
    "dfs_model_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "dfs_model_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "dfs_model_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int, C.c_int]),
    "dfs_model_saturation_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
    "dfs_cnn2d_score": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "dfs_cnn1d_score": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_void_p, C.c_int, C.c_void_p]),
    "dfs_cae_score": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_cae_forward": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfs_cae_debug_layer": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_score_host": (C.c_int, [C.c_void_p, C.POINTER(Features), C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_score_host_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_group_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int]),
    "dfs_group_destroy": (C.c_int, [C.c_void_p]),
    "dfs_group_stage_utts": (C.c_int64, [C.c_void_p]),
    "dfs_group_score_host": (C.c_int, [C.c_void_p, C.POINTER(Features), C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p]),
    "dfs_group_score_host_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_void_p]),
    "dfs_pinned_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "dfs_pinned_free": (C.c_int, [C.c_void_p]),
    "dfs_blend_f64": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_double,
                                C.c_int64, C.c_void_p, C.c_void_p]),
    "dfs_widen_f32_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "dfs_eer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(EerResult), C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfs_eer_select": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(EerResult), C.c_void_p]),
    "dfs_bce_with_logits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_void_p]),
    "dfs_set_global_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "dfs_confusion": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.POINTER(C.c_int64), C.c_void_p]),
    "dfs_fill_features": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_uint64, C.c_float, C.c_void_p]),
}

_lib = None


def load():
    """Load (once) and return the ctypes library with all prototypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python deep-fake-audio-classifier_b200/build.py` "
            "(or __graft_entry__.build()).  dfs_b200 has no Python/CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "dfs_b200"):
    if status != 0:
        msg = load().dfs_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


def launch_count() -> int:
    return int(load().dfs_launch_count())


def set_global_option(key: str, value: int):
    check(load().dfs_set_global_option(key.encode(), int(value)), "dfs_set_global_option")

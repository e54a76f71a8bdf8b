"""ctypes binding of lib/libdfs_b200_probes.so (csrc/probes.h): bring-up probes and micro-benchmarks.
Test / measurement code only -- the scoring path never loads this library."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libdfs_b200_probes.so")

SIGNATURES = {
    "dfs_probe_last_error": (C.c_char_p, []),
    "dfs_probe_umma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_probe_tma_window": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dfs_probe_umma_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int64), C.c_void_p]),
    "dfs_probe_tmem_ld_bench": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python deep-fake-audio-classifier_b200/build.py`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int, what: str = "dfs_probe"):
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}): {load().dfs_probe_last_error().decode(errors='replace')}")

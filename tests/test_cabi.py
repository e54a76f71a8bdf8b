"""CPU gate: the C-ABI library loads without a GPU and exports every symbol include/dfs_b200.h
declares; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from dfs_b200 import _native as N


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dfs_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dfs_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exists_and_loads():
    assert os.path.exists(N.LIB_PATH), "run __graft_entry__.build() first"
    lib = N.load()
    assert lib.dfs_version() >= 100


def test_every_declared_symbol_is_exported_and_bound():
    lib = C.CDLL(N.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in dfs_b200.h but not exported"
        assert name in N.SIGNATURES, f"{name} has no ctypes prototype in _native.py"
    assert sorted(N.SIGNATURES) == declared


def test_argument_validation_without_gpu():
    lib = N.load()
    h = C.c_void_p()
    w = N.Cnn2dWeights()
    w.in_features, w.base_channels = 200, 32
    assert lib.dfs_cnn2d_create(C.byref(h), 0, C.byref(w), 0) == -3          # DFS_ERR_UNSUPPORTED
    assert b"in_features=180" in lib.dfs_last_error()
    w.in_features = 180
    assert lib.dfs_cnn2d_create(C.byref(h), 0, C.byref(w), 0) == -1          # NULL tensors
    assert lib.dfs_model_destroy(None) == 0
    res = N.EerResult()
    assert lib.dfs_eer(None, 4, None, 10, C.byref(res), None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dfs_b200 import Cnn2dScorer, calculate_eer, synthetic
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Cnn2dScorer(synthetic.cnn2d_state(0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        calculate_eer([0.1, 0.9], [0, 1])


def test_every_option_key_is_documented_in_the_header():
    """Every key dfs_model_set_option / dfs_set_global_option accepts (csrc/api.cu, csrc/eer.cu) is described in include/dfs_b200.h."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    keys = set()
    for src in ("api.cu", "eer.cu"):
        text = open(os.path.join(root, "deep-fake-audio-classifier_b200", "csrc", src)).read()
        keys |= set(re.findall(r'strcmp\(key, "([a-z0-9_]+)"\)', text))
    header = open(os.path.join(root, "include", "dfs_b200.h")).read()
    assert keys and all(f'"{k}"' in header for k in keys), sorted(k for k in keys if f'"{k}"' not in header)


def test_global_options_validate_their_arguments_without_gpu():
    """dfs_set_global_option touches no device: keys and ranges are checked on the host."""
    lib = N.load()
    assert lib.dfs_set_global_option(b"eer_sort_overlap", 0) == 0 and lib.dfs_set_global_option(b"eer_sort_overlap", 1) == 0
    assert lib.dfs_set_global_option(b"eer_sort_onesweep", 6) == -1 and b"eer_sort_onesweep" in lib.dfs_last_error()
    assert lib.dfs_set_global_option(b"eer_sort_onesweep", 1) == 0
    assert lib.dfs_set_global_option(b"no_such_switch", 1) == -1 and b"unknown key" in lib.dfs_last_error()
    assert lib.dfs_set_global_option(None, 1) == -1

"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/ukbb_fcn.h declares (no compute calls here)."""
import ctypes
import os
import re

from ukbb_cardiac_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ukbb_fcn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ukbb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    syms = declared_symbols()
    assert "ukbb_fcn_create" in syms and "ukbb_fcn_forward" in syms and "ukbb_fcn_preprocess" in syms
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libukbb_fcn.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"


def test_error_convention_without_compute():
    lib = _lib.load()
    assert lib.ukbb_version().startswith(b"ukbb_fcn")
    out = ctypes.c_void_p()
    rc = lib.ukbb_fcn_create(None, 4, 0, _lib.MODE_FP32, ctypes.byref(out))
    assert rc == -1 and b"null" in lib.ukbb_last_error()
    fw = _lib.FcnWeights(3, None, 1e-3)
    rc = lib.ukbb_fcn_create(ctypes.byref(fw), 4, 0, _lib.MODE_FP32, ctypes.byref(out))
    assert rc == -1 and b"21" in lib.ukbb_last_error()
    assert lib.ukbb_fcn_launch_count(None) == 0
    lib.ukbb_fcn_destroy(None)


def test_struct_layout_matches_header():
    # ukbb_conv_weights: pointer, 4 ints, 5 pointers (LP64)
    assert ctypes.sizeof(_lib.ConvWeights) == 8 + 16 + 5 * 8
    assert ctypes.sizeof(_lib.FcnWeights) == 24

"""ctypes binding of libukbb_fcn.so (the C ABI declared in include/ukbb_fcn.h).

There is no CPU fallback: if the library cannot be loaded the import of the
engine fails loudly.  The library itself loads without a GPU (CUDA runtime is
linked statically, the driver is resolved lazily), so symbol checks run on CPU.
"""
from __future__ import annotations

import ctypes as C
import os

from . import tf_bundle

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libukbb_fcn.so")

MODE_FP32 = 0
MODE_BF16 = 2
MODE_FP16 = 3
MODE_BF16X3 = 4     # split operands (hi + lo 16-bit pairs, three tcgen05.mma per K step): 16-bit significand
MODE_FP16X3 = 5     # same with FP16 pieces: 22-bit significand -- the default tensor-core mode
MODE_FP16X2 = 6     # FP16 pieces, both correction products as one FP8 (E4M3) instruction: two tcgen05.mma per K step
N_CONV = 21
MAX_CLASS = 8


class ConvWeights(C.Structure):
    _fields_ = [("kernel", C.POINTER(C.c_float)),
                ("ksize", C.c_int), ("cin", C.c_int), ("cout", C.c_int), ("stride", C.c_int),
                ("gamma", C.POINTER(C.c_float)), ("beta", C.POINTER(C.c_float)),
                ("moving_mean", C.POINTER(C.c_float)), ("moving_variance", C.POINTER(C.c_float)),
                ("bias", C.POINTER(C.c_float))]


class FcnWeights(C.Structure):
    _fields_ = [("n_conv", C.c_int), ("conv", C.POINTER(ConvWeights)), ("bn_eps", C.c_float)]


# name -> (restype, argtypes); must list every symbol of include/ukbb_fcn.h
SIGNATURES = {
    "ukbb_fcn_create": (C.c_int, [C.POINTER(FcnWeights), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ukbb_fcn_destroy": (None, [C.c_void_p]),
    "ukbb_fcn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ukbb_fcn_preprocess": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_void_p]),
    "ukbb_fcn_rescale": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ukbb_fcn_segment_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ukbb_fcn_class_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ukbb_fcn_join": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ukbb_fcn_sync": (C.c_int, [C.c_void_p]),
    "ukbb_fcn_debug_conv": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "ukbb_fcn_debug_read": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
    "ukbb_fcn_debug_flags": (C.c_int, [C.c_void_p, C.c_int]),
    "ukbb_fcn_kernel_timer": (C.c_int, [C.c_void_p, C.c_int]),
    "ukbb_fcn_kernel_timer_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "ukbb_fcn_launch_count": (C.c_longlong, [C.c_void_p]),
    "ukbb_fcn_mode": (C.c_int, [C.c_void_p]),
    "ukbb_fcn_n_class": (C.c_int, [C.c_void_p]),
    "ukbb_cc_stats": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "ukbb_ao_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ukbb_ao_destroy": (None, [C.c_void_p]),
    "ukbb_ao_segment": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "ukbb_ao_launch_count": (C.c_longlong, [C.c_void_p]),
    "ukbb_crc32c": (C.c_uint32, [C.c_void_p, C.c_size_t]),
    "ukbb_last_error": (C.c_char_p, []),
    "ukbb_version": (C.c_char_p, []),
}

_lib = None


class UkbbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libukbb_fcn error %d: %s" % (code, msg))
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m ukbb_cardiac_b200.build` (nvcc, sm_100a). "
            "This package has no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError = stale library: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    tf_bundle.set_native_crc32c(lambda data: int(lib.ukbb_crc32c(bytes(data), len(data))))
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise UkbbError(rc, (load().ukbb_last_error() or b"").decode("utf-8", "replace"))

"""Loads liblbm_b200.so (the C-ABI of include/lbm_b200.h) with ctypes.

There is no fallback of any kind: if the library was not built, or a compute call finds no
CUDA device, the caller gets an exception."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "liblbm_b200.so")
EXE_PATH = os.path.join(HERE, "bin", "d2q9-bgk")

c_int_p = ctypes.POINTER(ctypes.c_int)
c_float_p = ctypes.POINTER(ctypes.c_float)
c_long_p = ctypes.POINTER(ctypes.c_long)
handle_t = ctypes.c_void_p

# name -> (restype, argtypes): every symbol include/lbm_b200.h declares
SIGNATURES = {
    "lbm_b200_abi_version": (ctypes.c_int, []),
    "lbm_b200_decompose": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_int_p, c_int_p]),
    "lbm_b200_plan_bands": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_int_p, c_int_p]),
    "lbm_b200_plan_bands_ex": (ctypes.c_int, [ctypes.c_int] * 6 + [c_int_p, c_int_p]),
    "lbm_b200_free_cells_inv": (ctypes.c_float, [c_int_p, ctypes.c_long]),
    "lbm_b200_last_error": (ctypes.c_char_p, []),
    "lbm_b200_device_count": (ctypes.c_int, []),
    "lbm_b200_create": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                       ctypes.c_float, ctypes.c_float, c_int_p, ctypes.c_int, c_int_p]),
    "lbm_b200_create_inplace": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_float, c_int_p, ctypes.c_int, c_int_p]),
    "lbm_b200_create_slab_inplace": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                    c_int_p, ctypes.c_int]),
    "lbm_b200_create_slab": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_float, c_int_p, ctypes.c_int]),
    "lbm_b200_create_ex": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_int_p,
                                          ctypes.c_int]),
    "lbm_b200_create_slab_ex": (ctypes.c_int, [ctypes.POINTER(handle_t), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                               ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int]),
    "lbm_b200_selftest": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]),
    "lbm_b200_ipc_blob_bytes": (ctypes.c_int, []),
    "lbm_b200_ipc_export": (ctypes.c_int, [handle_t, ctypes.c_void_p]),
    "lbm_b200_ipc_connect": (ctypes.c_int, [handle_t, ctypes.c_void_p, ctypes.c_void_p]),
    "lbm_b200_run": (ctypes.c_int, [handle_t, ctypes.c_int, c_float_p]),
    "lbm_b200_enqueue": (ctypes.c_int, [handle_t, ctypes.c_int]),
    "lbm_b200_sync": (ctypes.c_int, [handle_t]),
    "lbm_b200_elapsed_ms": (ctypes.c_int, [handle_t, c_float_p]),
    "lbm_b200_fetch_av_vels": (ctypes.c_int, [handle_t, ctypes.c_int, c_float_p]),
    "lbm_b200_shape": (ctypes.c_int, [handle_t, c_int_p, c_int_p, c_int_p]),
    "lbm_b200_get_cells": (ctypes.c_int, [handle_t, c_float_p]),
    "lbm_b200_set_cells": (ctypes.c_int, [handle_t, c_float_p]),
    "lbm_b200_get_final_state": (ctypes.c_int, [handle_t, c_float_p, c_float_p, c_float_p, c_float_p]),
    "lbm_b200_set_option": (ctypes.c_int, [handle_t, ctypes.c_char_p, ctypes.c_long]),
    "lbm_b200_get_option": (ctypes.c_int, [handle_t, ctypes.c_char_p, c_long_p]),
    "lbm_b200_destroy": (None, [handle_t]),
}

_lib = None


def library() -> ctypes.CDLL:
    """dlopen()s the in-tree library and declares every signature; raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built (run `make -C {HERE}` or __graft_entry__.build()); "
                "the CUDA library is the only implementation -- there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib

"""ctypes binding of libtheatergen.so (the C-ABI in include/theatergen.h).

The product path fails loudly when the library is missing -- there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtheatergen.so")
_lib = None

vp, ci, cf = C.c_void_p, C.c_int, C.c_float

_SIGS = {
    "tg_last_error": (C.c_char_p, []),
    "tg_version": (ci, []),
    "tg_launch_count": (C.c_ulonglong, []),
    "tg_gemm": (ci, [vp, ci, ci, vp, ci, vp, ci, ci, ci, vp, vp, ci, ci, vp, ci, ci, cf, vp, ci, vp]),
    "tg_conv3x3": (ci, [vp, ci, ci, ci, ci, vp, ci, vp, vp, ci, vp, ci, cf, vp, ci, vp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m theatergen_b200.build` "
                "(theatergen_b200 has no CPU / PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS.keys())


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().tg_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"theatergen {what} failed ({rc}): {msg}")


def launch_count() -> int:
    return int(lib().tg_launch_count())

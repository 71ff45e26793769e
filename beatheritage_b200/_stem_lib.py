"""ctypes binding of libbhstem.so (the C ABI declared in include/bhstem.h).

Built in-tree by `beatheritage_b200.build`; there is NO CPU fallback: if the library is missing
and cannot be built, `lib()` raises."""
from __future__ import annotations

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BHSTEM_LIB") or os.path.join(_PKG, "libbhstem.so")
OK = 0

_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); tests check this list against include/bhstem.h
SIGNATURES = {
    "bhstem_create": (ctypes.c_int, [_i32, _i32, _fp, _fp, _fp, _fp, ctypes.POINTER(_vp)]),
    "bhstem_destroy": (None, [_vp]),
    "bhstem_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "bhstem_forward_stage": (ctypes.c_int, [_vp, _i32, _vp, _i64, _i64, _vp, _vp]),
    "bhstem_prepare_split": (ctypes.c_int, [_vp, _i32]),
    "bhstem_forward_split": (ctypes.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "bhstem_set_option": (ctypes.c_int, [_vp, _i32, _i64]),
    "bhstem_version": (ctypes.c_int, []),
    "bhstem_last_error": (ctypes.c_char_p, []),
    "bhstem_launch_count": (_i64, [_vp]),
}


class BhstemError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbhstem error {code}: {message}")
        self.code = code


_lock = threading.Lock()
_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build_stem()
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise BhstemError(rc, lib().bhstem_last_error().decode("utf-8", "replace"))

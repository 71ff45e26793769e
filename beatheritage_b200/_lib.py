"""ctypes binding of libbhmel.so (the C ABI declared in include/bhmel.h).

The library is built in-tree by `beatheritage_b200.build`; there is NO CPU fallback: if it is
missing and cannot be built, importing this module's `lib()` raises."""
from __future__ import annotations

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BHMEL_LIB") or os.path.join(_PKG, "libbhmel.so")   # BHMEL_LIB: A/B builds

PAD_CONSTANT, PAD_REFLECT = 0, 1
OPT_BULK_COPY = 1
OPT_KERNEL = 2
OPT_STATIC_MEL = 3
OPT_PDL = 4
KERNEL_BARRIER, KERNEL_INDEPENDENT_WARPS, KERNEL_WARP_SPECIALIZED = 0, 1, 2
OK, EINVAL, ECUDA, ESHAPE, EDEVICE = 0, 1, 2, 3, 4


class BhmelParams(ctypes.Structure):
    _fields_ = [
        ("sample_rate", ctypes.c_int32),
        ("n_fft", ctypes.c_int32),
        ("hop_length", ctypes.c_int32),
        ("n_mels", ctypes.c_int32),
        ("f_min", ctypes.c_double),
        ("f_max", ctypes.c_double),
        ("pad_mode", ctypes.c_int32),
        ("log_scale", ctypes.c_int32),
        ("fb", ctypes.POINTER(ctypes.c_float)),
        ("window", ctypes.POINTER(ctypes.c_float)),
    ]


class BhmelOutDesc(ctypes.Structure):
    _fields_ = [("y", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("frame_pitch", ctypes.c_int64),
                ("row_pitch", ctypes.c_int64)]


OUT_F32, OUT_BF16 = 0, 1
LAYOUT_BTC, LAYOUT_BCT = 0, 1


class BhmelEncoderInputDesc(ctypes.Structure):
    _fields_ = [("y", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("layout", ctypes.c_int32),
                ("cond", ctypes.c_void_p), ("n_cond", ctypes.c_int64), ("scratch", ctypes.c_void_p)]
IN_F32, IN_PCM16 = 0, 1


class BhmelHostIO(ctypes.Structure):
    _fields_ = [("x_host", ctypes.c_void_p), ("x_dtype", ctypes.c_int32), ("scales", ctypes.c_void_p),
                ("y_host", ctypes.c_void_p), ("y_dtype", ctypes.c_int32)]


class BhmelError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libbhmel error {code}: {message}")
        self.code = code


_lock = threading.Lock()
_lib = None

_vp, _i64, _i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); also the list tests check against include/bhmel.h
SIGNATURES = {
    "bhmel_create": (ctypes.c_int, [ctypes.POINTER(BhmelParams), ctypes.POINTER(_vp)]),
    "bhmel_destroy": (None, [_vp]),
    "bhmel_set_fb": (ctypes.c_int, [_vp, _fp]),
    "bhmel_set_window": (ctypes.c_int, [_vp, _fp]),
    "bhmel_get_fb": (ctypes.c_int, [_vp, _fp]),
    "bhmel_get_window": (ctypes.c_int, [_vp, _fp]),
    "bhmel_num_frames": (_i64, [_vp, _i64]),
    "bhmel_forward": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "bhmel_forward_ex": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, ctypes.POINTER(BhmelOutDesc), _vp]),
    "bhmel_forward_encoder_input": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, ctypes.POINTER(BhmelEncoderInputDesc), _vp]),
    "bhmel_forward_gather": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp]),
    "bhmel_peak_scale_pcm16": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "bhmel_forward_gather_pcm16": (ctypes.c_int, [_vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "bhmel_forward_host": (ctypes.c_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "bhmel_forward_host_ex": (ctypes.c_int, [_vp, ctypes.POINTER(BhmelHostIO), _i64, _i64, _i64]),
    "bhmel_host_chunk_plan": (_i64, [_i64, _i64, _i32, ctypes.POINTER(_i64), _i64]),
    "bhmel_set_option": (ctypes.c_int, [_vp, _i32, _i64]),
    "bhmel_version": (ctypes.c_int, []),
    "bhmel_last_error": (ctypes.c_char_p, []),
    "bhmel_launch_count": (_i64, [_vp]),
    "bhmel_kernel_info": (None, [ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
}


def lib() -> ctypes.CDLL:
    """Load (building first if the .so is absent) and return the bound library."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build()
            handle = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
            _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise BhmelError(rc, lib().bhmel_last_error().decode("utf-8", "replace"))

"""ctypes binding of ``csrc/librf_b200.so`` (C ABI: ``include/rf_b200.h``).

The library is built in-tree by ``csrc/build.sh`` (see ``__graft_entry__.build``).  There is no
CPU fallback: :func:`lib` raises if the shared object is missing, and every compute wrapper
raises if the call does not return ``RF_OK``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librf_b200.so")

RF_OK, RF_EINVAL, RF_EUNSUPPORTED, RF_ECUDA, RF_ENOMEM = 0, 1, 2, 3, 4
RF_BF_GRAY_REPLICATED = 1
# cv::BorderTypes as rf_joint_bilateral_u8_border / rf_joint_bilateral_f32 take them (cv2.BORDER_*)
RF_BORDER_CONSTANT, RF_BORDER_REPLICATE, RF_BORDER_REFLECT, RF_BORDER_WRAP, RF_BORDER_REFLECT_101 = 0, 1, 2, 3, 4
RF_WHDR_PIXEL_COORDS = 1

_lib = None

# every symbol include/rf_b200.h declares: (restype, argtypes)
_vp, _sz, _i, _d, _u = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_uint
_ip = C.POINTER(C.c_int)
SIGNATURES = {
    "rf_version": (_i, []),
    "rf_last_error": (C.c_char_p, []),
    "rf_set_device": (_i, [_i]),
    "rf_device_info": (_i, [_ip, _ip, _ip, C.POINTER(_sz)]),
    "rf_launch_count": (C.c_ulonglong, []),
    "rf_cnn_create": (_i, [_vp, _vp, _i, _vp, C.POINTER(_vp)]),
    "rf_cnn_destroy": (None, [_vp]),
    "rf_cnn_forward_u8": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "rf_joint_bilateral_u8": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _i, _u, _vp]),
    "rf_joint_bilateral_u8_border": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _i, _u, _i, _vp]),
    "rf_joint_bilateral_f32_workspace_bytes": (_sz, [_i, _i]),
    "rf_joint_bilateral_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _d, _d, _i, _i, _vp, _sz, _vp]),
    "rf_joint_bilateral_geometry": (_i, [_d, _i, _ip, _ip]),
    "rf_joint_bilateral_max_radius": (_i, []),
    "rf_joint_bilateral_fast_max_radius": (_i, []),
    "rf_guided_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "rf_guided_u8": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _d, _vp, _sz, _vp]),
    "rf_guided_max_radius": (_i, []),
    "rf_guided_f32_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "rf_guided_f32": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _d, _vp, _sz, _vp]),
    "rf_guided_iterated_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "rf_guided_iterated_u8": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _d, _i, _vp, _sz, _vp]),
    "rf_guided_row_terms": (_i, [_i, _i, _i, _i, _ip, C.POINTER(C.c_float)]),
    "rf_replicate_gray_u8": (_i, [_vp, _vp, _sz, _vp]),
    "rf_extract_gray_u8": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "rf_accumulate_stats_u8": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "rf_whdr_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _d, _u, _vp, _vp, _vp]),
    "rf_colorize_workspace_bytes": (_sz, [_i, _i, _i]),
    "rf_colorize_u8": (_i, [_vp, _vp, _i, _i, _i, _d, C.c_ulonglong, C.c_ulonglong, _vp, _vp, _vp, _sz, _vp]),
}


class NativeError(RuntimeError):
    """A call into librf_b200.so returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__("librf_b200: %s (status %d)" % (message, status))
        self.status = status


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "%s is missing: build it with reflectance-filtering_b200/csrc/build.sh "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != RF_OK:
        msg = lib().rf_last_error()
        raise NativeError(status, msg.decode("utf-8", "replace") if msg else "unknown error")


def launch_count() -> int:
    return int(lib().rf_launch_count())

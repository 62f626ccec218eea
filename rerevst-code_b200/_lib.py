"""ctypes binding of ``csrc/librerevst_b200.so`` (C ABI declared in ``include/rerevst_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRV_LIB_PATH") or os.path.join(_HERE, "csrc", "librerevst_b200.so")   # override: kernel experiments

OUT_PLANES, OUT_F32_NHWC, OUT_F32_NCHW, OUT_BGR_F32, OUT_BGR_U8 = 0, 1, 2, 3, 4
TERMS_FULL, TERMS_NO_WLO, TERMS_NO_ALO = 0, 1, 2
ABI_VERSION = 3
IMPL_FFMA, IMPL_TCGEN05 = 0, 1

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64


class Epilogue(C.Structure):
    """``rrv_epilogue`` (include/rerevst_b200.h)."""
    _fields_ = [("bias", _vp), ("act", _i32), ("norm1", _vp), ("res_hi", _vp), ("res_lo", _vp),
                ("res_shift", _i32), ("res_H", _i32), ("res_W", _i32), ("res_batch_stride", _i64),
                ("norm2", _vp), ("affine", _vp), ("res_f32", _i32)]


class Conv(C.Structure):
    """``rrv_conv`` (include/rerevst_b200.h)."""
    _fields_ = [("N", _i32), ("H", _i32), ("W", _i32), ("Cin", _i32), ("Cout", _i32), ("ksize", _i32),
                ("ups", _i32), ("in_hi", _vp), ("in_lo", _vp), ("w_f32", _vp), ("w_tc", _vp),
                ("ep", Epilogue), ("out_mode", _i32), ("out_hi", _vp), ("out_lo", _vp), ("out_f32", _vp),
                ("out_C", _i32), ("Cin_used", _i32), ("pool", _i32), ("terms", _i32), ("out_img", _vp),
                ("crop_y0", _i32), ("crop_x0", _i32), ("crop_h", _i32), ("crop_w", _i32), ("stats", _vp), ("stats_minmax", _i32)]


# name -> (restype, argtypes); must list every symbol of include/rerevst_b200.h
SIGNATURES = {
    "rrv_abi_version": (C.c_int, []),
    "rrv_last_error": (C.c_char_p, []),
    "rrv_launch_count": (C.c_uint64, []),
    "rrv_set_lo_format": (C.c_int, [C.c_int]),
    "rrv_get_lo_format": (C.c_int, []),
    "rrv_conv2d": (C.c_int, [C.POINTER(Conv), C.c_int, _vp]),
    "rrv_tc_weight_bytes": (_i64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "rrv_pack_weights_tc": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_tc_tune": (C.c_int, [C.c_int, C.c_int]),
    "rrv_tc_tune_pair": (C.c_int, [C.c_int, C.c_int]),
    "rrv_tc_tune_merge": (C.c_int, [C.c_int]),
    "rrv_tc_tune_pdl": (C.c_int, [C.c_int]),
    "rrv_tc_timeline": (C.c_int, [_vp, C.c_int]),
    "rrv_pack_weights_f32": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_relu_backward": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "rrv_maxpool2x2_backward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_fold_filter": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rrv_first_layer": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rrv_maxpool2x2": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rrv_pointwise": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), C.c_int,
                                _vp, _vp, _vp, _vp]),
    "rrv_planes_to_nchw": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_nchw_to_planes": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rrv_reflect_pad_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_postprocess_bgr": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_postprocess_bgr_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "rrv_channel_stats": (C.c_int, [_vp, _i64, C.c_int, _vp, _vp]),
    "rrv_conv3x3_output_sum": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rrv_stats_init": (C.c_int, [_vp, C.c_int, C.c_double, _vp]),
    "rrv_stats_sums_to_m2": (C.c_int, [_vp, C.c_int, _vp]),
    "rrv_pointwise_stats": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Epilogue), C.c_int,
                                      _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rrv_stats_merge": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "rrv_stats_finalize": (C.c_int, [_vp, C.c_int, C.c_int, C.c_float, _vp, _vp]),
    "rrv_filter_fc": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "rrv_warp_nearest_border": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rrv_temporal_loss": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "rrv_warp_backward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with rerevst-code_b200/csrc/build.sh "
                               "(there is no CPU or PyTorch fallback for the CUDA path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.rrv_abi_version() != ABI_VERSION:
            raise RuntimeError("librerevst_b200.so: ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"librerevst_b200 {what}: {lib().rrv_last_error().decode()}")


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream

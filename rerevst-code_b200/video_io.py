"""Motion-JPEG .avi output for frames that are still on the device (``csrc/mjpg_io.cpp``, ``include/rerevst_b200_io.h``).

Replaces the tail of the reference script, ``test/generate_real_video.py:175-186``: there the written frames are read back from
disk and pushed through ``cv2.VideoWriter('MJPG')`` (a CPU JPEG encoder).  ``MjpgWriter`` takes the uint8 BGR frame the RGB head
wrote on the device, encodes it with nvJPEG on the GPU and downloads only the bitstream.  The side library is loaded on first use
and is not needed by anything else in the package; there is no CPU fallback (a missing library raises).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RRV_IO_LIB_PATH") or os.path.join(_HERE, "csrc", "librerevst_b200_io.so")
ABI_VERSION = 1
_vp, _i64 = C.c_void_p, C.c_int64
# name -> (restype, argtypes); must list every symbol of include/rerevst_b200_io.h
SIGNATURES = {
    "rrv_io_abi_version": (C.c_int, []),
    "rrv_io_last_error": (C.c_char_p, []),
    "rrv_mjpg_open": (_vp, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rrv_mjpg_encode": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "rrv_mjpg_flush": (C.c_int, [_vp, C.c_int, _vp]),
    "rrv_mjpg_retrieve": (C.c_int, [_vp, C.c_int, _vp, _vp, _i64, C.POINTER(_i64)]),
    "rrv_mjpg_write_jpeg": (C.c_int, [_vp, _vp, _i64]),
    "rrv_mjpg_frames": (_i64, [_vp]),
    "rrv_mjpg_bytes": (_i64, [_vp]),
    "rrv_mjpg_close": (C.c_int, [_vp]),
}
_handle = None


def lib():
    global _handle
    if _handle is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with rerevst-code_b200/csrc/build.sh (needs libnvjpeg)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.rrv_io_abi_version() != ABI_VERSION:
            raise RuntimeError("librerevst_b200_io.so: ABI version mismatch")
        _handle = h
    return _handle


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().rrv_io_last_error().decode()}")


class MjpgWriter:
    """``cv2.VideoWriter(path, fourcc('MJPG'), fps, (width, height))`` for device frames.

    ``write(frame)`` takes a contiguous uint8 CUDA tensor [height, width, 3] (BGR, what ``Stylization.transfer_device(...,
    out_dtype="u8")`` returns) and appends it; with ``states > 1``, ``encode(frame, k)`` / ``flush(k)`` keep several frames between
    the (asynchronous) GPU encode and the download of their bitstreams.  ``write_jpeg(bytes)`` appends an already encoded frame."""

    def __init__(self, path, fps, size, quality=75, states=1):
        width, height = size
        self._lib = lib()
        self._w = self._lib.rrv_mjpg_open(os.fsencode(path), int(width), int(height), int(fps), int(quality), int(states))
        if not self._w:
            raise RuntimeError("rrv_mjpg_open: " + self._lib.rrv_io_last_error().decode())
        self.size, self.states = (int(width), int(height)), int(states)
        self._keep = [None] * int(states)

    def _frame_ptr(self, frame):
        w, h = self.size
        if not (isinstance(frame, torch.Tensor) and frame.is_cuda and frame.dtype == torch.uint8 and tuple(frame.shape) == (h, w, 3)
                and frame.is_contiguous()):
            raise ValueError(f"expected a contiguous uint8 CUDA tensor of shape {(h, w, 3)} (BGR)")
        return frame.data_ptr()

    def encode(self, frame, state=0, stream=None):
        st = stream if stream is not None else torch.cuda.current_stream(frame.device)
        _check(self._lib.rrv_mjpg_encode(self._w, state, self._frame_ptr(frame), st.cuda_stream), "rrv_mjpg_encode")
        self._keep[state] = (frame, st)             # the frame must outlive the encode

    def flush(self, state=0):
        if self._keep[state] is None:
            raise RuntimeError("flush() without encode()")
        _, st = self._keep[state]
        _check(self._lib.rrv_mjpg_flush(self._w, state, st.cuda_stream), "rrv_mjpg_flush")
        self._keep[state] = None

    def retrieve(self, state=0):
        """The encoded frame of ``state`` as bytes, NOT appended to the file (append later with write_jpeg, in any order)."""
        if self._keep[state] is None:
            raise RuntimeError("retrieve() without encode()")
        _, st = self._keep[state]
        n = _i64(0)
        _check(self._lib.rrv_mjpg_retrieve(self._w, state, st.cuda_stream, None, 0, C.byref(n)), "rrv_mjpg_retrieve")
        buf = (C.c_char * n.value)()
        _check(self._lib.rrv_mjpg_retrieve(self._w, state, st.cuda_stream, C.cast(buf, _vp), n.value, C.byref(n)), "rrv_mjpg_retrieve")
        self._keep[state] = None
        return bytes(buf[:n.value])

    def write(self, frame):
        self.encode(frame, 0)
        self.flush(0)

    def write_jpeg(self, data):
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        _check(self._lib.rrv_mjpg_write_jpeg(self._w, C.cast(buf, _vp), len(data)), "rrv_mjpg_write_jpeg")

    @property
    def frames(self):
        return int(self._lib.rrv_mjpg_frames(self._w))

    @property
    def bytes(self):
        return int(self._lib.rrv_mjpg_bytes(self._w))

    def release(self):
        if self._w:
            w, self._w = self._w, None
            _check(self._lib.rrv_mjpg_close(w), "rrv_mjpg_close")

    close = release

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()

"""ctypes binding for oracle/_ref/libradiofm_ref_uecp.so: the UNMODIFIED reference cRDSGroupDecoder
(RDSGroupDecoder.cpp compiled in place) behind oracle/ref_uecp_harness.cpp.

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Returns the RAW frames handed to
cRadioReceiver::AddUECPDataFrame (before the transport framing) and the names handed to SetChannelName.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libradiofm_ref_uecp.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_uecp_create.restype = C.c_void_p
        for n in ("destroy", "reset"):
            getattr(L, f"ref_uecp_{n}").argtypes = [C.c_void_p]
        L.ref_uecp_set_setting_active.argtypes = [C.c_void_p, C.c_int]
        L.ref_uecp_decode.argtypes = [C.c_void_p, C.POINTER(C.c_uint16), C.c_uint]
        L.ref_uecp_take_frames.restype = C.c_uint
        L.ref_uecp_take_frames.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_uint]
        L.ref_uecp_take_names.restype = C.c_uint
        L.ref_uecp_take_names.argtypes = [C.c_void_p, C.c_char_p, C.c_uint]
        _lib = L
    return _lib


class RefGroupDecoder:
    def __init__(self):
        self._h = lib().ref_uecp_create()

    def close(self):
        if self._h:
            lib().ref_uecp_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        lib().ref_uecp_reset(self._h)

    def set_setting_active(self, active: bool):
        lib().ref_uecp_set_setting_active(self._h, int(active))

    def decode(self, groups) -> list[bytes]:
        """Feed groups [n, 4] u16; returns the raw frames produced, in order."""
        g = np.ascontiguousarray(groups, dtype=np.uint16).reshape(-1, 4)
        lib().ref_uecp_decode(self._h, g.ctypes.data_as(C.POINTER(C.c_uint16)), g.shape[0])
        n = lib().ref_uecp_take_frames(self._h, None, 0)
        buf = np.zeros(max(n, 1), dtype=np.uint8)
        lib().ref_uecp_take_frames(self._h, buf.ctypes.data_as(C.POINTER(C.c_uint8)), n)
        out, i = [], 0
        raw = buf[:n].tobytes()
        while i < n:
            ln = raw[i] | (raw[i + 1] << 8)
            out.append(raw[i + 2:i + 2 + ln])
            i += 2 + ln
        return out

    def take_names(self) -> list[bytes]:
        n = lib().ref_uecp_take_names(self._h, None, 0)
        buf = C.create_string_buffer(max(n, 1))
        lib().ref_uecp_take_names(self._h, buf, n)
        raw = buf.raw[:n]
        return [raw[i:i + 8] for i in range(0, n, 9)]

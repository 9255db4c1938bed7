"""ctypes binding for oracle/_ref/libradiofm_oracle.so -- the plain-C restatement.

TEST INFRASTRUCTURE ONLY -- see oracle/README.md and oracle/radiofm_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libradiofm_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)

TAP_NAMES = ("tuned", "demod_in", "baseband", "rds_dec", "rds_lp", "rds_pll", "rds_mf", "rds_sync",
             "mono_rs", "pilot38", "rawstereo", "stereo_rs", "lp", "deemph", "notch")
_TAP_COMPLEX = {"tuned", "demod_in", "rds_dec", "rds_lp"}


def build() -> None:
    """Compile the restatement (gcc only; needs nothing from /root/reference)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        L = C.CDLL(LIB_PATH)
        L.rfo_create.restype = C.c_void_p
        L.rfo_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_int]
        L.rfo_destroy.argtypes = [C.c_void_p]
        L.rfo_reset.argtypes = [C.c_void_p]
        L.rfo_u8_to_cf32.argtypes = [_u8p, C.c_uint, _f32p]
        L.rfo_process_cf32.restype = C.c_uint
        L.rfo_process_cf32.argtypes = [C.c_void_p, _f32p, C.c_uint, _f32p]
        L.rfo_process_u8.restype = C.c_uint
        L.rfo_process_u8.argtypes = [C.c_void_p, _u8p, C.c_uint, _f32p]
        L.rfo_take_groups.restype = C.c_uint
        L.rfo_take_groups.argtypes = [C.c_void_p, _u16p, C.c_uint]
        L.rfo_take_bits.restype = C.c_uint
        L.rfo_take_bits.argtypes = [C.c_void_p, _u8p, C.c_uint]
        L.rfo_status.argtypes = [C.c_void_p, _f32p]
        L.rfo_constants.argtypes = [C.c_void_p, _f64p]
        L.rfo_table.restype = C.c_uint
        L.rfo_table.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_uint]
        L.rfo_tap.restype = _f32p
        L.rfo_tap.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint)]
        L.rfo_last_stereo.restype = C.c_uint
        L.rfo_last_stereo.argtypes = [C.c_void_p]
        L.rfo_atan2f.restype = C.c_float
        L.rfo_atan2f.argtypes = [C.c_float, C.c_float]
        L.rfo_sincos.argtypes = [C.c_float, _f32p, _f32p]
        for name in ("freqshift", "downsample", "rdsdc", "fir", "iir", "pilot", "rdssync"):
            getattr(L, f"rfo_{name}_destroy").argtypes = [C.c_void_p]
        L.rfo_freqshift_create.restype = C.c_void_p
        L.rfo_freqshift_create.argtypes = [C.c_float, C.c_float]
        L.rfo_freqshift_reset.argtypes = [C.c_void_p]
        L.rfo_freqshift_process.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.rfo_downsample_create.restype = C.c_void_p
        L.rfo_downsample_create.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_int]
        L.rfo_downsample_reset.argtypes = [C.c_void_p]
        L.rfo_downsample_process_real.restype = C.c_uint
        L.rfo_downsample_process_real.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.rfo_downsample_process_complex.restype = C.c_uint
        L.rfo_downsample_process_complex.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.rfo_downsample_coeff.restype = C.c_uint
        L.rfo_downsample_coeff.argtypes = [C.c_void_p, _f32p]
        L.rfo_rdsdc_create.restype = C.c_void_p
        L.rfo_rdsdc_set_frequency.argtypes = [C.c_void_p, C.c_float]
        L.rfo_rdsdc_set_data_rate.restype = C.c_float
        L.rfo_rdsdc_set_data_rate.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.rfo_rdsdc_set_wfm_data_rate.restype = C.c_float
        L.rfo_rdsdc_set_wfm_data_rate.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.rfo_rdsdc_process.restype = C.c_int
        L.rfo_rdsdc_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
        L.rfo_rdsdc_stages.restype = C.c_int
        L.rfo_rdsdc_stages.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        L.rfo_fir_create.restype = C.c_void_p
        L.rfo_fir_init_lp.restype = C.c_int
        L.rfo_fir_init_lp.argtypes = [C.c_void_p, C.c_uint] + [C.c_float] * 5
        L.rfo_fir_init_hp.restype = C.c_int
        L.rfo_fir_init_hp.argtypes = [C.c_void_p, C.c_uint] + [C.c_float] * 5
        L.rfo_fir_init_const.argtypes = [C.c_void_p, C.c_uint, _f32p, C.c_float]
        L.rfo_fir_coef.restype = C.c_uint
        L.rfo_fir_coef.argtypes = [C.c_void_p, _f32p]
        L.rfo_fir_process_real.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.rfo_fir_process_complex.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.rfo_fir_process_two.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.rfo_iir_create.restype = C.c_void_p
        L.rfo_iir_init.restype = C.c_int
        L.rfo_iir_init.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.rfo_iir_coef.argtypes = [C.c_void_p, _f32p]
        L.rfo_iir_process_real.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.rfo_iir_process_complex.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.rfo_iir_process_two.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.rfo_pilot_create.restype = C.c_void_p
        L.rfo_pilot_create.argtypes = [C.c_float, C.c_float, C.c_float]
        L.rfo_pilot_process.restype = C.c_int
        L.rfo_pilot_process.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.rfo_pilot_level.restype = C.c_float
        L.rfo_pilot_level.argtypes = [C.c_void_p]
        L.rfo_rdssync_create.restype = C.c_void_p
        L.rfo_rdssync_reset.argtypes = [C.c_void_p]
        L.rfo_rdssync_push_bits.argtypes = [C.c_void_p, _u8p, C.c_uint]
        L.rfo_rdssync_take_groups.restype = C.c_uint
        L.rfo_rdssync_take_groups.argtypes = [C.c_void_p, _u16p, C.c_uint]
        L.rfo_rds_check_block.restype = C.c_uint32
        L.rfo_rds_check_block.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def u8_to_cf32(iq_u8: np.ndarray) -> np.ndarray:
    iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1, 2)
    out = np.empty((iq_u8.shape[0], 2), dtype=np.float32)
    lib().rfo_u8_to_cf32(_p(iq_u8, _u8p), iq_u8.shape[0], _p(out, _f32p))
    return out


class OracleFmDecoder:
    """Plain-C restatement of cFmDecoder (FmDecode.h:110-165)."""

    def __init__(self, fs_if, tuning_offset, fs_pcm=48000.0, bw_pcm=15000.0, downsample=1, usver=False):
        self._h = lib().rfo_create(fs_if, tuning_offset, fs_pcm, bw_pcm, downsample, int(usver))

    def close(self):
        if self._h:
            lib().rfo_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        lib().rfo_reset(self._h)

    def process_u8(self, iq_u8: np.ndarray) -> np.ndarray:
        iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1, 2)
        n = iq_u8.shape[0]
        audio = np.empty(max(2 * n, 2), dtype=np.float32)
        k = lib().rfo_process_u8(self._h, _p(iq_u8, _u8p), n, _p(audio, _f32p))
        return audio[:k].copy()

    def process_cf32(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1, 2)
        n = iq.shape[0]
        audio = np.empty(max(2 * n, 2), dtype=np.float32)
        k = lib().rfo_process_cf32(self._h, _p(iq, _f32p), n, _p(audio, _f32p))
        return audio[:k].copy()

    def tap(self, name: str) -> np.ndarray:
        n = C.c_uint(0)
        p = lib().rfo_tap(self._h, name.encode(), C.byref(n))
        if not p or n.value == 0:
            return np.zeros((0, 2) if name in _TAP_COMPLEX else 0, dtype=np.float32)
        a = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
        return a.reshape(-1, 2) if name in _TAP_COMPLEX else a

    def taps(self) -> dict:
        out = {name: self.tap(name) for name in TAP_NAMES}
        out["stereo"] = bool(lib().rfo_last_stereo(self._h))
        return out

    def take_groups(self, max_groups=4096) -> np.ndarray:
        out = np.zeros((max_groups, 4), dtype=np.uint16)
        k = lib().rfo_take_groups(self._h, _p(out, _u16p), max_groups)
        return out[:k].copy()

    def take_bits(self, max_bits=1 << 20) -> np.ndarray:
        out = np.zeros(max_bits, dtype=np.uint8)
        k = lib().rfo_take_bits(self._h, _p(out, _u8p), max_bits)
        return out[:k].copy()

    def status(self) -> dict:
        s = np.zeros(6, dtype=np.float32)
        lib().rfo_status(self._h, _p(s, _f32p))
        return {"stereo": bool(s[0]), "if_level": s[1], "bb_level": s[2], "bb_mean": s[3],
                "pilot_level": s[4], "tuning_offset": s[5]}

    def constants(self) -> np.ndarray:
        s = np.zeros(64, dtype=np.float64)
        lib().rfo_constants(self._h, _p(s, _f64p))
        return s

    def table(self, which: int) -> np.ndarray:
        out = np.zeros(4096, dtype=np.float32)
        k = lib().rfo_table(self._h, which, _p(out, _f32p), out.size)
        return out[:k].copy()

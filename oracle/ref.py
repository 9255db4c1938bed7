"""ctypes binding for oracle/_ref/libradiofm_ref.so (the UNMODIFIED reference chain).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  Imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the product.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libradiofm_ref.so")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


class RefTaps(C.Structure):
    _fields_ = [(n, _f32p) for n in (
        "tuned", "demod_in", "baseband", "rds_dec", "rds_lp", "rds_pll", "rds_mf", "rds_sync",
        "mono_rs", "pilot38", "rawstereo", "stereo_rs", "lp", "deemph", "notch")] + [
        ("nb", C.c_uint32), ("nr", C.c_uint32), ("na", C.c_uint32), ("stereo", C.c_uint32)]


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_fm_create.restype = C.c_void_p
        L.ref_fm_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint, C.c_int]
        L.ref_fm_destroy.argtypes = [C.c_void_p]
        L.ref_fm_reset.argtypes = [C.c_void_p]
        L.ref_fm_process_cf32.restype = C.c_uint
        L.ref_fm_process_cf32.argtypes = [C.c_void_p, _f32p, C.c_uint, _f32p]
        L.ref_fm_process_u8.restype = C.c_uint
        L.ref_fm_process_u8.argtypes = [C.c_void_p, _u8p, C.c_uint, _f32p]
        L.ref_fm_process_staged.restype = C.c_uint
        L.ref_fm_process_staged.argtypes = [C.c_void_p, _f32p, C.c_uint, _f32p, C.POINTER(RefTaps)]
        L.ref_fm_take_groups.restype = C.c_uint
        L.ref_fm_take_groups.argtypes = [C.c_void_p, _u16p, C.c_uint]
        L.ref_fm_take_bits.restype = C.c_uint
        L.ref_fm_take_bits.argtypes = [C.c_void_p, _u8p, C.c_uint]
        L.ref_fm_status.argtypes = [C.c_void_p, _f32p]
        L.ref_fm_constants.argtypes = [C.c_void_p, _f64p]
        L.ref_fm_table.restype = C.c_uint
        L.ref_fm_table.argtypes = [C.c_void_p, C.c_int, _f32p, C.c_uint]
        L.ref_u8_to_cf32.argtypes = [_u8p, C.c_uint, _f32p]
        for name in ("finetuner", "pilot", "freqshift", "downsample", "rdsdc", "fir", "iir", "rds"):
            getattr(L, f"ref_{name}_destroy").argtypes = [C.c_void_p]
        L.ref_finetuner_create.restype = C.c_void_p
        L.ref_finetuner_create.argtypes = [C.c_uint, C.c_int]
        L.ref_finetuner_process.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_pilot_create.restype = C.c_void_p
        L.ref_pilot_create.argtypes = [C.c_float, C.c_float, C.c_float]
        L.ref_pilot_process.restype = C.c_int
        L.ref_pilot_process.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_pilot_level.restype = C.c_float
        L.ref_pilot_level.argtypes = [C.c_void_p]
        L.ref_freqshift_create.restype = C.c_void_p
        L.ref_freqshift_create.argtypes = [C.c_float, C.c_float]
        L.ref_freqshift_reset.argtypes = [C.c_void_p]
        L.ref_freqshift_process.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_downsample_create.restype = C.c_void_p
        L.ref_downsample_create.argtypes = [C.c_uint, C.c_double, C.c_double, C.c_int]
        L.ref_downsample_reset.argtypes = [C.c_void_p]
        L.ref_downsample_process_real.restype = C.c_uint
        L.ref_downsample_process_real.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_downsample_process_complex.restype = C.c_uint
        L.ref_downsample_process_complex.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_downsample_coeff.restype = C.c_uint
        L.ref_downsample_coeff.argtypes = [C.c_void_p, _f32p]
        L.ref_rdsdc_create.restype = C.c_void_p
        L.ref_rdsdc_set_frequency.argtypes = [C.c_void_p, C.c_float]
        L.ref_rdsdc_set_data_rate.restype = C.c_float
        L.ref_rdsdc_set_data_rate.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_rdsdc_set_wfm_data_rate.restype = C.c_float
        L.ref_rdsdc_set_wfm_data_rate.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.ref_rdsdc_process.restype = C.c_int
        L.ref_rdsdc_process.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
        L.ref_rdsdc_stages.restype = C.c_int
        L.ref_rdsdc_stages.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        L.ref_fir_create.restype = C.c_void_p
        for nm in ("ref_fir_init_lp", "ref_fir_init_hp"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_uint] + [C.c_float] * 5
        L.ref_fir_init_const.argtypes = [C.c_void_p, C.c_uint, _f32p, C.c_float]
        L.ref_fir_init_const_iq.argtypes = [C.c_void_p, C.c_uint, _f32p, _f32p, C.c_float]
        L.ref_fir_coef.restype = C.c_uint
        L.ref_fir_coef.argtypes = [C.c_void_p, _f32p]
        L.ref_fir_process_real.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_fir_process_complex.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_fir_process_two.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_iir_create.restype = C.c_void_p
        L.ref_iir_init.restype = C.c_int
        L.ref_iir_init.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_iir_coef.argtypes = [C.c_void_p, _f32p]
        L.ref_iir_process_real.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_iir_process_complex.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_iir_process_two.argtypes = [C.c_void_p, _f32p, _f32p, C.c_uint]
        L.ref_rds_create.restype = C.c_void_p
        L.ref_rds_create.argtypes = [C.c_float]
        L.ref_rds_reset.argtypes = [C.c_void_p]
        L.ref_rds_process.argtypes = [C.c_void_p, _f32p, C.c_uint]
        L.ref_rds_take_groups.restype = C.c_uint
        L.ref_rds_take_groups.argtypes = [C.c_void_p, _u16p, C.c_uint]
        L.ref_rds_push_bits.argtypes = [C.c_void_p, _u8p, C.c_uint]
        L.ref_rds_check_block.restype = C.c_uint32
        L.ref_rds_check_block.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def u8_to_cf32(iq_u8: np.ndarray) -> np.ndarray:
    iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1, 2)
    out = np.empty((iq_u8.shape[0], 2), dtype=np.float32)
    lib().ref_u8_to_cf32(_p(iq_u8, _u8p), iq_u8.shape[0], _p(out, _f32p))
    return out


TAP_NAMES = ("tuned", "demod_in", "baseband", "rds_dec", "rds_lp", "rds_pll", "rds_mf", "rds_sync",
             "mono_rs", "pilot38", "rawstereo", "stereo_rs", "lp", "deemph", "notch")
_TAP_COMPLEX = {"tuned", "demod_in", "rds_dec", "rds_lp"}
_TAP_RATE = {"tuned": "n", "demod_in": "nb", "baseband": "nb", "rds_dec": "nr", "rds_lp": "nr",
             "rds_pll": "nr", "rds_mf": "nr", "rds_sync": "nr", "mono_rs": "na", "pilot38": "nb",
             "rawstereo": "nb", "stereo_rs": "na", "lp": "na2", "deemph": "na2", "notch": "na2"}


class RefFmDecoder:
    """The reference cFmDecoder (FmDecode.h:110-165) behind the harness."""

    def __init__(self, fs_if, tuning_offset, fs_pcm=48000.0, bw_pcm=15000.0, downsample=1, usver=False):
        self._h = lib().ref_fm_create(fs_if, tuning_offset, fs_pcm, bw_pcm, downsample, int(usver))

    def close(self):
        if self._h:
            lib().ref_fm_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        lib().ref_fm_reset(self._h)

    def process_u8(self, iq_u8: np.ndarray) -> np.ndarray:
        iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8).reshape(-1, 2)
        n = iq_u8.shape[0]
        audio = np.empty(2 * n, dtype=np.float32)
        k = lib().ref_fm_process_u8(self._h, _p(iq_u8, _u8p), n, _p(audio, _f32p))
        return audio[:k].copy()

    def process_cf32(self, iq: np.ndarray) -> np.ndarray:
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1, 2)
        n = iq.shape[0]
        audio = np.empty(2 * n, dtype=np.float32)
        k = lib().ref_fm_process_cf32(self._h, _p(iq, _f32p), n, _p(audio, _f32p))
        return audio[:k].copy()

    def process_staged(self, iq: np.ndarray, want=TAP_NAMES):
        """Returns (audio, taps dict incl. 'stereo').  iq is cf32 [n,2]."""
        iq = np.ascontiguousarray(iq, dtype=np.float32).reshape(-1, 2)
        n = iq.shape[0]
        audio = np.empty(2 * n, dtype=np.float32)
        taps = RefTaps()
        bufs = {}
        for name in want:
            bufs[name] = np.zeros(2 * n, dtype=np.float32)
            setattr(taps, name, _p(bufs[name], _f32p))
        k = lib().ref_fm_process_staged(self._h, _p(iq, _f32p), n, _p(audio, _f32p), C.byref(taps))
        sizes = {"n": n, "nb": taps.nb, "nr": taps.nr, "na": taps.na, "na2": 2 * taps.na}
        out = {"nb": taps.nb, "nr": taps.nr, "na": taps.na, "stereo": bool(taps.stereo)}
        for name in want:
            cnt = sizes[_TAP_RATE[name]]
            if name in _TAP_COMPLEX:
                out[name] = bufs[name][:2 * cnt].reshape(-1, 2).copy()
            else:
                out[name] = bufs[name][:cnt].copy()
        return audio[:k].copy(), out

    def take_groups(self, max_groups=4096) -> np.ndarray:
        out = np.zeros((max_groups, 4), dtype=np.uint16)
        k = lib().ref_fm_take_groups(self._h, _p(out, _u16p), max_groups)
        return out[:k].copy()

    def take_bits(self, max_bits=1 << 20) -> np.ndarray:
        out = np.zeros(max_bits, dtype=np.uint8)
        k = lib().ref_fm_take_bits(self._h, _p(out, _u8p), max_bits)
        return out[:k].copy()

    def status(self) -> dict:
        s = np.zeros(6, dtype=np.float32)
        lib().ref_fm_status(self._h, _p(s, _f32p))
        return {"stereo": bool(s[0]), "if_level": s[1], "bb_level": s[2], "bb_mean": s[3],
                "pilot_level": s[4], "tuning_offset": s[5]}

    def constants(self) -> np.ndarray:
        s = np.zeros(64, dtype=np.float64)
        lib().ref_fm_constants(self._h, _p(s, _f64p))
        return s

    def table(self, which: int) -> np.ndarray:
        out = np.zeros(4096, dtype=np.float32)
        k = lib().ref_fm_table(self._h, which, _p(out, _f32p), out.size)
        return out[:k].copy()

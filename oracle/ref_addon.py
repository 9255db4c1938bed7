"""ctypes binding for oracle/_ref/libradiofm_ref_addon.so: the UNMODIFIED reference add-on receive path (cRadioReceiver
+ cRtlSdrSource + cFmDecoder + RDS chain, every source file compiled in place) behind oracle/ref_addon_harness.cpp,
with stand-ins for the Kodi dev-kit, TinyXML and librtlsdr (oracle/stub/).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  One instance at a time (the stand-in device is process-wide).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libradiofm_ref_addon.so")
_lib = None
STREAMCHANGE = -11


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.refaddon_create.restype = C.c_void_p
        L.refaddon_create.argtypes = [C.c_float]
        L.refaddon_reload_channels.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.refaddon_open.argtypes = [C.c_void_p]
        L.refaddon_params.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint),
                                      C.POINTER(C.c_uint), C.POINTER(C.c_double)]
        L.refaddon_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.refaddon_read.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.c_void_p, C.c_int]
        L.refaddon_signal.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]
        L.refaddon_audio_level.restype = C.c_float
        L.refaddon_audio_level.argtypes = [C.c_void_p]
        L.refaddon_queued_samples.restype = C.c_ulonglong
        L.refaddon_queued_samples.argtypes = [C.c_void_p]
        L.refaddon_set_stream_change.argtypes = [C.c_void_p]
        L.refaddon_stream_properties.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int]
        L.refaddon_device_log.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.refaddon_close.argtypes = [C.c_void_p]
        L.refaddon_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class RefAddon:
    """cRadioReceiver with one channel at `channel_freq_hz`; open() = OpenLiveStream."""

    def __init__(self, channel_freq_hz=100.0e6):
        self._h = lib().refaddon_create(channel_freq_hz)
        self._buf = np.empty(1 << 20, np.uint8)

    def destroy(self):
        if self._h:
            lib().refaddon_destroy(self._h)
            self._h = None

    __del__ = destroy

    def reload_channels(self):
        f = C.c_float(0)
        n = lib().refaddon_reload_channels(self._h, C.byref(f))
        return n, f.value

    def open(self) -> bool:
        return bool(lib().refaddon_open(self._h))

    def close(self):
        lib().refaddon_close(self._h)

    def params(self):
        fs, off, tf = C.c_double(), C.c_double(), C.c_double()
        ds, blk = C.c_uint(), C.c_uint()
        lib().refaddon_params(self._h, C.byref(fs), C.byref(off), C.byref(ds), C.byref(blk), C.byref(tf))
        return dict(if_rate=fs.value, tuning_offset=off.value, downsample=ds.value, block_length=blk.value, tuner_freq=tf.value)

    def feed(self, iq_u8, short_read_after_next=False):
        a = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        lib().refaddon_feed(self._h, a.ctypes.data, a.size, int(short_read_after_next))

    def read(self):
        """One DemuxRead: (stream id, pts, duration, payload) -- payload float32 audio (id 1), bytes (id 2), None;
        None when the call would wait for the source."""
        sid, size = C.c_int(), C.c_int()
        pts, dur = C.c_double(), C.c_double()
        rc = lib().refaddon_read(self._h, C.byref(sid), C.byref(size), C.byref(pts), C.byref(dur), self._buf.ctypes.data,
                                 self._buf.size)
        if rc == 0:
            return None
        if rc != 1:
            raise RuntimeError(f"refaddon_read -> {rc}")
        raw = self._buf[:size.value]
        if sid.value == 1:
            return 1, pts.value, dur.value, raw.view(np.float32).copy()
        if sid.value == 2:
            return 2, pts.value, dur.value, raw.tobytes()
        return sid.value, pts.value, dur.value, None

    def read_all(self):
        out = []
        while True:
            p = self.read()
            if p is None:
                return out
            out.append(p)

    def signal(self):
        il, al = C.c_float(), C.c_float()
        st, sig, snr = C.c_int(), C.c_int(), C.c_int()
        text = C.create_string_buffer(256)
        if not lib().refaddon_signal(self._h, C.byref(il), C.byref(al), C.byref(st), C.byref(sig), C.byref(snr), text, 256):
            return None
        return dict(if_level_db=il.value, audio_level_db=al.value, stereo=bool(st.value), signal=sig.value, snr=snr.value,
                    status=text.value.decode("utf-8", "replace"))

    def audio_level(self):
        return np.float32(lib().refaddon_audio_level(self._h))

    def queued_samples(self):
        return int(lib().refaddon_queued_samples(self._h))

    def set_stream_change(self):
        lib().refaddon_set_stream_change(self._h)

    def stream_properties(self):
        out = (C.c_int * 24)()
        n = lib().refaddon_stream_properties(self._h, out, 4)
        keys = ("pid", "codec_type", "channels", "sample_rate", "bits", "bit_rate")
        return [dict(zip(keys, out[6 * i:6 * i + 6])) for i in range(n)]

    def device_log(self):
        text = C.create_string_buffer(1 << 14)
        lib().refaddon_device_log(self._h, text, 1 << 14)
        return text.value.decode().splitlines()

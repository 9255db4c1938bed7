"""TEST INFRASTRUCTURE -- CPU composition of the reference's classes for one station of a wideband capture
(SURVEY.md section 8d C5 / 8f N3), built on the plain-C restatement (oracle/port.py):

    u8 -> float (RTL_SDR_Source.cpp:207-211)
    mixer "osc":       CRDSDownConvert::SetFrequency(-f_k)                       (DownConvert.cpp:311-320,438-465)
    mixer "freqshift": cFreqShift(-f_k, Fs), Reset() before every front-end block (FreqShift.cpp:10-76), then
                       CRDSDownConvert at 0 Hz
    CRDSDownConvert::SetWfmDataRate(Fs, bw) + ProcessData per front-end block     (DownConvert.cpp:378-399,412-489)
    cFmDecoder(out_rate, 0, 48000, 15000, 1)::ProcessStream per `blocks_per_call` blocks (FmDecode.cpp:417-502)

Only tests/ and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import port

_f32p = C.POINTER(C.c_float)


def _P(a):
    return a.ctypes.data_as(_f32p)


class OracleStation:
    def __init__(self, f_station: float, fs: float = 50.0e6, front_block: int = 32000, blocks_per_call: int = 64,
                 mixer: str = "osc", max_bw: float = 100000.0):
        self.L = L = port.lib()
        self.mixer, self.front_block, self.blocks_per_call = mixer, front_block, blocks_per_call
        self.dc = L.rfo_rdsdc_create()
        L.rfo_rdsdc_set_frequency(self.dc, np.float32(-f_station if mixer == "osc" else 0.0))
        self.rate = L.rfo_rdsdc_set_wfm_data_rate(self.dc, np.float32(fs), np.float32(max_bw))
        self.shift = L.rfo_freqshift_create(np.float32(-f_station), np.float32(fs)) if mixer == "freqshift" else None
        self.dec = port.OracleFmDecoder(float(self.rate), 0.0, downsample=1)

    def close(self):
        if self.dc:
            self.L.rfo_rdsdc_destroy(self.dc)
            self.dc = None
        if self.shift:
            self.L.rfo_freqshift_destroy(self.shift)
            self.shift = None

    __del__ = close

    def baseband(self, capture_u8: np.ndarray) -> np.ndarray:
        """capture [blocks * front_block, 2] u8 -> decimated complex baseband [m, 2]."""
        x = port.u8_to_cf32(capture_u8)
        out = []
        for b in range(x.shape[0] // self.front_block):
            z = np.array(x[b * self.front_block:(b + 1) * self.front_block], dtype=np.float32, order="C")
            if self.shift:
                self.L.rfo_freqshift_reset(self.shift)
                self.L.rfo_freqshift_process(self.shift, _P(z), z.shape[0])
            y = np.zeros_like(z)
            k = self.L.rfo_rdsdc_process(self.dc, z.shape[0], _P(z), _P(y))
            out.append(y[:k].copy())
        return np.concatenate(out)

    def process_u8(self, capture_u8: np.ndarray) -> np.ndarray:
        """One demodulator call: blocks_per_call front-end blocks -> interleaved L,R audio."""
        return self.dec.process_cf32(self.baseband(capture_u8))

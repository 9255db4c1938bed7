"""CRDSDownConvert on the GPU (rfm_downconvert_*, include/radiofm_b200.h) against the oracle's restatement of
DownConvert.cpp:271-727, bit for bit: planner, NCO_OSC mixer, every stage kind, carried state over several calls, the
time-chunked path of long calls, the fused u8 conversion of a shared wideband capture."""
import ctypes as C

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

_f32p = C.POINTER(C.c_float)


def P(a):
    return a.ctypes.data_as(_f32p)


class OracleDc:
    def __init__(self, port, freq, in_rate, max_bw, wfm):
        self.L = port.lib()
        self.h = self.L.rfo_rdsdc_create()
        self.L.rfo_rdsdc_set_frequency(self.h, np.float32(freq))
        fn = self.L.rfo_rdsdc_set_wfm_data_rate if wfm else self.L.rfo_rdsdc_set_data_rate
        self.rate = fn(self.h, np.float32(in_rate), np.float32(max_bw))
        lens = (C.c_int * 16)()
        ns = self.L.rfo_rdsdc_stages(self.h, lens, 16)
        self.stages = list(lens[:ns])

    def set_frequency(self, f):
        self.L.rfo_rdsdc_set_frequency(self.h, np.float32(f))

    def process(self, x):
        z = np.array(x, dtype=np.float32, order="C").reshape(-1, 2)
        y = np.zeros((z.shape[0], 2), dtype=np.float32)
        k = self.L.rfo_rdsdc_process(self.h, z.shape[0], P(z), P(y))
        return y[:k].copy()

    def __del__(self):
        self.L.rfo_rdsdc_destroy(self.h)


def _rand_iq(rng, rows, n):
    return (rng.standard_normal((rows, n, 2)) * 0.3).astype(np.float32)


@pytest.mark.parametrize("in_rate,max_bw,freqs,lengths", [
    # fixed 11-tap, HB15, HB23, HB47; every row the same frequency: one shared oscillator table
    (250000.0, 4800.0, [-57000.0, -57000.0, -57000.0], (8192, 16384, 4096, 16384)),
    # fixed 11-tap, HB15, HB23; one table per row
    (218181.8125, 4800.0, [-57000.0, 12345.0, 0.0, 801.5], (8192, 16384, 1024, 16384)),
    # HB15, HB23, HB51 (the chain cRDSRxSignalProcessor never plans, a wide output)
    (2.4e6, 100000.0, [-57000.0, 3000.0], (8192, 16384, 400, 16384)),
    # CIC3 first, 8 stages
    (4.0e6, 4800.0, [100000.0, -250000.0], (16384, 32768, 12288, 16384)),
])
def test_set_data_rate_chain(rfm, port, in_rate, max_bw, freqs, lengths):
    rng = np.random.default_rng(5)
    d = rfm.DownConvertBatch(freqs, in_rate, max_bw, wfm=False, max_len=32768)
    oracles = [OracleDc(port, f, in_rate, max_bw, False) for f in freqs]
    assert d.stage_taps == oracles[0].stages and np.float32(d.output_rate) == np.float32(oracles[0].rate)
    for n in lengths:
        x = _rand_iq(rng, len(freqs), n)
        y = d.process_cf32(x)
        assert y.shape[1] == n >> d.n_stages
        for r, o in enumerate(oracles):
            assert bits_equal(y[r], o.process(x[r])), (in_rate, r, n)
    # SetFrequency keeps the carried phasor (DownConvert.cpp:311-320)
    f2 = [f + 1000.0 for f in freqs]
    d.set_frequency(f2)
    x = _rand_iq(rng, len(freqs), 16384)
    y = d.process_cf32(x)
    for r, o in enumerate(oracles):
        o.set_frequency(f2[r])
        assert bits_equal(y[r], o.process(x[r])), ("after SetFrequency", r)
    d.close()


def test_wfm_chain_shared_u8_capture(rfm, port):
    """SetWfmDataRate(50 MS/s): 7 x HB51 -> 390 625 S/s; stations of one shared u8 capture, 32000-sample blocks."""
    rng = np.random.default_rng(6)
    fs, blk = 50.0e6, 32000
    freqs = [-7.0e6, 200000.0, 0.0, 3.2e6, -400000.0]
    d = rfm.DownConvertBatch(freqs, fs, 100000.0, wfm=True, max_len=blk)
    assert d.stage_taps == [51] * 7 and d.output_rate == 390625.0
    oracles = [OracleDc(port, f, fs, 100000.0, True) for f in freqs]
    for b in range(3):
        cap = rng.integers(0, 256, size=(blk, 2), dtype=np.uint8)
        y = d.process_u8(cap, shared_capture=True)
        x = port.u8_to_cf32(cap)
        assert y.shape == (len(freqs), 250, 2)
        for r, o in enumerate(oracles):
            assert bits_equal(y[r], o.process(x)), (b, r)
    d.close()


def test_zero_hz_oscillator_cycle(rfm, port):
    """At 0 Hz the NCO_OSC phasor falls into a short cycle after ~120 samples; the table is then completed in parallel.
    Several calls, lengths with every remainder mod 4 of the samples left after detection."""
    rng = np.random.default_rng(7)
    fs = 50.0e6
    d = rfm.DownConvertBatch([0.0, 0.0, 0.0], fs, 100000.0, wfm=True, max_len=32000)
    o = OracleDc(port, 0.0, fs, 100000.0, True)
    for n in (32000, 12800, 6400 + 128, 32000 - 128):
        x = _rand_iq(rng, 1, n)
        y = d.process_cf32(np.repeat(x, 3, axis=0))
        ref = o.process(x[0])
        for r in range(3):
            assert bits_equal(y[r], ref), (n, r)
    d.close()


def test_long_call_time_chunks(rfm, port):
    """A call long enough to be cut into time chunks (warm-up re-derives the stage histories) equals the sequential
    reference, for the uniform HB51 chain and for a mixed chain; state carried into a following short call."""
    rng = np.random.default_rng(8)
    for in_rate, bw, wfm, freqs, n in ((50.0e6, 100000.0, True, [1.0e6, -2.2e6], 64 * 6400),
                                       (2.0e6, 4800.0, False, [100000.0], 1 << 18)):
        d = rfm.DownConvertBatch(freqs, in_rate, bw, wfm=wfm, max_len=n)
        oracles = [OracleDc(port, f, in_rate, bw, wfm) for f in freqs]
        for m in (n, n // 8):
            x = _rand_iq(rng, len(freqs), m)
            y = d.process_cf32(x)
            for r, o in enumerate(oracles):
                assert bits_equal(y[r], o.process(x[r])), (in_rate, m, r)
        d.close()


def test_rejects_what_the_reference_misfilters(rfm):
    d = rfm.DownConvertBatch([0.0], 50.0e6, 100000.0, wfm=True, max_len=32000)
    x = np.zeros((1, 12864, 2), dtype=np.float32)
    with pytest.raises(rfm.RadioFmError):
        d.process_cf32(x[:, :12800 + 64])       # not a multiple of 2^7
    with pytest.raises(rfm.RadioFmError):
        d.process_cf32(x[:, :6400 - 128])       # last stage would see fewer than 2 * (51 - 1) samples
    assert d.process_cf32(x[:, :12800]).shape == (1, 100, 2)
    d.close()

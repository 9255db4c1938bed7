"""Stand-alone filter primitives (rfm_iir / rfm_fir / rfm_rdsproc, include/radiofm_b200.h) against the oracle's
restatements of cIirFilter, cFirFilter and cRDSRxSignalProcessor, bit for bit, over several calls with carried state."""
import ctypes as C

import numpy as np
import pytest

from conftest import RATES, bits_equal, station

pytestmark = pytest.mark.gpu

_f32p = C.POINTER(C.c_float)


def P(a):
    return a.ctypes.data_as(_f32p)


@pytest.mark.parametrize("ftype,f0,q,fs", [(3, 19000.0, 5.0, 48000.0), (2, 1187.5, 500.0, 31250.0),
                                          (1, 30.0, 2.0, 48000.0), (0, 15000.0, 0.7071, 48000.0)])
def test_iir_all_types_and_entry_points(rfm, port, ftype, f0, q, fs):
    rng = np.random.default_rng(21)
    L = port.lib()
    rows = 35                       # more than one warp of rows, ragged
    f = rfm.IirFilterBatch(rows, max_len=4096)
    assert f.init(ftype, f0, q, fs)
    hs = [L.rfo_iir_create() for _ in range(rows)]
    for h in hs:
        assert L.rfo_iir_init(h, ftype, f0, q, fs)
    co = np.zeros(5, dtype=np.float32)
    L.rfo_iir_coef(hs[0], P(co))
    assert bits_equal(f.coefficients(), co)
    for n in (1000, 37, 4096):
        x = rng.standard_normal((rows, n)).astype(np.float32)
        y = f.process_real(x)
        for r, h in enumerate(hs):
            ref = x[r].copy()
            L.rfo_iir_process_real(h, P(ref), n)
            assert bits_equal(y[r], ref), ("real", r, n)
    # ProcessTwo and complex share the delay pairs a / b: continue on the same objects
    for n in (512, 33):
        a, b = rng.standard_normal((2, rows, n)).astype(np.float32)
        ya, yb = f.process_two(a, b)
        z = rng.standard_normal((rows, n, 2)).astype(np.float32)
        yz = f.process_complex(z)
        for r, h in enumerate(hs):
            ra, rb, rz = a[r].copy(), b[r].copy(), z[r].copy()
            L.rfo_iir_process_two(h, P(ra), P(rb), n)
            L.rfo_iir_process_complex(h, P(rz), n)
            assert bits_equal(ya[r], ra) and bits_equal(yb[r], rb) and bits_equal(yz[r], rz), (r, n)
    assert not f.init(7, f0, q, fs)          # unknown type: Init returns false
    for h in hs:
        L.rfo_iir_destroy(h)
    f.close()


def test_fir_lowpass_and_const_taps(rfm, port):
    rng = np.random.default_rng(22)
    L = port.lib()
    rows = 5
    f = rfm.FirFilterBatch(rows, max_len=8192)
    hs = [L.rfo_fir_create() for _ in range(rows)]
    # the audio low-pass of the chain (FmDecode.cpp:286) and the RDS low-pass (RDSProcess.cpp:99)
    # ... and a high-pass design (InitHPFilter, FirFilter.cpp:195-264: no caller in the reference)
    for kind, spec in (("lp", (0, 1.0, 60.0, 15000.0, 21000.0, 48000.0)), ("lp", (0, 1.0, 40.0, 2400.0, 3120.0, 31250.0)),
                       ("hp", (0, 1.0, 50.0, 6000.0, 3000.0, 48000.0))):
        nt = getattr(f, "init_" + kind)(*spec)
        for h in hs:
            assert getattr(L, "rfo_fir_init_" + kind)(h, *spec) == nt
        co = np.zeros(80, dtype=np.float32)
        L.rfo_fir_coef(hs[0], P(co))
        assert bits_equal(f.taps(), co[:nt])
        for n in (1000, 29, 8192, 3):        # the summation start rotates with the running sample count
            x = rng.standard_normal((rows, n)).astype(np.float32)
            y = f.process_real(x)
            z = rng.standard_normal((rows, n, 2)).astype(np.float32)
            for r, h in enumerate(hs):
                ref = x[r].copy()
                L.rfo_fir_process_real(h, P(ref), n)
                assert bits_equal(y[r], ref), ("real", spec, r, n)
        for n in (777, 2048):
            a, b = rng.standard_normal((2, rows, n)).astype(np.float32)
            ya, yb = f.process_two(a, b)
            for r, h in enumerate(hs):
                ra, rb = a[r].copy(), b[r].copy()
                L.rfo_fir_process_two(h, P(ra), P(rb), n)
                assert bits_equal(ya[r], ra) and bits_equal(yb[r], rb), ("two", spec, r, n)
        for n in (600, 75):                  # complex Process shares the delay line with ProcessTwo
            z = rng.standard_normal((rows, n, 2)).astype(np.float32)
            yz = f.process_complex(z)
            for r, h in enumerate(hs):
                rz = z[r].copy()
                L.rfo_fir_process_complex(h, P(rz), n)
                assert bits_equal(yz[r], rz), ("complex", spec, r, n)
    taps = rng.standard_normal(52).astype(np.float32)
    f.init_const(taps, 31250.0)
    for h in hs:
        L.rfo_fir_init_const(h, taps.size, P(taps), 31250.0)
    for n in (2048, 100):
        x = rng.standard_normal((rows, n)).astype(np.float32)
        y = f.process_real(x)
        for r, h in enumerate(hs):
            ref = x[r].copy()
            L.rfo_fir_process_real(h, P(ref), n)
            assert bits_equal(y[r], ref), ("const", r, n)
    for h in hs:
        L.rfo_fir_destroy(h)
    f.close()


@pytest.mark.parametrize("ntaps,rows", [(8, 1), (9, 33), (10, 3), (11, 70), (29, 33), (44, 2), (50, 65), (75, 31), (74, 4), (73, 5), (5, 3)])
def test_fir_tap_count_sweep(rfm, port, ntaps, rows):
    """the lane = stream rotating FIR (groups of four outputs, rotation cycles of `ntaps` outputs, one-output path for
    the N mod 4 leftovers and the block edges) for every residue of N mod 4, row counts that leave partial warps, and
    call lengths around the cycle length; 5 taps takes the thread-per-output form.  InitConstFir only sets the real
    path's taps (FirFilter.cpp:302-320), so this sweep is the real Process."""
    rng = np.random.default_rng(100 + ntaps)
    L = port.lib()
    f = rfm.FirFilterBatch(rows, max_len=4096)
    hs = [L.rfo_fir_create() for _ in range(rows)]
    taps = rng.standard_normal(ntaps).astype(np.float32)
    f.init_const(taps, 48000.0)
    for h in hs:
        L.rfo_fir_init_const(h, taps.size, P(taps), 48000.0)
    for n in (1, ntaps - 1, ntaps, ntaps + 1, 4 * ntaps + 3, 1000, 2, 4096):
        x = rng.standard_normal((rows, n)).astype(np.float32)
        y = f.process_real(x)
        for r, h in enumerate(hs):
            ref = x[r].copy()
            L.rfo_fir_process_real(h, P(ref), n)
            assert bits_equal(y[r], ref), ("real", ntaps, r, n)
    for h in hs:
        L.rfo_fir_destroy(h)
    f.close()


@pytest.mark.parametrize("spec,rows", [((0, 1.0, 60.0, 15000.0, 21000.0, 48000.0), 33), ((0, 1.0, 40.0, 2400.0, 3120.0, 31250.0), 70),
                                       ((0, 2.0, 50.0, 5000.0, 9000.0, 48000.0), 3), ((0, 1.0, 30.0, 1000.0, 6000.0, 48000.0), 65),
                                       ((0, 0.5, 45.0, 3000.0, 5500.0, 48000.0), 1)])
def test_fir_two_channel_and_complex_interleaved(rfm, port, spec, rows):
    """ProcessTwo and the complex Process interleaved (they share one delay line and the rotation index,
    FirFilter.cpp:330-413), short and long calls, for several Kaiser designs (different tap counts).  The real
    Process is not mixed in: it has its own delay line but the SAME rotation index, so in the reference a real call
    after complex calls finds stale slots in its line -- a use the chain never makes and the library does not
    reproduce."""
    rng = np.random.default_rng(int(spec[3]))
    L = port.lib()
    f = rfm.FirFilterBatch(rows, max_len=4096)
    hs = [L.rfo_fir_create() for _ in range(rows)]
    nt = f.init_lp(*spec)
    for h in hs:
        assert L.rfo_fir_init_lp(h, *spec) == nt
    for n in (1, nt - 1, nt + 1, 4 * nt + 3, 1000, 2, 4096):
        a, b = rng.standard_normal((2, rows, n)).astype(np.float32)
        ya, yb = f.process_two(a, b)
        z = rng.standard_normal((rows, n, 2)).astype(np.float32)
        yz = f.process_complex(z)
        for r, h in enumerate(hs):
            ra, rb = a[r].copy(), b[r].copy()
            L.rfo_fir_process_two(h, P(ra), P(rb), n)
            rz = z[r].copy()
            L.rfo_fir_process_complex(h, P(rz), n)
            assert bits_equal(ya[r], ra) and bits_equal(yb[r], rb), ("two", nt, r, n)
            assert bits_equal(yz[r], rz), ("complex", nt, r, n)
    for h in hs:
        L.rfo_fir_destroy(h)
    f.close()


@pytest.mark.parametrize("rate", ["1.0M", "2.4M"])
def test_rds_processor_from_baseband(rfm, port, rate):
    """cRDSRxSignalProcessor on its own: fed with the demodulated baseband of the oracle decoder (its `baseband` tap is
    exactly what FmDecode.cpp:436 passes to RDSProcess::Process), it must emit the decoder's RDS bits and groups."""
    fs, ds, blk = RATES[rate]
    nblk = 6
    streams = [station(rate, nblk, stream_id=s)[0] for s in range(2)]
    oracles = [port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds) for _ in streams]
    rp = rfm.RdsProcessorBatch(len(streams), fs / ds, max_len=blk)
    for b in range(nblk):
        bbs = []
        for o, iq in zip(oracles, streams):
            o.process_u8(iq[b * blk:(b + 1) * blk])
            bbs.append(o.tap("baseband"))
        rp.process(np.stack(bbs))
    for s, o in enumerate(oracles):
        bits = o.take_bits()
        assert bits.size > 150
        assert np.array_equal(rp.take_bits(s), bits)
        groups = o.take_groups()
        assert np.array_equal(rp.take_groups(s), groups) and (rate != "1.0M" or groups.shape[0] >= 1)
    rp.close()


def test_downsample_filter_both_live_forms(rfm, port):
    """cDownsampleFilter as cFmDecoder builds it (FmDecode.cpp:257-273): the complex integer decimator in front of the
    demodulator and the real fractional resampler behind it; several calls (positions and histories carried)."""
    rng = np.random.default_rng(23)
    L = port.lib()
    rows = 3
    # complex, integer: order 8 * ds, cutoff 0.6 / ds (1.0 MS/s: ds 4; 2.4 MS/s: ds 11)
    for ds in (4, 11, 1):
        order, cutoff = 8 * ds, 0.6 / ds
        f = rfm.DownsampleFilterBatch(rows, order, cutoff, ds, True, max_len=8192)
        hs = [L.rfo_downsample_create(order, cutoff, float(ds), 1) for _ in range(rows)]
        co = np.zeros(1024, dtype=np.float32)
        k = L.rfo_downsample_coeff(hs[0], P(co))
        assert bits_equal(f.coefficients(), co[:k])
        for n in (8192, 1000, 4097, 8192):
            x = rng.standard_normal((rows, n, 2)).astype(np.float32)
            y = f.process_complex(x)
            for r, h in enumerate(hs):
                ref = np.zeros((n, 2), dtype=np.float32)
                m = L.rfo_downsample_process_complex(h, P(x[r].copy()), P(ref), n)
                assert y.shape[1] == m and bits_equal(y[r], ref[:m]), ("complex", ds, r, n)
        f.reset()
        for h in hs:
            L.rfo_downsample_reset(h)
        x = rng.standard_normal((rows, 2048, 2)).astype(np.float32)
        y = f.process_complex(x)
        for r, h in enumerate(hs):
            ref = np.zeros((2048, 2), dtype=np.float32)
            m = L.rfo_downsample_process_complex(h, P(x[r].copy()), P(ref), 2048)
            assert bits_equal(y[r], ref[:m]), ("complex after Reset", ds, r)
            L.rfo_downsample_destroy(h)
        f.close()
    # real, fractional: order int(Fb / 1000), cutoff 0.75 * 48000 / Fb ... as FmDecode.cpp:263-267
    for fb in (250000.0, 218181.8125):
        order = int(fb / 1000.0)
        ratio = fb / 48000.0
        cutoff = 15000.0 / fb
        f = rfm.DownsampleFilterBatch(rows, order, cutoff, ratio, False, max_len=16384)
        hs = [L.rfo_downsample_create(order, cutoff, ratio, 0) for _ in range(rows)]
        for n in (16384, 5952, 1000, 16384):
            x = rng.standard_normal((rows, n)).astype(np.float32)
            y = f.process_real(x)
            for r, h in enumerate(hs):
                ref = np.zeros(n, dtype=np.float32)
                m = L.rfo_downsample_process_real(h, P(x[r].copy()), P(ref), n)
                assert y.shape[1] == m and bits_equal(y[r], ref[:m]), ("real", fb, r, n)
        for h in hs:
            L.rfo_downsample_destroy(h)
        with pytest.raises(rfm.RadioFmError):
            f.process_complex(np.zeros((rows, 4096, 2), dtype=np.float32))
        f.close()


def test_downsample_filter_real_integer(rfm, port):
    """cDownsampleFilter real input + integer factor (DownConvert.cpp:164-192; no caller in the reference, provided for
    users of the class): call lengths that leave every residue of the output position, Reset in between."""
    rng = np.random.default_rng(29)
    L = port.lib()
    rows = 3
    for order, ds, cutoff in ((32, 4, 0.15), (40, 5, 0.12), (24, 1, 0.4), (130, 7, 0.07)):
        f = rfm.DownsampleFilterBatch(rows, order, cutoff, ds, True, max_len=8192)
        hs = [L.rfo_downsample_create(order, cutoff, float(ds), 1) for _ in range(rows)]
        for rep in range(2):
            for n in (8192, 1001, 4097, 130, 8191):
                x = rng.standard_normal((rows, n)).astype(np.float32)
                y = f.process_real(x)
                for r, h in enumerate(hs):
                    ref = np.zeros(n, dtype=np.float32)
                    m = L.rfo_downsample_process_real(h, P(x[r].copy()), P(ref), n)
                    assert y.shape[1] == m and bits_equal(y[r], ref[:m]), ("real integer", order, ds, rep, r, n)
            f.reset()
            for h in hs:
                L.rfo_downsample_reset(h)
        for h in hs:
            L.rfo_downsample_destroy(h)
        f.close()

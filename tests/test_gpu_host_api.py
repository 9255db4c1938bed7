"""The reference-facing C++ classes (pvr.rtl.radiofm_b200/host/FmDecode.h, FreqShift.h) driven from a C++ program the
way cRadioReceiver drives the reference (tests/cpp/host_api_driver.cpp), and the cFreqShift primitive; compared with
the oracle.  GPU box only."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, RATES, bits_equal, station

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostapi")
    exe = d / "host_api_driver"
    lib_dir = os.path.join(ROOT, "pvr.rtl.radiofm_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(lib_dir, "host"),
                           os.path.join(ROOT, "tests", "cpp", "host_api_driver.cpp"), "-o", str(exe),
                           "-L", lib_dir, "-lradiofm_b200", f"-Wl,-rpath,{lib_dir}"])
    return str(exe), d


def test_cfmdecoder_class_matches_oracle(driver, port):
    exe, d = driver
    fs, ds, blk = RATES["1.0M"]
    nblk = 10
    iq, sent = station("1.0M", nblk)
    (d / "iq.bin").write_bytes(iq.tobytes())
    out = subprocess.run([exe, str(d / "iq.bin"), str(fs), str(-0.15 * fs), str(ds), str(blk), str(d / "audio.bin"),
                          str(d / "groups.bin"), "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    ref_audio = np.concatenate([o.process_cf32(port.u8_to_cf32(iq[b * blk:(b + 1) * blk])) for b in range(nblk)])
    audio = np.fromfile(d / "audio.bin", dtype=np.float32)
    assert bits_equal(audio, ref_audio)
    groups = np.fromfile(d / "groups.bin", dtype=np.uint16).reshape(-1, 4)
    assert np.array_equal(groups, o.take_groups()) and len(groups) >= 5
    words = out.stdout.split()
    st = o.status()
    assert int(words[3]) == 3                                         # stereo blocks 8, 9, 10
    assert np.float32(words[7]) == st["if_level"] and np.float32(words[9]) == st["bb_level"]
    assert np.float32(words[11]) == st["pilot_level"] and np.float32(words[13]) == st["tuning_offset"]


def test_unchanged_caller_gets_uecp_frames_through_the_receiver(driver, port):
    """cFmDecoder built with a receiver and no sink (what RadioReceiver.cpp:296-300 does): the frames reach
    cRadioReceiver::AddUECPDataFrame from inside ProcessStream, as in the reference (RDSProcess.cpp:312,355,
    RDSGroupDecoder.cpp:979-991) -- including the re-announcement after a Reset() (RDSProcess.cpp:92)."""
    from oracle import uecp_port
    exe, d = driver
    fs, ds, blk = RATES["1.0M"]
    nblk, cut = 14, 7
    iq, _ = station("1.0M", 7)
    iq2 = np.concatenate([iq, iq])                      # the same station again after the reset
    (d / "iq3.bin").write_bytes(iq2.tobytes())
    out = subprocess.run([exe, str(d / "iq3.bin"), str(fs), str(-0.15 * fs), str(ds), str(blk), str(d / "audio3.bin"),
                          str(d / "groups3.bin"), "0", str(d / "uecp3.bin"), str(cut)], capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0, out.stderr
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    u = uecp_port.OracleGroupDecoder()
    want, audio = [], []
    for b in range(nblk):
        audio.append(o.process_cf32(port.u8_to_cf32(iq2[b * blk:(b + 1) * blk])))
        want += u.decode(o.take_groups())
        if b + 1 == cut:
            o.reset()
            u.reset()
    assert bits_equal(np.fromfile(d / "audio3.bin", dtype=np.float32), np.concatenate(audio))
    raw = (d / "uecp3.bin").read_bytes()
    got, i = [], 0
    while i < len(raw):
        ln = raw[i] | (raw[i + 1] << 8)
        got.append(raw[i + 2:i + 2 + ln])
        i += 2 + ln
    assert len(want) >= 8 and got == want


def test_cfreqshift_then_decoder(driver, port):
    """cFreqShift::Process in front of the decoder (the wideband composition): shift the capture by +50 kHz with the
    GPU mixer, decode with tuning offset -100 kHz; must equal the oracle's cFreqShift + chain."""
    exe, d = driver
    fs, ds, blk = RATES["1.0M"]
    nblk = 3
    iq, _ = station("1.0M", nblk)
    (d / "iq2.bin").write_bytes(iq.tobytes())
    out = subprocess.run([exe, str(d / "iq2.bin"), str(fs), str(-0.10 * fs), str(ds), str(blk), str(d / "audio2.bin"),
                          str(d / "groups2.bin"), "50000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    import ctypes as C
    L = port.lib()
    h = L.rfo_freqshift_create(50000.0, fs)
    o = port.OracleFmDecoder(fs, -0.10 * fs, downsample=ds)
    ref = []
    for b in range(nblk):
        x = port.u8_to_cf32(iq[b * blk:(b + 1) * blk])
        L.rfo_freqshift_process(h, x.ctypes.data_as(C.POINTER(C.c_float)), blk)
        ref.append(o.process_cf32(x))
    L.rfo_freqshift_destroy(h)
    assert bits_equal(np.fromfile(d / "audio2.bin", dtype=np.float32), np.concatenate(ref))


def test_freqshift_batch_primitive(rfm, port):
    """rows x cFreqShift: cf32 in place over two calls (carried phase), the fused u8 form on a shared capture
    (wideband: every station mixes the same samples), reset."""
    import ctypes as C
    rng = np.random.default_rng(4)
    freqs = np.array([37000.0, -200000.0, 0.0, 1234.5, 450000.0], dtype=np.float32)
    fs, n = 1.0e6, 6000
    L = port.lib()
    x = (rng.standard_normal((freqs.size, 2 * n, 2)) * 0.3).astype(np.float32)
    fsb = rfm.FreqShiftBatch(freqs, fs, max_len=n)
    got = np.concatenate([fsb.process_cf32(x[:, :n]), fsb.process_cf32(x[:, n:])], axis=1)
    for r, f in enumerate(freqs):
        h = L.rfo_freqshift_create(float(f), fs)
        y = x[r].copy()
        L.rfo_freqshift_process(h, y.ctypes.data_as(C.POINTER(C.c_float)), n)
        L.rfo_freqshift_process(h, y[n:].ctypes.data_as(C.POINTER(C.c_float)), n)
        L.rfo_freqshift_destroy(h)
        assert bits_equal(got[r], y), f
    u8 = rng.integers(0, 256, (n, 2), dtype=np.uint8)
    fsb.reset()
    got = fsb.process_u8(u8, shared_capture=True)
    for r, f in enumerate(freqs):
        h = L.rfo_freqshift_create(float(f), fs)
        y = port.u8_to_cf32(u8)
        L.rfo_freqshift_process(h, y.ctypes.data_as(C.POINTER(C.c_float)), n)
        L.rfo_freqshift_destroy(h)
        assert bits_equal(got[r], y), f


def test_primitive_classes_match_oracle(tmp_path, port):
    """cIirFilter, cFirFilter, CRDSDownConvert and cRDSRxSignalProcessor (host/*.h) driven from C++ through the reference's
    member signatures (tests/cpp/host_primitives_driver.cpp); every output file equals the oracle's, bit for bit."""
    import ctypes as C
    P = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    lib_dir = os.path.join(ROOT, "pvr.rtl.radiofm_b200")
    exe = tmp_path / "host_primitives_driver"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(lib_dir, "host"),
                           os.path.join(ROOT, "tests", "cpp", "host_primitives_driver.cpp"), "-o", str(exe),
                           "-L", lib_dir, "-lradiofm_b200", f"-Wl,-rpath,{lib_dir}"])
    # input: the demodulated baseband of a 1.0 MS/s station (250 kS/s, carries pilot + RDS), 8 blocks of 16384
    fs, ds, blk = RATES["1.0M"]
    iq, _ = station("1.0M", 8)
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    bb = []
    for b in range(8):
        o.process_u8(iq[b * blk:(b + 1) * blk])
        bb.append(o.tap("baseband"))
    x = np.concatenate(bb).astype(np.float32)
    n, bl = x.size, 16384
    (tmp_path / "in.f32").write_bytes(x.tobytes())
    out = subprocess.run([str(exe), str(tmp_path / "in.f32"), str(n), str(bl), str(tmp_path / "o")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    rd = lambda name, dt=np.float32: np.fromfile(tmp_path / f"o_{name}", dtype=dt)
    L = port.lib()
    half = n // 2
    # cIirFilter
    h = L.rfo_iir_create()
    L.rfo_iir_init(h, 3, 19000.0, 5.0, 48000.0)
    a, b = x[:half].copy(), x[half:].copy()
    L.rfo_iir_process_two(h, P(a), P(b), half)
    assert bits_equal(rd("iir_a.f32"), a) and bits_equal(rd("iir_b.f32"), b)
    L.rfo_iir_init(h, 2, 1187.5, 500.0, 31250.0)
    r = x.copy()
    L.rfo_iir_process_real(h, P(r), n)
    L.rfo_iir_destroy(h)
    assert bits_equal(rd("iir_r.f32"), r)
    # cFirFilter
    h = L.rfo_fir_create()
    nt = L.rfo_fir_init_lp(h, 0, 1.0, 60.0, 15000.0, 21000.0, 48000.0)
    a, b = x[:half].copy(), x[half:].copy()
    L.rfo_fir_process_two(h, P(a), P(b), half)
    assert f"fir_taps {nt}" in out.stdout
    assert bits_equal(rd("fir_a.f32"), a) and bits_equal(rd("fir_b.f32"), b)
    L.rfo_fir_init_lp(h, 0, 1.0, 40.0, 2400.0, np.float32(1.3) * np.float32(2400.0), 31250.0)
    z = x.copy()
    L.rfo_fir_process_complex(h, P(z), half)
    assert bits_equal(rd("fir_z.f32"), z)
    nh = L.rfo_fir_init_hp(h, 0, 1.0, 50.0, 6000.0, 3000.0, 48000.0)   # InitHPFilter + the two-buffer complex Process
    z = x.copy()
    L.rfo_fir_process_complex(h, P(z), half)
    L.rfo_fir_destroy(h)
    assert f"fir_hp_taps {nh}" in out.stdout and bits_equal(rd("fir_hp.f32"), z)
    # CRDSDownConvert (block-wise, like the driver)
    h = L.rfo_rdsdc_create()
    rate = L.rfo_rdsdc_set_data_rate(h, 250000.0, 8000.0)
    L.rfo_rdsdc_set_frequency(h, -57000.0)
    zz = x.copy().reshape(-1, 2)
    ys = []
    for i in range(0, half - bl + 1, bl):
        zi = zz[i:i + bl].copy()
        y = np.zeros_like(zi)
        k = L.rfo_rdsdc_process(h, bl, P(zi), P(y))
        ys.append(y[:k])
    L.rfo_rdsdc_destroy(h)
    assert bits_equal(rd("dc.f32").reshape(-1, 2), np.concatenate(ys))
    assert f"dc_rate {float(np.float32(rate)):.9g}" in out.stdout
    # cDownsampleFilter
    h = L.rfo_downsample_create(32, 0.6 / 4, 4.0, 1)
    ys = []
    for i in range(0, half - bl + 1, bl):
        y = np.zeros((bl, 2), dtype=np.float32)
        k = L.rfo_downsample_process_complex(h, P(zz[i:i + bl].copy()), P(y), bl)
        ys.append(y[:k])
    L.rfo_downsample_destroy(h)
    assert bits_equal(rd("ds_c.f32").reshape(-1, 2), np.concatenate(ys))
    h = L.rfo_downsample_create(250, 15000.0 / 250000.0, 250000.0 / 48000.0, 0)
    ys = []
    for i in range(0, n - bl + 1, bl):
        y = np.zeros(bl, dtype=np.float32)
        k = L.rfo_downsample_process_real(h, P(x[i:i + bl].copy()), P(y), bl)
        ys.append(y[:k])
    L.rfo_downsample_destroy(h)
    assert bits_equal(rd("ds_r.f32"), np.concatenate(ys))
    # cRDSRxSignalProcessor: the decoder's own RDS output for this baseband
    bits, groups = o.take_bits(), o.take_groups()
    assert np.array_equal(rd("rds_bits.u8", np.uint8), bits) and bits.size > 400
    assert np.array_equal(rd("rds_groups.u16", np.uint16).reshape(-1, 4), groups) and len(groups) >= 3

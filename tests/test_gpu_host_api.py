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

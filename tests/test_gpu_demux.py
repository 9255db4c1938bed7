"""the caller of the hot path -- IQ block queue + DemuxRead packetiser
(RadioReceiver.cpp:420-542) -- through the C ABI (rfm_demux_*) against the oracle's restatement (oracle/demux_port.py):
same packet order, ids, sizes, pts / duration (exact doubles), bit-identical audio, identical UECP bytes."""
from __future__ import annotations

import os
import threading

import numpy as np
import pytest

from conftest import ROOT, station

pytestmark = pytest.mark.gpu


def read_all(dm):
    out = []
    while True:
        p = dm.read()
        if p is None:
            return out
        out.append(p)


def same_packets(got, want):
    assert [p[0] for p in got] == [p[0] for p in want]
    for a, b in zip(got, want):
        assert a[1] == b[1] and a[2] == b[2], (a[:3], b[:3])
        if a[0] == 1:
            assert a[3].size == b[3].size and np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))
        elif a[0] == 2:
            assert a[3] == b[3]


def test_demux_packets_match_the_oracle(rfm):
    from oracle import demux_port
    fs, ds, blk, nblk = 1.0e6, 4, 65536, 8
    iq, _ = station("1.0M", nblk)
    dm = rfm.Demux(fs, -0.15 * fs, downsample=ds, max_block_len=blk)
    om = demux_port.OracleDemux(fs, -0.15 * fs, downsample=ds)
    with pytest.raises(rfm.RadioFmError):
        dm.signal_status()                     # GetSignalStatus is false until the stream-change packet went out
    for b in range(nblk):                      # everything queued up front: read-ahead is active throughout
        dm.write_u8(iq[b * blk:(b + 1) * blk])
        om.write_u8(iq[b * blk:(b + 1) * blk])
    assert dm.queued_samples() == nblk * blk
    dm.end()
    got, want = read_all(dm), []
    while True:
        p = om.read()
        if p is None:
            break
        want.append(p)
    same_packets(got, want)
    ids = [p[0] for p in got]
    assert ids[0] == rfm.DEMUX_STREAMCHANGE and ids.count(1) == nblk and ids.count(2) >= 2
    assert got[1][1] == 1000000.0              # first pts = STREAM_TIME_BASE
    assert np.float32(dm.audio_level()) == om.audio_level
    il, al, stereo = dm.signal_status()
    assert stereo and np.isfinite(il) and np.float32(al) == np.float32(20 * np.log10(float(om.audio_level)) + 3.01)
    assert dm.queued_samples() == 0 and dm.read() is None


def test_demux_with_a_live_producer_thread(rfm):
    """source thread writes while the demux thread reads (blocks arrive one at a time: no read-ahead possible at
    first, then a burst); ragged last block; stream change re-armed in the middle"""
    from oracle import demux_port
    fs, ds, blk, nblk = 1.2e6, 5, 65520, 5
    iq, _ = station("1.2M", nblk)
    lens = [blk, blk, blk, blk, blk // 2 // 80 * 80]
    dm = rfm.Demux(fs, -0.15 * fs, downsample=ds, max_block_len=blk)
    om = demux_port.OracleDemux(fs, -0.15 * fs, downsample=ds)
    gate = threading.Event()

    def producer():
        off = 0
        for i, n in enumerate(lens):
            if i == 2:
                gate.wait(10.0)
            dm.write_u8(iq[off:off + n])
            off += n
        dm.end()

    off = 0
    for n in lens:
        om.write_u8(iq[off:off + n])
        off += n
    th = threading.Thread(target=producer)
    th.start()
    got = []
    while True:
        p = dm.read()
        if p is None:
            break
        got.append(p)
        if sum(1 for q in got if q[0] == 1) == 2:
            gate.set()
    th.join()
    want = []
    while True:
        p = om.read()
        if p is None:
            break
        want.append(p)
    same_packets(got, want)
    assert [p[3].size for p in got if p[0] == 1][-1] < [p[3].size for p in got if p[0] == 1][0]
    dm.set_stream_change()
    assert dm.read() is not None and dm.read() is None


def test_source_callback_block_rule_and_short_reads(rfm):
    """cRtlSdrSource without librtlsdr: block-length rule, async-read callback, short reads dropped"""
    from oracle import demux_port
    L = rfm.lib()
    assert [L.rfm_source_block_length(v) for v in (0, 4095, 4096, 65536, 70000, 1 << 21)] == [4096, 4096, 4096, 65536, 69632, 1 << 20]
    fs, ds, blk, nblk = 1.0e6, 4, 65536, 3
    iq, _ = station("1.0M", nblk)
    dm = rfm.Demux(fs, -0.15 * fs, downsample=ds, max_block_len=blk)
    om = demux_port.OracleDemux(fs, -0.15 * fs, downsample=ds)
    dm.source_cb(iq[:blk])                     # no block length configured yet: dropped
    dm.set_source_block_length(blk + 100)      # 65636 -> 65536
    for b in range(nblk):
        dm.source_cb(iq[b * blk:(b + 1) * blk])
        if b == 0:
            dm.source_cb(iq[:blk // 2])        # short read: samples lost, nothing queued
        om.write_u8(iq[b * blk:(b + 1) * blk])
    dm.end()
    assert dm.short_reads() == 2 and dm.queued_samples() == nblk * blk
    want = []
    while True:
        p = om.read()
        if p is None:
            break
        want.append(p)
    same_packets(read_all(dm), want)


def test_rtlsdr_source_adapter_against_a_stand_in_library(rfm, tmp_path):
    """cRtlSdrSource end to end: librtlsdr is bound with dlopen, so a stand-in library
    (tests/cpp/fake_rtlsdr.c, a file player) exercises the adapter -- Open, Configure's call sequence and block-length
    rule, the reader thread around rtlsdr_read_async, the short read it must drop, the restart after a failed read,
    Close -- and the packets that come out of the demux are the oracle's."""
    import subprocess
    from oracle import demux_port
    so = tmp_path / "libfake_rtlsdr.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "fake_rtlsdr.c"), "-o", str(so)])
    fs, ds, blk, nblk = 1.0e6, 4, 65536, 4
    iq, _ = station("1.0M", nblk)
    (tmp_path / "iq.bin").write_bytes(iq.tobytes())
    os.environ["FAKE_RTLSDR_FILE"] = str(tmp_path / "iq.bin")
    os.environ["FAKE_RTLSDR_LOG"] = str(tmp_path / "calls.log")
    os.environ["FAKE_RTLSDR_FAIL"] = "1"
    try:
        dm = rfm.Demux(fs, -0.15 * fs, downsample=ds, max_block_len=blk)
        src = rfm.RtlSdrSource(dm, str(so), 0)
        src.configure(1000000, 98500000, tuner_gain=297, block_length=blk + 1000, agcmode=True)   # 66536 -> 65536
        assert (src.block_length, src.sample_rate, src.frequency, src.tuner_gain) == (blk, 1000000, 98500000, 297)
        om = demux_port.OracleDemux(fs, -0.15 * fs, downsample=ds)
        for b in range(nblk):
            om.write_u8(iq[b * blk:(b + 1) * blk])
        want = []
        while True:
            p = om.read()
            if p is None:
                break
            want.append(p)
        got = []
        while sum(1 for q in got if q[0] == 1) < nblk:     # the file player never ends by itself: read what it holds
            p = dm.read()
            assert p is not None
            got.append(p)
        src.set_frequency(101300000)
        assert src.frequency == 101300000 and src.restarts == 1 and dm.short_reads() == 1
        src.close()                                          # cancel_async, join, EndDataBuffer
        while True:
            p = dm.read()
            if p is None:
                break
            got.append(p)
        same_packets(got, want)
        calls = (tmp_path / "calls.log").read_text().split("\n")
        first = [c for c in calls if c][:8]
        assert first == ["open 0", "sample_rate 1000000", "center_freq 98500000", "gain_mode 1", "gain 297", "agc 1",
                         "reset_buffer", f"read_async 15 {2 * blk}"], first
        assert "close" in calls
        with pytest.raises(rfm.RadioFmError):
            rfm.RtlSdrSource(dm, str(tmp_path / "no_such_library.so"), 0)
    finally:
        for k in ("FAKE_RTLSDR_FILE", "FAKE_RTLSDR_LOG", "FAKE_RTLSDR_FAIL"):
            os.environ.pop(k, None)


def test_demux_and_source_adapter_against_the_compiled_reference_addon(rfm, tmp_path):
    """The product against the UNMODIFIED reference add-on itself (oracle/ref_addon_harness.cpp: cRadioReceiver +
    cRtlSdrSource + cFmDecoder + the RDS chain, every source file compiled in place): the device-configuration calls of
    OpenLiveStream / Configure, then every DemuxRead packet -- order, ids, pts / duration, audio bits, UECP bytes -- the
    audio level and the signal status in dB."""
    import subprocess
    from oracle import ref_addon
    if not ref_addon.available():
        pytest.skip("oracle/_ref/libradiofm_ref_addon.so not built")
    nblk = 8
    iq, _ = station("1.0M", nblk)
    addon = ref_addon.RefAddon(100.0e6)
    assert addon.open()
    par = addon.params()
    fs, off, ds, blk = par["if_rate"], par["tuning_offset"], par["downsample"], par["block_length"]
    # the source adapter, configured the way OpenLiveStream configures cRtlSdrSource, talks to its device identically
    so = tmp_path / "libfake_rtlsdr.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", os.path.join(ROOT, "tests", "cpp", "fake_rtlsdr.c"), "-o", str(so)])
    (tmp_path / "iq.bin").write_bytes(iq[:nblk * blk].tobytes())
    os.environ["FAKE_RTLSDR_FILE"] = str(tmp_path / "iq.bin")
    os.environ["FAKE_RTLSDR_LOG"] = str(tmp_path / "calls.log")
    try:
        dm = rfm.Demux(fs, off, downsample=ds, max_block_len=blk)
        src = rfm.RtlSdrSource(dm, str(so), 0)
        src.configure(int(fs), int(par["tuner_freq"]), tuner_gain=197, block_length=65536, agcmode=True)
        assert src.block_length == blk
        for b in range(nblk):
            addon.feed(iq[b * blk:(b + 1) * blk], short_read_after_next=(b == 0))   # the file player does the same
        want = addon.read_all()
        got = []
        while sum(1 for q in got if q[0] == 1) < nblk:
            p = dm.read()
            assert p is not None
            got.append(p)
        src.close()
        while True:
            p = dm.read()
            if p is None:
                break
            got.append(p)
        same_packets(got, want)
        calls = [c for c in (tmp_path / "calls.log").read_text().split("\n") if c]
        assert calls[:8] == addon.device_log()[:8]
        assert np.float32(dm.audio_level()) == addon.audio_level()
        il, al, stereo = dm.signal_status()
        sig = addon.signal()
        assert (np.float32(il), np.float32(al), bool(stereo)) == \
            (np.float32(sig["if_level_db"]), np.float32(sig["audio_level_db"]), sig["stereo"])
    finally:
        for k in ("FAKE_RTLSDR_FILE", "FAKE_RTLSDR_LOG"):
            os.environ.pop(k, None)
        addon.close()
        addon.destroy()

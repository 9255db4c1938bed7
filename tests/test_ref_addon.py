"""The UNMODIFIED reference add-on receive path -- cRadioReceiver::OpenLiveStream / DemuxRead over cRtlSdrSource,
cFmDecoder and the RDS chain, every source file compiled in place behind oracle/ref_addon_harness.cpp -- against the
oracle's restatement of it (oracle/demux_port.py).  This is what pins SURVEY.md 8f rows N2 / N4 and the UECP transport
framing: before it, RadioReceiver.cpp:387-542 and RTL_SDR_Source.cpp were only restated ("parity unpinned").
CPU only; needs oracle/_ref/libradiofm_ref_addon.so (built by __graft_entry__.build() where /root/reference exists;
the prebuilt file travels to the GPU box)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import station

from oracle import demux_port, ref_addon

pytestmark = pytest.mark.skipif(not ref_addon.available(), reason="oracle/_ref/libradiofm_ref_addon.so not built")

FS, DS, BLK = 1.0e6, 4, 65536


def level_db(x, add=None):
    """20 * log10(x) as RadioReceiver.cpp:551-552 computes it: float log10, float product, "+ 3.01" in double"""
    v = np.float32(20) * np.float32(np.log10(float(np.float32(x))))
    return v if add is None else np.float32(float(v) + add)


def port_packets(om):
    out = []
    while True:
        p = om.read()
        if p is None:
            return out
        out.append(p)


def same_packets(got, want):
    assert [p[0] for p in got] == [p[0] for p in want]
    for a, b in zip(got, want):
        assert a[1] == b[1] and a[2] == b[2], (a[:3], b[:3])            # pts / duration: exact doubles
        if a[0] == 1:
            assert a[3].size == b[3].size and np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))
        elif a[0] == 2:
            assert a[3] == bytes(b[3])


@pytest.fixture()
def addon():
    a = ref_addon.RefAddon(100.0e6)
    yield a
    a.close()
    a.destroy()


def test_open_live_stream_derives_the_decoder_and_source_parameters(addon):
    """OpenLiveStream (RadioReceiver.cpp:176-353): 1 MS/s, tuner 0.15 fs above the channel, down-sampling 4, gain 19.7 dB
    with AGC, the default 65536-sample block -- the parameters every N2 / N4 test of the product uses."""
    assert addon.reload_channels() == (1, 100.0e6)        # SaveChannelData -> LoadChannelData round trip
    assert addon.signal() is None                         # no decoder yet
    assert addon.open()
    assert addon.params() == dict(if_rate=FS, tuning_offset=-0.15 * FS, downsample=DS, block_length=BLK, tuner_freq=100.15e6)
    assert addon.device_log() == ["open 0", "sample_rate 1000000", "center_freq 100150000", "gain_mode 1", "gain 197", "agc 1",
                                  "reset_buffer", f"read_async 15 {2 * BLK}"]
    props = addon.stream_properties()
    assert [(p["pid"], p["channels"], p["sample_rate"], p["bits"], p["bit_rate"]) for p in props] == \
        [(1, 2, 48000, 32, 3072000), (2, 2, 48000, 32, 3072000)]
    assert addon.signal() is None                         # m_StreamChange still set (RadioReceiver.cpp:548)


def test_reference_addon_packets_equal_the_restatement(addon):
    nblk = 8
    iq, _ = station("1.0M", nblk)
    assert addon.open()
    om = demux_port.OracleDemux(FS, -0.15 * FS, downsample=DS)
    for b in range(nblk):
        addon.feed(iq[b * BLK:(b + 1) * BLK], short_read_after_next=(b == 1))   # the half-length buffer is dropped
        om.write_u8(iq[b * BLK:(b + 1) * BLK])
    got, want = addon.read_all(), port_packets(om)
    same_packets(got, want)
    ids = [p[0] for p in got]
    assert ids[0] == ref_addon.STREAMCHANGE and ids.count(1) == nblk and ids.count(2) >= 2
    assert got[1][1] == 1000000.0 and addon.queued_samples() == 0
    assert addon.audio_level() == om.audio_level
    sig = addon.signal()
    assert sig["stereo"] and np.float32(sig["audio_level_db"]) == level_db(om.audio_level, 3.01)
    assert np.float32(sig["if_level_db"]) == level_db(om.dec.status()["if_level"])
    assert sig["signal"] == int(2.5 * (sig["if_level_db"] + 40) * 656) and sig["snr"] == int((sig["audio_level_db"] + 100) * 656)


def test_stream_change_and_partial_feeds(addon):
    """SetStreamChange mid-stream goes out before anything else; bytes fed in arbitrary pieces come out as whole blocks"""
    nblk = 3
    iq, _ = station("1.0M", nblk)
    raw = iq[:nblk * BLK].reshape(-1)
    assert addon.open()
    om = demux_port.OracleDemux(FS, -0.15 * FS, downsample=DS)
    cuts = [0, 1000, 2 * BLK - 2, 2 * BLK + 4096, 5 * BLK, raw.size]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        addon.feed(raw[lo:hi])
    for b in range(nblk):
        om.write_u8(iq[b * BLK:(b + 1) * BLK])
    got = [addon.read(), addon.read()]
    want = [om.read(), om.read()]
    addon.set_stream_change()
    om.stream_change = True
    got += addon.read_all()
    want += port_packets(om)
    same_packets(got, want)
    assert [p[0] for p in got].count(ref_addon.STREAMCHANGE) == 2 and got[2][0] == ref_addon.STREAMCHANGE

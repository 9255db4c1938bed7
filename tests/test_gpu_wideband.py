"""Wideband front end (C5): stations of one shared 50 MS/s u8 capture through WidebandReceiver (device composition of
rfm_freqshift / rfm_downconvert / rfm_decoder) against the oracle composition of the reference classes
(oracle/wideband.py), bit for bit, for both mixers."""
import importlib

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu

FS, BLK, BPC = 50.0e6, 32000, 64
FREQS = [-1.0e6, 200000.0, 7.0e6]


@pytest.fixture(scope="module")
def capture(synth):
    return synth.make_wideband_u8(FS, 2 * BPC * BLK, FREQS)


@pytest.mark.parametrize("mixer", ["osc", "freqshift", "freqshift_unfused"])
def test_wideband_stations_bit_exact(rfm, port, capture, mixer):
    import torch
    wb_mod = importlib.import_module("radiofm_b200.wideband")
    from oracle.wideband import OracleStation
    wb = wb_mod.WidebandReceiver(torch, FREQS, FS, BLK, BPC, mixer=mixer, device=0)
    oracles = [OracleStation(f, FS, BLK, BPC, mixer=mixer.split("_")[0]) for f in FREQS]
    n_call = BPC * BLK
    peak = 0.0
    for c in range(2):
        cap = capture[c * n_call:(c + 1) * n_call]
        audio = wb.process_u8(cap)
        bb = wb.bb.cpu().numpy()
        for s, o in enumerate(oracles):
            ref_bb = o.baseband(cap)
            assert bits_equal(bb[s], ref_bb), (mixer, c, s, "decimated baseband")
            ref_audio = o.dec.process_cf32(ref_bb)
            assert bits_equal(audio[s], ref_audio), (mixer, c, s, "audio")
            peak = max(peak, float(np.max(np.abs(ref_audio))))
    assert peak > 0.02          # the stations really are demodulated (tones present)
    for s, o in enumerate(oracles):
        assert np.array_equal(wb.dec.take_bits(s), o.dec.take_bits())
        so, sd = o.dec.status(), wb.dec.status(s)
        assert all(np.float32(so[k]) == np.float32(sd[k]) for k in so), (so, sd)
    wb.close()


@pytest.mark.parametrize("mixer", ["osc", "freqshift"])
def test_wideband_100_stations_on_the_raster(rfm, port, mixer):
    """BASELINE.json configs[4] at its stated width: 100 stations on the 200 kHz raster f_k = k * 200 kHz, k = -50 .. 49,
    one shared 50 MS/s capture (made on the device, bench scaffolding: radiofm_b200.synth_device), two demodulator calls
    of carried state.  Eight stations -- both band edges (+-10 MHz), the centre, and five in between -- are compared bit
    for bit with the oracle composition of the reference classes; all hundred must be alive."""
    import torch
    wb_mod = importlib.import_module("radiofm_b200.wideband")
    synth_device = importlib.import_module("radiofm_b200.synth_device")
    from oracle.wideband import OracleStation
    freqs = [(k - 50) * 200000.0 for k in range(100)]
    checked = [0, 13, 37, 49, 50, 57, 72, 99]            # k = -50, -37, -13, -1, 0, 7, 22, 49
    n_call = BPC * BLK
    cap_d = synth_device.make_wideband_u8(torch, FS, 2 * n_call, freqs, torch.device("cuda", 0))
    cap = cap_d.cpu().numpy()
    wb = wb_mod.WidebandReceiver(torch, freqs, FS, BLK, BPC, mixer=mixer, device=0)
    oracles = {s: OracleStation(freqs[s], FS, BLK, BPC, mixer=mixer) for s in checked}
    for c in range(2):
        seg = cap[c * n_call:(c + 1) * n_call]
        audio = wb.process_u8(seg)
        bb = wb.bb.cpu().numpy()
        for s, o in oracles.items():
            ref_bb = o.baseband(seg)
            assert bits_equal(bb[s], ref_bb), (mixer, c, s, "decimated baseband")
            assert bits_equal(audio[s], o.dec.process_cf32(ref_bb)), (mixer, c, s, "audio")
        assert np.all(np.max(np.abs(audio), axis=1) > 1e-3), "every station produces audio"
    for s, o in oracles.items():
        assert np.array_equal(wb.dec.take_bits(s), o.dec.take_bits())
        so, sd = o.dec.status(), wb.dec.status(s)
        assert all(np.float32(so[k]) == np.float32(sd[k]) for k in so), (s, so, sd)
    wb.close()


def test_companion_stream_is_usable_and_placement_only(rfm, port, capture):
    """rfm_decoder_companion_stream: with and without an SM partition the receiver produces the same bits (the partition
    and the stream are placement, not arithmetic), and the handle is a live stream owned by the decoder."""
    import torch
    wb_mod = importlib.import_module("radiofm_b200.wideband")
    n_call = BPC * BLK
    outs = []
    for lanes_sms, n_slots in ((8, 3), (1, 2)):
        wb = wb_mod.WidebandReceiver(torch, FREQS, FS, BLK, BPC, mixer="freqshift", device=0, lanes_sms=lanes_sms, n_slots=n_slots)
        h = wb.dec.companion_stream()
        assert h != 0 and h == wb.dec.companion_stream()
        a = [wb.process_u8(capture[c * n_call:(c + 1) * n_call]).copy() for c in range(2)]
        outs.append((a, [wb.dec.take_bits(s) for s in range(len(FREQS))]))
        wb.close()
    for c in range(2):
        assert bits_equal(outs[0][0][c], outs[1][0][c])
    for s in range(len(FREQS)):
        assert np.array_equal(outs[0][1][s], outs[1][1][s])

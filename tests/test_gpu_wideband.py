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

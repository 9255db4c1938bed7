"""rfm_steps.cuh (the bodies of the lane kernels) compiled for the host with g++ and run sample by sample against the
oracle on the CPU: exact and branch-free ("fast", with the sticky replay flag) forms of the FM-demodulator PLL and
the 19 kHz pilot PLL.  Test-only shim; the product runs these functions on the GPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, RATES, bits_equal, station

SHIM = r'''
#include "rfm_steps.cuh"
using namespace rfm;
extern "C" {
// returns number of samples flagged `bad` by the fast form; out = baseband (exact form), out_fast likewise
unsigned run_demod(const float* z, unsigned n, const float* k5, float* out, float* out_fast) {
  DemodConst k = {k5[0], k5[1], k5[2], k5[3], k5[4]};
  DemodState a = {0, 0}, b = {0, 0};
  float dca = 0, dcb = 0; unsigned flagged = 0;
  for (unsigned i = 0; i < n; ++i) {
    demod_step(a, z[2*i], z[2*i+1], k);
    out[i] = demod_output(a.incr, dca, k.gain);
    bool bad = false; DemodState save = b;
    demod_step_fast(b, z[2*i], z[2*i+1], k, bad);
    if (bad) { ++flagged; b = save; demod_step(b, z[2*i], z[2*i+1], k); }
    out_fast[i] = demod_output(b.incr, dcb, k.gain);
  }
  return flagged;
}
unsigned run_pilot(const float* x, unsigned n, const float* k8, float freq0, float* out, float* out_fast, float* lvl) {
  PilotConstDev k = {k8[0], k8[1], k8[2], k8[3], k8[4], k8[5], k8[6], k8[7], 0};
  PilotState a = {0, freq0, 0, 0, 0, 0, 0, 1000.f}, b = a; unsigned flagged = 0;
  for (unsigned i = 0; i < n; ++i) {
    out[i] = pilot_step(a, x[i], k);
    bool bad = false; PilotState save = b;
    float o = pilot_step_fast(b, x[i], k, rfm_sincos_regs(), bad);
    if (bad) { ++flagged; b = save; o = pilot_step(b, x[i], k); }
    out_fast[i] = o;
  }
  lvl[0] = a.level; lvl[1] = b.level;
  return flagged;
}
// speculative-sincos form (pilot_step_spec): per-sample replay on `bad`; returns the number of flagged samples,
// mism[0] = samples where the UNFLAGGED speculative step differs from the exact one in any state word or output
unsigned run_pilot_spec(const float* x, unsigned n, const float* k8, float freq0, float* out_spec, unsigned* mism) {
  PilotConstDev k = {k8[0], k8[1], k8[2], k8[3], k8[4], k8[5], k8[6], k8[7], 0};
  PilotState a = {0, freq0, 0, 0, 0, 0, 0, 1000.f}, b = a; unsigned flagged = 0; mism[0] = 0;
  const SinCosRegs R = rfm_sincos_regs();
  PilotCarry sc; rfm_sincos(b.phase, &sc.ps, &sc.pc);
  for (unsigned i = 0; i < n; ++i) {
    const float oa = pilot_step(a, x[i], k);
    bool bad = false; PilotState save = b;
    float o = pilot_step_spec(b, sc, x[i], k, R, bad);
    float es, ec; rfm_sincos(b.phase, &es, &ec);
    if (bad) { ++flagged; b = save; o = pilot_step(b, x[i], k); rfm_sincos(b.phase, &sc.ps, &sc.pc); }
    else if (memcmp(&a, &b, sizeof(a)) != 0 || f2u(o) != f2u(oa) || f2u(es) != f2u(sc.ps) || f2u(ec) != f2u(sc.pc)) ++mism[0];
    out_spec[i] = o;
  }
  return flagged;
}
}
'''


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("stepshim")
    (d / "shim.cpp").write_text(SHIM)
    so = d / "shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I",
                           os.path.join(ROOT, "pvr.rtl.radiofm_b200", "csrc"), str(d / "shim.cpp"), "-o", str(so), "-lm"])
    L = C.CDLL(str(so))
    f32p = C.POINTER(C.c_float)
    L.run_demod.argtypes = [f32p, C.c_uint, f32p, f32p, f32p]
    L.run_pilot.argtypes = [f32p, C.c_uint, f32p, C.c_float, f32p, f32p, f32p]
    L.run_pilot_spec.argtypes = [f32p, C.c_uint, f32p, C.c_float, f32p, C.POINTER(C.c_uint)]
    L.run_demod.restype = L.run_pilot.restype = L.run_pilot_spec.restype = C.c_uint
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("rate", ["1.0M", "2.4M"])
def test_demod_steps_match_oracle(shim, port, synth, rate):
    fs, ds, blk = RATES[rate]
    rng = np.random.default_rng(8)
    noise = np.clip(np.rint(127.5 + 40 * rng.standard_normal((2 * blk, 2))), 0, 255).astype(np.uint8)
    silence = np.full((blk, 2), 128, dtype=np.uint8)
    silence[::2] = 127
    for name, iq in (("station", station(rate, 3)[0]), ("noise", noise), ("silence", silence)):
        o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
        z, bb = [], []
        for b in range(iq.shape[0] // blk):
            o.process_u8(iq[b * blk:(b + 1) * blk])
            z.append(o.tap("demod_in"))
            bb.append(o.tap("baseband"))
        z, bb = np.ascontiguousarray(np.concatenate(z)), np.concatenate(bb)
        c = o.constants()
        k5 = np.array([c[3], c[4], c[5], c[6], c[7]], dtype=np.float32)
        out, outf = np.zeros(z.shape[0], dtype=np.float32), np.zeros(z.shape[0], dtype=np.float32)
        flagged = shim.run_demod(_p(z), z.shape[0], _p(k5), _p(out), _p(outf))
        assert bits_equal(out, bb), name
        assert bits_equal(outf, bb), name
        if name == "station":
            assert flagged <= 4   # the replay path is for degenerate operands only (e.g. an exact zero)


def test_pilot_steps_match_oracle(shim, port):
    fs, ds, blk = RATES["1.0M"]
    iq, _ = station("1.0M", 4)
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    c = o.constants()
    k8 = np.array([c[9], c[10], c[11], c[12], c[13], c[14], c[15], c[17]], dtype=np.float32)
    bb, p38 = [], []
    for b in range(4):
        o.process_u8(iq[b * blk:(b + 1) * blk])
        bb.append(o.tap("baseband"))
        p38.append(o.tap("pilot38"))
    bb, p38 = np.ascontiguousarray(np.concatenate(bb)), np.concatenate(p38)
    for x in (bb, np.zeros_like(bb)):
        out, outf, lvl = np.zeros_like(x), np.zeros_like(x), np.zeros(2, dtype=np.float32)
        flagged = shim.run_pilot(_p(x), x.size, _p(k8), float(np.float32(c[16])), _p(out), _p(outf), _p(lvl))
        assert bits_equal(out, outf) and lvl[0] == lvl[1]
        if x is bb:
            assert bits_equal(out, p38) and flagged == 0
    # the stand-alone oracle primitive on a noisy pilot
    L = port.lib()
    rng = np.random.default_rng(11)
    n = 30000
    t = np.arange(n) / 250000.0
    pil = (0.1 * np.sin(2 * np.pi * 19000.0 * t) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    h = L.rfo_pilot_create(np.float32(19000.0 / 250000.0), np.float32(50.0 / 250000.0), np.float32(0.04))
    y = np.zeros(n, dtype=np.float32)
    L.rfo_pilot_process(h, _p(pil), _p(y), n)
    L.rfo_pilot_destroy(h)
    out, outf, lvl = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32), np.zeros(2, dtype=np.float32)
    shim.run_pilot(_p(pil), n, _p(k8), float(np.float32(c[16])), _p(out), _p(outf), _p(lvl))
    assert bits_equal(out, y) and bits_equal(outf, y)


def test_speculative_pilot_step_is_exact_or_flagged(shim, port):
    """pilot_step_spec (sincos predicted off the dependent chain, corrected by the phase difference) must equal
    pilot_step bit for bit -- state, output and carried sincos -- on every sample it does not flag; the flag rate in
    lock must stay negligible (a flagged tile is replayed with the direct routine on the GPU)."""
    fs, ds, blk = RATES["2.4M"]
    nblk = 24
    iq, _ = station("2.4M", nblk)
    o = port.OracleFmDecoder(fs, -0.15 * fs, downsample=ds)
    c = o.constants()
    k8 = np.array([c[9], c[10], c[11], c[12], c[13], c[14], c[15], c[17]], dtype=np.float32)
    bb, p38 = [], []
    for b in range(nblk):
        o.process_u8(iq[b * blk:(b + 1) * blk])
        bb.append(o.tap("baseband"))
        p38.append(o.tap("pilot38"))
    bb, p38 = np.ascontiguousarray(np.concatenate(bb)), np.concatenate(p38)
    rng = np.random.default_rng(5)
    noise = (0.3 * rng.standard_normal(200000)).astype(np.float32)
    t = np.arange(400000) / (fs / ds)
    drift = (0.1 * np.sin(2 * np.pi * (19000.0 + 8.0 * np.sin(2 * np.pi * 3.0 * t)) * t)
             + 0.02 * rng.standard_normal(t.size)).astype(np.float32)
    for name, x, max_flag_rate in (("station", bb, 2e-3), ("noise", noise, 1.0), ("zeros", np.zeros(50000, np.float32), 1.0),
                                   ("drifting pilot", drift, 0.05)):
        out, mism = np.zeros_like(x), C.c_uint(0)
        flagged = shim.run_pilot_spec(_p(x), x.size, _p(k8), float(np.float32(c[16])), _p(out), C.byref(mism))
        assert mism.value == 0, (name, mism.value)
        assert flagged <= max_flag_rate * x.size, (name, flagged, x.size)
        if name == "station":
            assert bits_equal(out, p38)
            locked = shim.run_pilot_spec(_p(x[-40000:]), 40000, _p(k8), float(np.float32(c[16])), _p(out[:40000]), C.byref(mism))
            print("flagged", flagged, "of", x.size)

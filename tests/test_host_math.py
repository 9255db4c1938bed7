"""rfm_math.cuh is __host__ __device__: compile the very same header with g++ (test-only shim, never shipped)
and pin its restated libm routines on the CPU: rfm_atan2f == glibc atan2f bit for bit; rfm_sincos ==
float(sin/cos(double)) (what x87 fsincos -> float gives); the fmod helpers == libm."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SHIM = r'''
#include "rfm_math.cuh"
#include <math.h>
extern "C" {
unsigned long long cmp_atan2f(const float* y, const float* x, unsigned n) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n; ++i) {
    float a = rfm::rfm_atan2f(y[i], x[i]), b = atan2f(y[i], x[i]);
    bad += rfm::f2u(a) != rfm::f2u(b) && !(a != a && b != b);
  }
  return bad;
}
unsigned long long cmp_sincos(const float* p, unsigned n) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n; ++i) {
    float s, c; rfm::rfm_sincos(p[i], &s, &c);
    bad += rfm::f2u(s) != rfm::f2u((float)sin((double)p[i]));
    bad += rfm::f2u(c) != rfm::f2u((float)cos((double)p[i]));
  }
  return bad;
}
unsigned long long cmp_wraps(const float* p, unsigned n) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n; ++i) {
    float a = p[i];
    if ((double)a >= RFM_K_2PI) a = (float)fmod((double)a, RFM_K_2PI);   // FmDecode.cpp:404-409
    while (a < 0.0f) a = (float)(a + RFM_K_2PI);
    bad += rfm::f2u(rfm::rfm_wrap_demod(p[i])) != rfm::f2u(a);
    float b = p[i];
    if (b > RFM_K_2PI) b = (float)(b - RFM_K_2PI);                        // FmDecode.cpp:203-205
    if (p[i] > 0.0f) bad += rfm::f2u(rfm::rfm_wrap_pilot(p[i])) != rfm::f2u(b);
  }
  return bad;
}
unsigned long long cmp_sincos_generic(const float* p, unsigned n) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n; ++i) {
    float s, c; rfm::rfm_sincos_generic(p[i], &s, &c);
    bad += rfm::f2u(s) != rfm::f2u((float)sin((double)p[i]));
    bad += rfm::f2u(c) != rfm::f2u((float)cos((double)p[i]));
  }
  return bad;
}
unsigned long long cmp_atan2f_generic(const float* y, const float* x, unsigned n) {
  unsigned long long bad = 0;
  for (unsigned i = 0; i < n; ++i) {
    float a = rfm::rfm_atan2f_generic(y[i], x[i]), b = atan2f(y[i], x[i]);
    bad += rfm::f2u(a) != rfm::f2u(b) && !(a != a && b != b);
  }
  return bad;
}
// integer-pipe conversions (rfm_f2d_bits / rfm_d2f_bits) and the routines built on them, against the conversion
// instructions: returns unflagged mismatches; *flagged gets the number of operands the bit forms refused
unsigned long long cmp_bit_conversions(const unsigned* fbits, const double* d, unsigned n, unsigned long long* flagged) {
  unsigned long long bad = 0; *flagged = 0;
  const rfm::SinCosRegs R = rfm::rfm_sincos_regs();
  for (unsigned i = 0; i < n; ++i) {
    bool f1 = false; const float x = rfm::u2f(fbits[i]);
    const double w = rfm::rfm_f2d_bits(x, f1); const double we = (double)x;
    if (f1) ++*flagged; else bad += memcmp(&w, &we, 8) != 0;
    bool f2 = false; const float nrw = rfm::rfm_d2f_bits(d[i], f2);
    if (f2) ++*flagged; else bad += rfm::f2u(nrw) != rfm::f2u((float)d[i]);
    const float ph = fabsf(x) < 6.9f ? fabsf(x) : (float)fmod(fabs((double)x), 6.9);
    if (ph == ph && ph != 0.0f) {
      bool f3 = false; float s, c, es, ec;
      rfm::rfm_sincos_core_b(ph, R, &s, &c, f3); rfm::rfm_sincos_core_a(ph, R, &es, &ec);
      if (f3) ++*flagged; else bad += rfm::f2u(s) != rfm::f2u(es) || rfm::f2u(c) != rfm::f2u(ec);
    }
  }
  return bad;
}
unsigned long long cmp_fmod(const float* p, unsigned n) {
  unsigned long long bad = 0;
  const float twopif = (float)RFM_K_2PI;
  for (unsigned i = 0; i < n; ++i) {
    bad += rfm::f2u(rfm::rfm_fmodf_small(p[i], twopif)) != rfm::f2u(fmodf(p[i], twopif));
    double x = fabs((double)p[i]);
    if (x < 2 * RFM_K_2PI) bad += rfm::rfm_fmod_2pi_small(x) != fmod(x, RFM_K_2PI);
  }
  return bad;
}
}
'''


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    d = tmp_path_factory.mktemp("mathshim")
    src = d / "shim.cpp"
    src.write_text(SHIM)
    so = d / "shim.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "pvr.rtl.radiofm_b200", "csrc"), str(src), "-o", str(so), "-lm"])
    L = C.CDLL(str(so))
    for f in (L.cmp_atan2f, L.cmp_sincos, L.cmp_fmod, L.cmp_wraps, L.cmp_sincos_generic, L.cmp_atan2f_generic,
              L.cmp_bit_conversions):
        f.restype = C.c_ulonglong
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def test_atan2f_bit_exact_vs_glibc(shim):
    rng = np.random.default_rng(1)
    n = 4_000_000
    for sy, sx in ((1.0, 1.0), (1e-3, 1.0), (1.0, 1e-4), (1e-20, 1e20), (0.5, -0.5)):
        y = (rng.standard_normal(n) * sy).astype(np.float32)
        x = (rng.standard_normal(n) * sx).astype(np.float32)
        assert shim.cmp_atan2f(_p(y), _p(x), n) == 0
    # every binade of atanf through atan2f(y, 1) and signed zeros / infinities
    bits = np.arange(0, 0x7f800001, 997, dtype=np.uint32)
    y = np.concatenate([bits.view(np.float32), (bits | 0x80000000).view(np.float32)])
    x = np.ones_like(y)
    assert shim.cmp_atan2f(_p(y), _p(x), y.size) == 0
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, 3e38, 1e-39, 2e-20, 5e19], dtype=np.float32)
    yy, xx = [np.ascontiguousarray(v.ravel()) for v in np.meshgrid(sp, sp)]
    assert shim.cmp_atan2f(_p(yy), _p(xx), yy.size) == 0
    assert shim.cmp_atan2f_generic(_p(yy), _p(xx), yy.size) == 0
    y = rng.standard_normal(n).astype(np.float32)
    x = rng.standard_normal(n).astype(np.float32)
    assert shim.cmp_atan2f_generic(_p(y), _p(x), n) == 0


def test_sincos_is_rounded_double_sincos(shim):
    rng = np.random.default_rng(2)
    n = 8_000_000
    p = (rng.random(n) * 6.6 - 0.2).astype(np.float32)          # the PLL phases live in [0, 2 pi]
    assert shim.cmp_sincos(_p(p), n) == 0
    p = (rng.standard_normal(n) * 300.0).astype(np.float32)     # far outside (generic path), still exact
    assert shim.cmp_sincos(_p(p), n) == 0
    p = (rng.random(n) * 32.0 - 16.0).astype(np.float32)
    assert shim.cmp_sincos_generic(_p(p), n) == 0
    # the floats next to k * pi/2 (worst cases of the argument reduction) and the zeros
    k = np.arange(-10, 11, dtype=np.float64) * (np.pi / 2)
    near = np.concatenate([np.nextafter(k.astype(np.float32), np.float32(s)) for s in (-100, 100)] + [k.astype(np.float32),
                          np.array([0.0, -0.0], dtype=np.float32)])
    assert shim.cmp_sincos(_p(np.ascontiguousarray(near)), near.size) == 0


def test_phase_wraps_equal_the_double_expressions(shim):
    """tools/exhaustive_math.cpp enumerates every float of the domain; here a dense random + boundary subset."""
    rng = np.random.default_rng(4)
    n = 4_000_000
    p = (rng.random(n) * 18.4 - 6.0).astype(np.float32)
    assert shim.cmp_wraps(_p(p), n) == 0
    two_pi = np.float32(6.2831855)
    edge = [two_pi]
    for _ in range(64):
        edge.append(np.nextafter(edge[-1], np.float32(100)))
    lo = [np.float32(6.283185)]
    for _ in range(64):
        lo.append(np.nextafter(lo[-1], np.float32(-100)))
    tiny = -np.float32(2.0) ** -np.arange(1, 140, dtype=np.float32)
    e = np.ascontiguousarray(np.concatenate([np.array(edge + lo, dtype=np.float32), tiny, -np.array(edge, dtype=np.float32) + 1]))
    assert shim.cmp_wraps(_p(e), e.size) == 0


def test_fmod_helpers(shim):
    rng = np.random.default_rng(3)
    n = 2_000_000
    p = (rng.standard_normal(n) * 20.0).astype(np.float32)
    assert shim.cmp_fmod(_p(p), n) == 0


def test_osc_gain_float_form_is_exact():
    """rfm_osc_gain_fast (rfm_dsp.cuh): float(1.95 - double(q)) of the NCO_OSC oscillator (DownConvert.cpp:441) from
    five float additions -- restated here in numpy float32 and checked on EVERY float of its domain [0.5, 1.69]."""
    f = np.float32
    lo, hi = f(0.5).view(np.uint32), f(1.69).view(np.uint32)
    q = np.arange(lo, hi + 1, dtype=np.uint32).view(np.float32)
    ref = (np.float64(1.95) - q.astype(np.float64)).astype(np.float32)
    ch = f(1.95)
    cl = f(np.float64(1.95) - np.float64(ch))
    assert float(cl).hex() == "-0x1.99999a0000000p-25"
    s = (ch - q).astype(f)
    e = ((ch - s).astype(f) - q).astype(f)
    r = (s + (e + cl).astype(f)).astype(f)
    assert q.size == 14176749
    assert np.array_equal(r.view(np.uint32), ref.view(np.uint32))


def test_integer_pipe_conversions_are_exact_or_flagged(shim):
    """rfm_f2d_bits / rfm_d2f_bits (float <-> double as bit manipulation, for the XU-free lanes experiment) and
    rfm_sincos_core_b on top of them: bit-identical to the conversion instructions wherever they do not raise the
    replay flag -- random bit patterns, doubles across and beyond the float range, exact ties and near-ties."""
    rng = np.random.default_rng(7)
    n = 4_000_000
    fbits = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    f = (rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & np.uint32(0xbfffffff)).view(np.float32)
    ulp = (np.nextafter(f, np.float32(np.inf)).astype(np.float64) - f.astype(np.float64))
    k = rng.integers(-2, 3, n)
    d = f.astype(np.float64) + 0.5 * ulp + k * ulp * 2.0 ** -29          # around the rounding boundary, ties included
    wide = rng.integers(0, 1 << 63, n // 4, dtype=np.uint64).view(np.float64)
    d[: n // 4] = np.where(np.isfinite(wide), wide, 1.0)
    d = np.ascontiguousarray(np.nan_to_num(d, nan=1.0, posinf=3e38, neginf=-3e38))
    flagged = C.c_ulonglong(0)
    bad = shim.cmp_bit_conversions(fbits.ctypes.data_as(C.POINTER(C.c_uint)), d.ctypes.data_as(C.POINTER(C.c_double)), n,
                                   C.byref(flagged))
    assert bad == 0 and 0 < flagged.value < n

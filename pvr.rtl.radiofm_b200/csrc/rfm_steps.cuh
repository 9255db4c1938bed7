// rfm_steps.cuh -- one-sample update functions of the chain's recurrences (the bodies of the lane kernels).
// __host__ __device__ like rfm_math.cuh, so tests/ can compile them with g++ and check them sample by sample
// against the oracle on the CPU; the product only runs them inside the sm_100a kernels (rfm_kernels.cu).
// Every operation is individually rounded, in the reference's order (no FMA contraction).
#pragma once

#include "rfm_math.cuh"

namespace rfm
{

struct DemodConst
{
  float gain, lo, hi, alpha, beta;
};

struct PilotConstDev
{
  float minfreq, maxfreq, b0, a1, a2, lb0, lb1, minsignal;
  int lock_delay;
};

struct BiquadDev
{
  float A1, A2, B0, B1, B2;
};

// ---- FM-demodulator PLL, cFmDecoder::PhaseLockedLoop (FmDecode.cpp:361-415), loop part ---------------------------
// State (phase, incr); returns nothing: the caller reads st.incr (the NCO frequency after the update), from which
// the DC tracker and the output are formed by demod_output().
struct DemodState
{
  float phase, incr;
};

RFM_HD void demod_step(DemodState& st, float xr, float xi, const DemodConst& k)
{
  float Sin, Cos;
  rfm_sincos(st.phase, &Sin, &Cos);                                   // :386-392
  const float dre = subf(mulf(Cos, xr), mulf(Sin, xi));               // :394 std::complex product
  const float dim = addf(mulf(Cos, xi), mulf(Sin, xr));
  float at;
  if (rfm_atan2f_special(dim, dre))
    at = rfm_atan2f_generic(dim, dre);
  else
    at = rfm_atan2f_main(dim, dre, absf(divf(dim, dre)));
  const float err = negf(at);                                         // :395
  float incr = addf(st.incr, mulf(k.beta, err));                      // :397
  if (incr < k.lo)                                                    // :398-401
    incr = k.lo;
  if (incr > k.hi)
    incr = k.hi;
  st.incr = incr;
  st.phase = rfm_wrap_demod(addf(st.phase, addf(incr, mulf(k.alpha, err)))); // :403-409
}

// Branch-free form of demod_step for the lane kernels.  Identical arithmetic; anything it cannot handle raises the
// sticky `bad` flag and the caller replays the tile with demod_step (see k_bb_lanes).
RFM_HD void demod_step_fast(DemodState& st, float xr, float xi, const DemodConst& k, bool& bad)
{
  float Sin, Cos;
  rfm_sincos_core(st.phase, &Sin, &Cos);                              // phase is always in [0, 2 pi]
  const float dre = subf(mulf(Cos, xr), mulf(Sin, xi));
  const float dim = addf(mulf(Cos, xi), mulf(Sin, xr));
  const float err = negf(rfm_atan2f_fast(dim, dre, bad));
  const float incr = fminf(fmaxf(addf(st.incr, mulf(k.beta, err)), k.lo), k.hi);
  bad = bad | (!(absf(st.phase) < 16.0f));
  st.incr = incr;
  st.phase = rfm_wrap_demod_fast(addf(st.phase, addf(incr, mulf(k.alpha, err))), bad);
}

// DC tracker + output scaling (FmDecode.cpp:410-412): dc is float, the update is evaluated in double.
RFM_HD float demod_output(float incr, float& dc, float gain)
{
  const float pinc = mulf(2.0f, incr);
  dc = d2f(addd(muld(1 - 0.0001, (double)dc), muld(0.0001, (double)pinc)));
  return mulf(subf(pinc, dc), gain);
}

// The same with the three conversions on the integer pipe (rfm_math.cuh: rfm_f2d_bits / rfm_d2f_bits); `bad` -> replay
RFM_HD float demod_output_bits(float incr, float& dc, float gain, bool& bad)
{
  const float pinc = mulf(2.0f, incr);
  const double v = addd(muld(1 - 0.0001, rfm_f2d_bits(dc, bad)), muld(0.0001, rfm_f2d_bits(pinc, bad)));
  // the tracker's value is tiny and may be zero at a stream's start: zero is not "normal" for the narrowing
  bool nb = false;
  const float f = rfm_d2f_bits(v, nb);
  const bool vz = v == 0.0;
  bad = bad | (nb & !vz);
  dc = vz ? u2f(rfm_d_hi(v) & 0x80000000u) : f;
  return mulf(subf(pinc, dc), gain);
}

// ---- 19 kHz pilot PLL, cPilotPhaseLock::Process loop body (FmDecode.cpp:149-216); returns sin(2 phi) -------------
struct PilotState
{
  float phase, freq, i1, i2, q1, q2, x1, level;
};

RFM_HD float pilot_step(PilotState& st, float x, const PilotConstDev& k)
{
  float ps, pc;
  rfm_sincos(st.phase, &ps, &pc);                                     // :167-173
  const float out = mulf(mulf(2.0f, ps), pc);                         // :176
  float pi = mulf(ps, x);                                             // :179-180
  float pq = mulf(pc, x);
  pi = subf(subf(mulf(k.b0, pi), mulf(k.a1, st.i1)), mulf(k.a2, st.i2)); // :185-186
  pq = subf(subf(mulf(k.b0, pq), mulf(k.a1, st.q1)), mulf(k.a2, st.q2));
  st.i2 = st.i1;
  st.i1 = pi;
  st.q2 = st.q1;
  st.q1 = pq;
  float err;                                                          // :190-193
  if (pi > absf(pq))
    err = divf(pq, pi);
  else if (pq > 0.0f)
    err = 1.0f;
  else
    err = -1.0f;
  st.level = (pi < st.level) ? pi : st.level;                         // :196
  float freq = addf(st.freq, addf(mulf(k.lb0, err), mulf(k.lb1, st.x1))); // :199
  st.x1 = err;
  const float t = (freq < k.maxfreq) ? freq : k.maxfreq;              // :203 max(minfreq, min(maxfreq, f))
  freq = (k.minfreq < t) ? t : k.minfreq;
  st.freq = freq;
  st.phase = rfm_wrap_pilot(addf(st.phase, freq));                    // :206-208
  return out;
}

// FAKE_SINCOS: timing experiment only (RFM_DEBUG_FAKE_SINCOS: what would a shorter pilot chain buy?) -- the hardware's
// approximate sincos, results are NOT the reference's
template <bool FAKE_SINCOS = false, bool XUFREE = false>
RFM_HD float pilot_step_fast(PilotState& st, float x, const PilotConstDev& k, const SinCosRegs& sca, bool& bad)
{
  float ps, pc;
  bad = bad | (!(absf(st.phase) < 16.0f));
#if defined(__CUDA_ARCH__)
  if (FAKE_SINCOS)
    __sincosf(st.phase, &ps, &pc);
  else
#endif
  if (XUFREE)
  {
    bad = bad | (st.phase == 0.0f);        // sin(+-0): the narrowing has no zero; the replay handles it
    rfm_sincos_core_b(st.phase, sca, &ps, &pc, bad);
  }
  else
    rfm_sincos_core_a(st.phase, sca, &ps, &pc);
  const float out = mulf(mulf(2.0f, ps), pc);
  float pi = mulf(ps, x);
  float pq = mulf(pc, x);
  pi = subf(subf(mulf(k.b0, pi), mulf(k.a1, st.i1)), mulf(k.a2, st.i2));
  pq = subf(subf(mulf(k.b0, pq), mulf(k.a1, st.q1)), mulf(k.a2, st.q2));
  st.i2 = st.i1;
  st.i1 = pi;
  st.q2 = st.q1;
  st.q1 = pq;
  const bool use_div = pi > absf(pq);
  bad = bad | (use_div & rfm_div_unsafe_below(pq, pi));
  const float ediv = rfm_div_fast(pq, pi);
  const float esat = (pq > 0.0f) ? 1.0f : -1.0f;
  const float err = use_div ? ediv : esat;
  st.level = fminf(pi, st.level);
  const float freq = fmaxf(fminf(addf(st.freq, addf(mulf(k.lb0, err), mulf(k.lb1, st.x1))), k.maxfreq), k.minfreq);
  st.x1 = err;
  st.freq = freq;
  st.phase = rfm_wrap_pilot_fast(addf(st.phase, freq), bad);
  return out;
}

// The same step with the sincos taken off the dependent chain (rfm_sincos_predict / rfm_sincos_correct, rfm_math.cuh):
// (ps, pc) = sincos(st.phase) is carried from the previous step; this step predicts the next phase from the OLD
// frequency, evaluates the double kernels there while the chain runs, and corrects by the (tiny) difference once the
// new phase is known.  `bad` (sticky) means "not guaranteed identical to pilot_step": the caller replays the tile.
struct PilotCarry
{
  float ps, pc;
};

RFM_HD float pilot_step_spec(PilotState& st, PilotCarry& sc, float x, const PilotConstDev& k, const SinCosRegs& sca, bool& bad)
{
  const float ph_hat = rfm_wrap_pilot_fast(addf(st.phase, st.freq), bad);
  const SinCosPred pred = rfm_sincos_predict(ph_hat, sca);
  const float ps = sc.ps, pc = sc.pc;
  const float out = mulf(mulf(2.0f, ps), pc);
  float pi = mulf(ps, x);
  float pq = mulf(pc, x);
  pi = subf(subf(mulf(k.b0, pi), mulf(k.a1, st.i1)), mulf(k.a2, st.i2));
  pq = subf(subf(mulf(k.b0, pq), mulf(k.a1, st.q1)), mulf(k.a2, st.q2));
  st.i2 = st.i1;
  st.i1 = pi;
  st.q2 = st.q1;
  st.q1 = pq;
  const bool use_div = pi > absf(pq);
  bad = bad | (use_div & rfm_div_unsafe_below(pq, pi));
  const float ediv = rfm_div_fast(pq, pi);
  const float esat = (pq > 0.0f) ? 1.0f : -1.0f;
  const float err = use_div ? ediv : esat;
  st.level = fminf(pi, st.level);
  const float freq = fmaxf(fminf(addf(st.freq, addf(mulf(k.lb0, err), mulf(k.lb1, st.x1))), k.maxfreq), k.minfreq);
  st.x1 = err;
  st.freq = freq;
  st.phase = rfm_wrap_pilot_fast(addf(st.phase, freq), bad);
  bad = bad | (st.phase == 0.0f);
  rfm_sincos_correct(pred, st.phase, ph_hat, &sc.ps, &sc.pc, bad);
  return out;
}

// ---- cIirFilter DF-II biquad (IirFilter.cpp:78-105) ---------------------------------------------------------------
RFM_HD float biquad_step(const BiquadDev& c, float x, float& w1, float& w2)
{
  const float w0 = subf(subf(x, mulf(c.A1, w1)), mulf(c.A2, w2));
  const float y = addf(addf(mulf(c.B0, w0), mulf(c.B1, w1)), mulf(c.B2, w2));
  w2 = w1;
  w1 = w0;
  return y;
}

// ---- RDS Costas loop phase detector: the reference's rational arctan2 (RDSProcess.cpp:187-217) --------------------
RFM_HD float arctan2_approx(float y, float x)
{
  if (x == 0.0f)
  {
    if (y > 0.0f)
      return d2f(RFM_K_PI2);
    if (y == 0.0f)
      return 0.0f;
    return d2f(-RFM_K_PI2);
  }
  float angle;
  const float z = divf(y, x);
  if (absf(z) < 1.0f)
  {
    angle = d2f(divd((double)z, addd(1.0, muld(muld(0.2854, (double)z), (double)z))));
    if (x < 0.0f)
    {
      if (y < 0.0f)
        return d2f(subd((double)angle, RFM_K_PI));
      return d2f(addd((double)angle, RFM_K_PI));
    }
  }
  else
  {
    angle = d2f(subd(RFM_K_PI2, divd((double)z, addd((double)mulf(z, z), 0.2854))));
    if (y < 0.0f)
      return d2f(subd((double)angle, RFM_K_PI));
  }
  return angle;
}

// one sample of cRDSRxSignalProcessor::ProcessRdsPll (RDSProcess.cpp:222-268); returns Im of the rotated sample
RFM_HD float rds_pll_step(float& phase, float& freq, float xr, float xi, float lo, float hi, float alpha, float beta)
{
  float Sin, Cos;
  rfm_sincos(phase, &Sin, &Cos);
  const float tre = subf(mulf(Cos, xr), mulf(Sin, xi));
  const float tim = addf(mulf(Cos, xi), mulf(Sin, xr));
  const float err = negf(arctan2_approx(tim, tre));
  freq = addf(freq, mulf(beta, err));
  if (freq > hi)
    freq = hi;
  else if (freq < lo)
    freq = lo;
  phase = addf(phase, addf(freq, mulf(alpha, err)));
  return tim;
}

} // namespace rfm

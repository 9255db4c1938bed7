// rfm_plan.cpp -- the reference's CONSTRUCTORS restated on the host: filter designs, PLL constants, the decimation-chain
// planner (interface: rfm_plan.h).  Host only, run once per decoder.
//
// Attribution.  The kernels of this library consume coefficient tables and constants that must be bit-identical to
// what the reference computes (one different ulp in a tap moves every output), so this file -- unlike the rest of the
// library -- follows the reference's initialisation code expression by expression, in its precision and operation
// order.  It is a derived work of pvr.rtl.radiofm (GPL-2.0-or-later):
//     Izero / PlanKaiserLP / PlanKaiserHP   <- src/FirFilter.cpp:39-58,78-148,195-264
//                                              Copyright (C) 2010-2013 Moe Wheatley, (C) 2015-2020 Alwin Esch (Team KODI)
//     PlanBiquad                            <- src/IirFilter.cpp:11-51          (same authors)
//     PlanLanczos, PlanDecimationChain      <- src/DownConvert.cpp:18-56,327-371,378-399
//                                              Copyright (C) 2010-2013 Moe Wheatley, (C) 2013 Joris van Rantwijk,
//                                              (C) 2015-2020 Alwin Esch (Team KODI)
//     PlanPilot, PlanDecoder                <- src/FmDecode.cpp:45-64,88-139,237-324
//                                              Copyright (C) 2013 Joris van Rantwijk, (C) 2015-2020 Alwin Esch (Team KODI)
//     RDS process constants                 <- src/RDSProcess.cpp:43-118         (Moe Wheatley / Alwin Esch)
//     halfband_taps.inc                     <- src/filtercoef.h (tap DATA; Moe Wheatley, Simplified-BSD in CuteSDR)
// and is distributed under the same licence terms (SPDX-License-Identifier: GPL-2.0-or-later).
//
// Build with plain g++ -O2 -ffp-contract=off (no -march / -ffast-math): the tables must come out bit-identical to what
// the reference's constructors compute with the same libm -- tests/test_abi_host.py and tests/test_oracle_port.py
// check every constant and table against the compiled reference and the golden fixtures.
#include "rfm_plan.h"

#include <math.h>

namespace rfm
{

#define K_2PI (2.0 * 3.14159265358979323846)
#define K_PI (3.14159265358979323846)

#include "halfband_taps.inc"

bool PlanBiquad(int type, float F0, float Q, float Fs, Biquad* o)
{
  const float w0 = (float)(K_2PI * F0 / Fs);
  const float alpha = (float)(sinf(w0) / (2.0 * Q));
  const float A = (float)(1.0 / (1.0 + alpha));
  switch (type)
  {
    case 0:
      o->B0 = (float)(A * ((1.0 - cosf(w0)) / 2.0));
      o->B1 = (float)(A * (1.0 - cosf(w0)));
      o->B2 = (float)(A * ((1.0 - cosf(w0)) / 2.0));
      break;
    case 1:
      o->B0 = (float)(A * ((1.0 + cosf(w0)) / 2.0));
      o->B1 = (float)(-A * (1.0 + cosf(w0)));
      o->B2 = (float)(A * ((1.0 + cosf(w0)) / 2.0));
      break;
    case 2:
      o->B0 = A * alpha;
      o->B1 = 0.0f;
      o->B2 = A * -alpha;
      break;
    case 3:
      o->B0 = (float)(A * 1.0);
      o->B1 = (float)(A * (-2.0 * cosf(w0)));
      o->B2 = (float)(A * 1.0);
      break;
    default:
      return false;
  }
  o->A1 = (float)(A * (-2.0 * cosf(w0)));
  o->A2 = (float)(A * (1.0 - alpha));
  return true;
}

std::vector<float> PlanLanczos(unsigned order, double cutoff)
{
  // The ctor passes (order - 1) as filter_order; coefficients live at [1 .. order], [0] and
  // [order + 1] stay zero so that (order + 1) interpolated taps can always be formed.
  const unsigned fo = order - 1;
  std::vector<float> c(order + 2, 0.0f);
  double ysum = 0.0;
  for (int i = 1; i <= (int)fo + 1; i++)
  {
    const int t2 = 2 * i - (int)fo;
    double y;
    if (t2 == 0)
    {
      y = 1.0;
    }
    else
    {
      const double x1 = cutoff * t2;
      const double x2 = t2 / double(fo + 2);
      y = (sinf((float)(K_PI * x1)) / K_PI / x1) * (sinf((float)(K_PI * x2)) / K_PI / x2);
    }
    c[i] = (float)y;
    ysum += y;
  }
  for (unsigned i = 1; i <= fo + 1; i++)
    c[i] = (float)(c[i] / ysum);
  return c;
}

static float Izero(float x)
{
  const float x2 = x / 2.0f;
  float sum = 1.0f, ds = 1.0f, di = 1.0f, tmp;
  const float errorlimit = (float)1e-9;
  do
  {
    tmp = x2 / di;
    tmp *= tmp;
    ds *= tmp;
    sum += ds;
    di = (float)(di + 1.0);
  } while (ds >= errorlimit * sum);
  return sum;
}

std::vector<float> PlanKaiserLP(unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs)
{
  float Beta;
  const float normFpass = Fpass / Fs;
  const float normFstop = Fstop / Fs;
  const float normFcut = (normFstop + normFpass) / 2.0f;
  if (Astop < 20.96f)
    Beta = 0;
  else if (Astop >= 50.0f)
    Beta = (float)(.1102 * (Astop - 8.71f));
  else
    Beta = (float)(.5842 * powf((Astop - 20.96f), (float)0.4) + .07886f * (Astop - 20.96f));
  unsigned taps = (unsigned)((Astop - 8.0f) / (2.285f * K_2PI * (normFstop - normFpass)) + 1);
  if (taps > kMaxFirTaps)
    taps = kMaxFirTaps;
  if (taps < 3)
    taps = 3;
  if (NumTaps)
    taps = NumTaps;
  std::vector<float> coef(taps);
  const float fCenter = (float)(.5 * (float)(taps - 1));
  const float izb = Izero(Beta);
  for (unsigned n = 0; n < taps; ++n)
  {
    float x = (float)n - fCenter;
    float c;
    if ((float)n == fCenter)
      c = (float)(2.0 * normFcut);
    else
      c = (float)(sinf((float)(K_2PI * x * normFcut)) / (K_PI * x));
    x = ((float)n - ((float)taps - 1.0f) / 2.0f) / (((float)taps - 1.0f) / 2.0f);
    coef[n] = Scale * c * Izero(Beta * sqrtf(1 - (x * x))) / izb;
  }
  return coef;
}

// cFirFilter::InitHPFilter, FirFilter.cpp:195-264 (RealType = float; the double sub-expressions are the reference's:
// its literals 2.0 / 1.0 and K_PI are doubles, Beta's are floats here -- unlike InitLPFilter)
std::vector<float> PlanKaiserHP(unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs)
{
  float Beta;
  const float normFpass = Fpass / Fs;
  const float normFstop = Fstop / Fs;
  const float normFcut = (float)((normFstop + normFpass) / 2.0);
  if (Astop < 20.96f)
    Beta = 0;
  else if (Astop >= 50.0f)
    Beta = .1102f * (Astop - 8.71f);
  else
    Beta = .5842f * powf((Astop - 20.96f), 0.4f) + .07886f * (Astop - 20.96f);
  unsigned taps = (unsigned)((Astop - 8.0f) / (2.285f * K_2PI * (normFpass - normFstop)) + 1);
  if (taps > kMaxFirTaps - 1)
    taps = kMaxFirTaps - 1;
  if (taps < 3)
    taps = 3;
  taps |= 1;
  if (NumTaps)
    taps = NumTaps;
  std::vector<float> coef(taps);
  const float izb = Izero(Beta);
  const float fCenter = .5f * (float)(taps - 1);
  for (unsigned n = 0; n < taps; n++)
  {
    float x = (float)((float)n - (float)(taps - 1) / 2.0);
    float c;
    if ((float)n == fCenter)
      c = (float)(1.0 - 2.0 * normFcut);
    else
      c = (float)(sinf((float)(K_PI * x)) / (K_PI * x) - sinf((float)(K_2PI * x * normFcut)) / (K_PI * x));
    x = ((float)n - ((float)taps - 1.0f) / 2.0f) / (((float)taps - 1.0f) / 2.0f);
    coef[n] = Scale * c * Izero(Beta * sqrtf(1 - (x * x))) / izb;
  }
  return coef;
}

float PlanDecimationChain(float InRate, float MaxBW, bool wfm, std::vector<HalfBandStage>* st)
{
  st->clear();
  float f = InRate;
  if (wfm)
  {
    while (f > 400000.0)
    {
      st->push_back({51, false, HB51TAP_H});
      f = (float)(f / 2.0);
    }
    return f;
  }
  while ((f > (MaxBW / HB51TAP_MAX)) && (f > (7900.0 * 2.0)))
  {
    if (f >= (MaxBW / CIC3_MAX))
      st->push_back({3, false, nullptr});
    else if (f >= (MaxBW / HB11TAP_MAX))
      st->push_back({11, true, HB11TAP_H});
    else if (f >= (MaxBW / HB15TAP_MAX))
      st->push_back({15, false, HB15TAP_H});
    else if (f >= (MaxBW / HB19TAP_MAX))
      st->push_back({19, false, HB19TAP_H});
    else if (f >= (MaxBW / HB23TAP_MAX))
      st->push_back({23, false, HB23TAP_H});
    else if (f >= (MaxBW / HB27TAP_MAX))
      st->push_back({27, false, HB27TAP_H});
    else if (f >= (MaxBW / HB31TAP_MAX))
      st->push_back({31, false, HB31TAP_H});
    else if (f >= (MaxBW / HB35TAP_MAX))
      st->push_back({35, false, HB35TAP_H});
    else if (f >= (MaxBW / HB39TAP_MAX))
      st->push_back({39, false, HB39TAP_H});
    else if (f >= (MaxBW / HB43TAP_MAX))
      st->push_back({43, false, HB43TAP_H});
    else if (f >= (MaxBW / HB47TAP_MAX))
      st->push_back({47, false, HB47TAP_H});
    else if (f >= (MaxBW / HB51TAP_MAX))
      st->push_back({51, false, HB51TAP_H});
    f = (float)(f / 2.0);
  }
  return f;
}

NcoOsc PlanNcoOsc(float nco_freq, float in_rate)
{
  NcoOsc o;
  o.inc = (float)(K_2PI * nco_freq / in_rate);
  o.cosv = cosf(o.inc);
  o.sinv = sinf(o.inc);
  return o;
}

PilotConst PlanPilot(float freq, float bandwidth, float minsignal)
{
  PilotConst p;
  p.minfreq = (float)((freq - bandwidth) * K_2PI);
  p.maxfreq = (float)((freq + bandwidth) * K_2PI);
  p.minsignal = minsignal;
  p.lock_delay = int(20.0f / bandwidth);
  const float p1 = (float)exp(-1.146f * bandwidth * K_2PI);
  const float p2 = (float)exp(-5.331f * bandwidth * K_2PI);
  p.a1 = -p1 - p2;
  p.a2 = p1 * p2;
  p.b0 = 1 + p.a1 + p.a2;
  p.lb0 = (float)(0.62f * bandwidth * K_2PI);
  p.lb1 = (float)(-p.lb0 * exp(-0.1153 * bandwidth * K_2PI));
  p.freq0 = (float)(freq * K_2PI);
  return p;
}

DecoderPlan PlanDecoder(double fs_if, double tuning_offset, double fs_pcm, double bw_pcm, unsigned downsample,
                        bool usver)
{
  DecoderPlan p;
  p.fs_if = (float)fs_if;
  p.fs_bb = (float)(fs_if / downsample);
  p.freq_dev = (float)60000.0;
  p.downsample = downsample;
  p.tuning_shift = (int)lrint(-64.0 * tuning_offset / fs_if);
  p.demod_gain = (float)(1.0 / (60000.0 / p.fs_bb * K_2PI));

  for (int u = 0; u < 256; ++u)
    p.u8lut[u] = (float)(u / (255.0 / 2.0) - 1.0);

  {
    const float phase_step = (float)(K_2PI / float(kTunerTable));
    for (unsigned i = 0; i < kTunerTable; ++i)
    {
      const int64_t r = ((int64_t)p.tuning_shift * (int64_t)i) % (int64_t)kTunerTable;
      const float phi = (float)r * phase_step;
      p.tuner[2 * i] = cosf(phi) * 2.0f;
      p.tuner[2 * i + 1] = sinf(phi) * 2.0f;
    }
  }

  p.in_order = 8 * downsample;
  p.in_coeff = PlanLanczos(p.in_order, 0.6 / downsample);
  p.pilot = PlanPilot((float)(19000.0 / p.fs_bb), 50 / p.fs_bb, 0.04f);
  p.a_order = (unsigned)int(p.fs_bb / 1000.0);
  p.a_coeff = PlanLanczos(p.a_order, bw_pcm / p.fs_bb);
  p.a_ratio = p.fs_bb / fs_pcm;
  p.a_pstep = (float)p.a_ratio;

  PlanBiquad(3, (float)19000.0, 5, (float)fs_pcm, &p.notch);
  p.lp_coef = PlanKaiserLP(0, 1.0f, 60.0f, 15000.0f, (float)(1.4 * 15000.0), (float)fs_pcm);
  {
    const float Time = usver ? (float)75E-6 : (float)50E-6, SampleRate = (float)fs_pcm;
    p.de_alpha = (1.0f - expf(-1.0f / (SampleRate * Time)));
  }
  {
    const float fac = (float)(K_2PI / p.fs_bb);
    const float bandwidth = 0.85f * p.fs_bb;
    const float maxFreqDev = 0.95f * (0.5f * p.fs_bb);
    p.nco_lo = (-maxFreqDev) * fac;
    p.nco_hi = (+maxFreqDev) * fac;
    p.pll_alpha = 0.125f * bandwidth * fac;
    p.pll_beta = (p.pll_alpha * p.pll_alpha) / 2.0f;
  }

  // cRDSRxSignalProcessor
  p.rds_rate = PlanDecimationChain(p.fs_bb, 8000.0f, false, &p.rds_stages);
  p.rds_osc = PlanNcoOsc(-57000.0f, p.fs_bb);
  {
    const float norm = (float)(K_2PI / p.rds_rate);
    const float ncofreq = 0.0f;
    p.rpll_lo = (float)((ncofreq - 12.0) * norm);
    p.rpll_hi = (float)((ncofreq + 12.0) * norm);
    p.rpll_alpha = (float)(2.0 * 0.707 * 1.00 * norm);
    p.rpll_beta = (float)((p.rpll_alpha * p.rpll_alpha) / (4.0 * 0.707 * 0.707));
  }
  {
    const double bitrate = 57000.0 / 48.0;
    const unsigned L = (unsigned)(p.rds_rate / bitrate);
    std::vector<float> c(2 * L + 1, 0.0f);
    for (unsigned i = 0; i <= L; i++)
    {
      const float t = (float)i / p.rds_rate;
      const float x = (float)(t * bitrate);
      const float x64 = (float)(64.0 * x);
      c[i + L] = (float)(.75 * cosf((float)(2.0 * K_2PI * x)) * ((1.0 / (1.0 / x - x64)) - (1.0 / (9.0 / x - x64))));
      c[L - i] = (float)(-.75 * cosf((float)(2.0 * K_2PI * x)) * ((1.0 / (1.0 / x - x64)) - (1.0 / (9.0 / x - x64))));
    }
    unsigned taps = 2 * L;
    if (taps > kMaxFirTaps)
      taps = kMaxFirTaps; // InitConstFir clamp, FirFilter.cpp:306-309
    p.mf_coef.assign(c.begin(), c.begin() + taps);
  }
  p.rlp_coef = PlanKaiserLP(0, 1.0f, 40.0f, 2400.0f, (float)(1.3 * 2400.0), p.rds_rate);
  PlanBiquad(2, (float)(57000.0 / 48.0), 500, p.rds_rate, &p.rsync);
  return p;
}

unsigned FractionalOutputs(float pos_frac, float pstep, unsigned n, float* pos_frac_out)
{
  const float p = pos_frac;
  float pf = p;
  unsigned pi = (unsigned)int(pf);
  unsigned i = 0;
  while (pi < n)
  {
    i++;
    pf = p + i * pstep;
    pi = (unsigned)int(pf);
  }
  float np = pf - n;
  if (np < 0)
    np = 0;
  if (pos_frac_out)
    *pos_frac_out = np;
  return i;
}

} // namespace rfm

// rfm_plan.h -- host-side planner: every constant and coefficient table of the chain, evaluated with
// the same expressions and the same float/double precision as the reference constructors
// (SURVEY.md Appendix C).  Host-only (libm); the result is uploaded once per decoder.
#pragma once

#include <stdint.h>

#include <vector>

namespace rfm
{

constexpr unsigned kMaxFirTaps = 75;  // cFirFilter MAX_NUMCOEF, FirFilter.h:15
constexpr unsigned kMaxDecStages = 9; // MAX_DECSTAGES - 1, DownConvert.h:63
constexpr unsigned kTunerTable = 64;  // m_TuningTableSize, FmDecode.cpp:249

struct Biquad
{
  float A1, A2, B0, B1, B2; // IirFilter.h:25-29
};

// cIirFilter::Init, IirFilter.cpp:11-60.  type: 0 LP, 1 HP, 2 BP, 3 BR.  Returns false for others.
bool PlanBiquad(int type, float F0, float Q, float Fs, Biquad* out);

// MakeLanczosCoeff as called by cDownsampleFilter's ctor (DownConvert.cpp:18-56,78): order+2 entries.
std::vector<float> PlanLanczos(unsigned order, double cutoff);

// cFirFilter::InitLPFilter, FirFilter.cpp:78-148.  Returns the taps.
std::vector<float> PlanKaiserLP(unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs);
std::vector<float> PlanKaiserHP(unsigned NumTaps, float Scale, float Astop, float Fpass, float Fstop, float Fs); // FirFilter.cpp:195-264

struct HalfBandStage
{
  int len;        // 3 = CIC3, 11 with fixed11 = unrolled 11-tap, else generic half-band length
  bool fixed11;
  const float* h; // taps (nullptr for CIC3)
};

// CRDSDownConvert::SetDataRate / SetWfmDataRate planners, DownConvert.cpp:327-399.
float PlanDecimationChain(float InRate, float MaxBW, bool wfm, std::vector<HalfBandStage>* stages);

struct NcoOsc
{
  float inc, cosv, sinv; // CRDSDownConvert::SetFrequency, DownConvert.cpp:311-320
};
NcoOsc PlanNcoOsc(float nco_freq, float in_rate);

struct PilotConst // cPilotPhaseLock ctor, FmDecode.cpp:88-140
{
  float minfreq, maxfreq, b0, a1, a2, lb0, lb1, freq0, minsignal;
  int lock_delay;
};
PilotConst PlanPilot(float freq, float bandwidth, float minsignal);

struct DecoderPlan // cFmDecoder ctor, FmDecode.cpp:237-314 (+ cRDSRxSignalProcessor ctor, RDSProcess.cpp:43-88)
{
  float fs_if, fs_bb, freq_dev;
  unsigned downsample;
  int tuning_shift;
  float demod_gain, nco_lo, nco_hi, pll_alpha, pll_beta;
  float u8lut[256];                 // RTL_SDR_Source.cpp:207-211
  float tuner[2 * kTunerTable];     // cFineTuner table (re, im), FmDecode.cpp:45-59
  unsigned in_order;                // 8 * downsample
  std::vector<float> in_coeff;      // in_order + 2
  PilotConst pilot;
  unsigned a_order;                 // int(Fb / 1000)
  std::vector<float> a_coeff;       // a_order + 2
  double a_ratio;                   // Fb / fs_pcm
  float a_pstep;                    // float(a_ratio), DownConvert.cpp:203
  std::vector<float> lp_coef;       // audio 29-tap Kaiser LP
  float de_alpha;
  Biquad notch;
  // RDS
  float rds_rate;
  std::vector<HalfBandStage> rds_stages;
  NcoOsc rds_osc;
  std::vector<float> rlp_coef;      // 2.4 kHz Kaiser LP at rds_rate
  float rpll_lo, rpll_hi, rpll_alpha, rpll_beta;
  std::vector<float> mf_coef;       // biphase matched filter
  Biquad rsync;
};

DecoderPlan PlanDecoder(double fs_if, double tuning_offset, double fs_pcm, double bw_pcm, unsigned downsample,
                        bool usver);

// Number of outputs and the carried position of one cDownsampleFilter fractional-branch call
// (DownConvert.cpp:203-232) -- data independent, so the host tracks it for all streams at once.
unsigned FractionalOutputs(float pos_frac, float pstep, unsigned n, float* pos_frac_out);

} // namespace rfm

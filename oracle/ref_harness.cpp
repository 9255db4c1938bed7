// oracle/ref_harness.cpp -- C-ABI harness around the UNMODIFIED reference DSP sources.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (pvr.rtl.radiofm_b200/, include/) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use the library built from it.
//
// What it is: a thin extern "C" wrapper that drives the reference's own classes
// (cFmDecoder, cFineTuner, cPilotPhaseLock, cDownsampleFilter, CRDSDownConvert, cFirFilter,
// cIirFilter, cFreqShift, cRDSRxSignalProcessor) compiled IN PLACE from /root/reference/src
// by oracle/Makefile into oracle/_ref/libradiofm_ref.so.  No reference source is copied into
// this repository.  The only reference code that is *restated* here is
//   * the one-line u8 -> complex<float> conversion of RTL_SDR_Source.cpp:207-211 (that file
//     needs librtlsdr and cannot be compiled), and
//   * the call ORDER of cFmDecoder::ProcessStream (FmDecode.cpp:417-502) and
//     cRDSRxSignalProcessor::Process (RDSProcess.cpp:120-180) in ref_fm_process_staged(),
//     which calls the reference's own member functions one by one so that every intermediate
//     buffer can be tapped.  tests/test_oracle_ref.py checks staged == unstaged bit for bit.
// cRDSGroupDecoder (RDSGroupDecoder.cpp, out of scope: SURVEY.md section 2 row 8) is replaced by
// a recording stand-in that stores the 4 x u16 block words of each decoded group.
//
// Build flags must stay plain "-O2 -std=c++17" (no -march, no -ffast-math, no FMA contraction):
// SURVEY.md section 0.5 / Appendix A.

#include <assert.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <complex>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

// The private helpers of cFmDecoder are declared `inline` (FmDecode.h:170-176) and have no
// external symbol, so FmDecode.cpp is unity-included here; private -> public gives access to
// the intermediate buffers and filter objects.  All std headers are included above first.
#define private public
#define protected public
#include "FmDecode.cpp"
#undef private
#undef protected

// ---------------------------------------------------------------------------------------------
// Recording stand-in for cRDSGroupDecoder (declared in RDSGroupDecoder.h:24-31).
// ---------------------------------------------------------------------------------------------
namespace
{
std::mutex g_sinkMutex;
std::map<const cRDSGroupDecoder*, std::vector<uint16_t>*> g_sinks;

void RegisterSink(const cRDSGroupDecoder* dec, std::vector<uint16_t>* sink)
{
  std::lock_guard<std::mutex> lock(g_sinkMutex);
  g_sinks[dec] = sink;
}

void UnregisterSink(const cRDSGroupDecoder* dec)
{
  std::lock_guard<std::mutex> lock(g_sinkMutex);
  g_sinks.erase(dec);
}
} // namespace

cRDSGroupDecoder::cRDSGroupDecoder(cRadioReceiver* proc) : m_RadioProc(proc)
{
}

void cRDSGroupDecoder::Reset()
{
}

void cRDSGroupDecoder::DecodeRDS(uint16_t* blockData)
{
  std::vector<uint16_t>* sink = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_sinkMutex);
    auto it = g_sinks.find(this);
    if (it != g_sinks.end())
      sink = it->second;
  }
  if (sink)
    sink->insert(sink->end(), blockData, blockData + 4);
}

// ---------------------------------------------------------------------------------------------
// Handles
// ---------------------------------------------------------------------------------------------
struct RefFm
{
  cFmDecoder* dec = nullptr;
  std::vector<uint16_t> groups; // 4 words per group
  std::vector<uint8_t> bits; // data bits (after differential decode), staged path only
  std::vector<ComplexType> iq; // conversion scratch
};

struct RefRds
{
  cRDSRxSignalProcessor* rds = nullptr;
  std::vector<uint16_t> groups;
};

// Caller-provided tap buffers for ref_fm_process_staged (any pointer may be NULL).
struct RefTaps
{
  float* tuned; // cf32 [n]            after cFineTuner
  float* demod_in; // cf32 [nb]        after m_ReSampleInput
  float* baseband; // f32 [nb]         after PhaseLockedLoop
  float* rds_dec; // cf32 [nr]         after CRDSDownConvert::ProcessData
  float* rds_lp; // cf32 [nr]          after m_RdsLPFilter
  float* rds_pll; // f32 [nr]          after ProcessRdsPll
  float* rds_mf; // f32 [nr]           after m_RdsMatchedFilter
  float* rds_sync; // f32 [nr]         after m_RdsBitSyncFilter
  float* mono_rs; // f32 [na]          after m_ReSampleMono
  float* pilot38; // f32 [nb]          cPilotPhaseLock output (sin 2phi)
  float* rawstereo; // f32 [nb]        pilot38 * 2 * baseband
  float* stereo_rs; // f32 [na]        after m_ReSampleStereo
  float* lp; // f32 [2*na]             stereo then mono, after m_LPFilter.ProcessTwo
  float* deemph; // f32 [2*na]         after ProcessDeemphasisFilter
  float* notch; // f32 [2*na]          after m_NotchFilter.ProcessTwo
  uint32_t nb; // out: baseband samples
  uint32_t nr; // out: RDS-rate samples
  uint32_t na; // out: audio frames
  uint32_t stereo; // out: m_StereoDetected for this block
};

static void CopyOut(float* dst, const void* src, size_t nfloats)
{
  if (dst)
    memcpy(dst, src, nfloats * sizeof(float));
}

extern "C"
{

// --- conversion: RTL_SDR_Source.cpp:207-211 -------------------------------------------------
__attribute__((visibility("default"))) void ref_u8_to_cf32(const uint8_t* buf,
                                                           unsigned n,
                                                           float* out)
{
  ComplexType* o = reinterpret_cast<ComplexType*>(out);
  for (unsigned i = 0; i < n; ++i)
    o[i] = ComplexType((buf[2 * i] / (255.0 / 2.0) - 1.0), (buf[2 * i + 1] / (255.0 / 2.0) - 1.0));
}

// --- cFmDecoder -------------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_fm_create(double fs_if,
                                                           double tuning_offset,
                                                           double fs_pcm,
                                                           double bw_pcm,
                                                           unsigned downsample,
                                                           int usver)
{
  RefFm* h = new RefFm;
  h->dec = new cFmDecoder(nullptr, fs_if, tuning_offset, fs_pcm, bw_pcm, downsample, usver != 0);
  // never initialised by the reference (RDSProcess.h:99,104); SURVEY.md Appendix A
  h->dec->m_RDSProcess.m_InBitStream = 0;
  h->dec->m_RDSProcess.m_BlockErrors = 0;
  RegisterSink(&h->dec->m_RDSProcess.m_Decoder, &h->groups);
  return h;
}

__attribute__((visibility("default"))) void ref_fm_destroy(void* hv)
{
  RefFm* h = static_cast<RefFm*>(hv);
  if (!h)
    return;
  UnregisterSink(&h->dec->m_RDSProcess.m_Decoder);
  delete h->dec;
  delete h;
}

__attribute__((visibility("default"))) void ref_fm_reset(void* hv)
{
  static_cast<RefFm*>(hv)->dec->Reset();
}

__attribute__((visibility("default"))) unsigned ref_fm_process_cf32(void* hv,
                                                                    const float* iq,
                                                                    unsigned n,
                                                                    float* audio)
{
  RefFm* h = static_cast<RefFm*>(hv);
  return h->dec->ProcessStream(reinterpret_cast<const ComplexType*>(iq), n, audio);
}

__attribute__((visibility("default"))) unsigned ref_fm_process_u8(void* hv,
                                                                  const uint8_t* iq,
                                                                  unsigned n,
                                                                  float* audio)
{
  RefFm* h = static_cast<RefFm*>(hv);
  h->iq.resize(n);
  ref_u8_to_cf32(iq, n, reinterpret_cast<float*>(h->iq.data()));
  return h->dec->ProcessStream(h->iq.data(), n, audio);
}

__attribute__((visibility("default"))) unsigned ref_fm_take_groups(void* hv,
                                                                   uint16_t* out,
                                                                   unsigned max_groups)
{
  RefFm* h = static_cast<RefFm*>(hv);
  unsigned n = std::min<size_t>(h->groups.size() / 4, max_groups);
  if (out && n)
    memcpy(out, h->groups.data(), n * 4 * sizeof(uint16_t));
  h->groups.erase(h->groups.begin(), h->groups.begin() + n * 4);
  return n;
}

__attribute__((visibility("default"))) unsigned ref_fm_take_bits(void* hv,
                                                                 uint8_t* out,
                                                                 unsigned max_bits)
{
  RefFm* h = static_cast<RefFm*>(hv);
  unsigned n = std::min<size_t>(h->bits.size(), max_bits);
  if (out && n)
    memcpy(out, h->bits.data(), n);
  h->bits.erase(h->bits.begin(), h->bits.begin() + n);
  return n;
}

// out[0]=stereo out[1]=IF level out[2]=baseband level out[3]=baseband mean
// out[4]=pilot level out[5]=tuning offset (FmDecode.h:140-165)
__attribute__((visibility("default"))) void ref_fm_status(void* hv, float* out)
{
  const cFmDecoder* d = static_cast<RefFm*>(hv)->dec;
  out[0] = d->StereoDetected() ? 1.0f : 0.0f;
  out[1] = d->GetInterfaceLevel();
  out[2] = d->GetBasebandLevel();
  out[3] = d->m_BasebandMean;
  out[4] = d->GetPilotLevel();
  out[5] = d->GetTuningOffset();
}

// Derived constants and tables of a constructed decoder, for KAT checks of the product's
// host-side planner.  Scalars go to `s` (see index list below); tables are copied on request.
__attribute__((visibility("default"))) void ref_fm_constants(void* hv, double* s)
{
  const cFmDecoder* d = static_cast<RefFm*>(hv)->dec;
  const cRDSRxSignalProcessor& r = d->m_RDSProcess;
  const cPilotPhaseLock& p = d->m_PilotPLL;
  int i = 0;
  s[i++] = d->m_SampleRate_Interface; // 0
  s[i++] = d->m_SampleRate_Baseband; // 1
  s[i++] = d->m_TuningShift; // 2
  s[i++] = d->m_FMDeModGain; // 3
  s[i++] = d->m_NcoLLimit; // 4
  s[i++] = d->m_NcoHLimit; // 5
  s[i++] = d->m_PLLAlpha; // 6
  s[i++] = d->m_PLLBeta; // 7
  s[i++] = d->m_DeemphasisAlpha; // 8
  s[i++] = p.m_minfreq; // 9
  s[i++] = p.m_maxfreq; // 10
  s[i++] = p.m_phasor_b0; // 11
  s[i++] = p.m_phasor_a1; // 12
  s[i++] = p.m_phasor_a2; // 13
  s[i++] = p.m_loopfilter_b0; // 14
  s[i++] = p.m_loopfilter_b1; // 15
  s[i++] = p.m_freq; // 16  (current value; == initial right after construction)
  s[i++] = p.m_minsignal; // 17
  s[i++] = p.m_lock_delay; // 18
  s[i++] = d->m_ReSampleInput.m_stateOrderSize; // 19
  s[i++] = d->m_ReSampleInput.m_downsample_int; // 20
  s[i++] = d->m_ReSampleMono.m_stateOrderSize; // 21
  s[i++] = d->m_ReSampleMono.m_downsample; // 22
  s[i++] = r.m_ProcessRate; // 23
  s[i++] = r.m_RdsNcoLLimit; // 24
  s[i++] = r.m_RdsNcoHLimit; // 25
  s[i++] = r.m_RdsPllAlpha; // 26
  s[i++] = r.m_RdsPllBeta; // 27
  s[i++] = r.m_RdsMatchCoefLength; // 28
  s[i++] = r.m_RdsLPFilter.m_NumTaps; // 29
  s[i++] = r.m_DownConvert.m_NcoInc; // 30
  s[i++] = r.m_DownConvert.m_OscCos; // 31
  s[i++] = r.m_DownConvert.m_OscSin; // 32
  s[i++] = r.m_RdsBitSyncFilter.m_A1; // 33
  s[i++] = r.m_RdsBitSyncFilter.m_A2; // 34
  s[i++] = r.m_RdsBitSyncFilter.m_B0; // 35
  s[i++] = r.m_RdsBitSyncFilter.m_B1; // 36
  s[i++] = r.m_RdsBitSyncFilter.m_B2; // 37
  s[i++] = d->m_NotchFilter.m_A1; // 38
  s[i++] = d->m_NotchFilter.m_A2; // 39
  s[i++] = d->m_NotchFilter.m_B0; // 40
  s[i++] = d->m_NotchFilter.m_B1; // 41
  s[i++] = d->m_NotchFilter.m_B2; // 42
  s[i++] = d->m_LPFilter.m_NumTaps; // 43
  int stages = 0;
  while (r.m_DownConvert.m_pDecimatorPtrs[stages])
    ++stages;
  s[i++] = stages; // 44
  for (int k = 0; k < 6; ++k) // 45..50: half-band lengths per stage (0 = none, 3 = CIC3)
  {
    double len = 0;
    if (k < stages)
    {
      auto* st = r.m_DownConvert.m_pDecimatorPtrs[k];
      if (auto* hb = dynamic_cast<CRDSDownConvert::CHalfBandDecimateBy2*>(st))
        len = hb->m_FirLength;
      else if (dynamic_cast<CRDSDownConvert::CHalfBand11TapDecimateBy2*>(st))
        len = 11;
      else
        len = 3;
    }
    s[i++] = len;
  }
}

// which: 0 fine-tuner table (cf32[64]) 1 input-FIR coeff (order+2) 2 audio-resampler coeff
// (order+2) 3 RDS LP coef (NumTaps) 4 RDS matched coef (len) 5 audio LP coef (NumTaps)
__attribute__((visibility("default"))) unsigned ref_fm_table(void* hv,
                                                             int which,
                                                             float* out,
                                                             unsigned max_floats)
{
  const cFmDecoder* d = static_cast<RefFm*>(hv)->dec;
  const float* src = nullptr;
  unsigned n = 0;
  switch (which)
  {
    case 0:
      src = reinterpret_cast<const float*>(d->m_FineTuner.m_table);
      n = 2 * d->m_FineTuner.m_tableSize;
      break;
    case 1:
      src = d->m_ReSampleInput.m_coeff;
      n = d->m_ReSampleInput.m_stateOrderSize + 2;
      break;
    case 2:
      src = d->m_ReSampleMono.m_coeff;
      n = d->m_ReSampleMono.m_stateOrderSize + 2;
      break;
    case 3:
      src = d->m_RDSProcess.m_RdsLPFilter.m_Coef;
      n = d->m_RDSProcess.m_RdsLPFilter.m_NumTaps;
      break;
    case 4:
      src = d->m_RDSProcess.m_RdsMatchedFilter.m_Coef;
      n = d->m_RDSProcess.m_RdsMatchedFilter.m_NumTaps;
      break;
    case 5:
      src = d->m_LPFilter.m_Coef;
      n = d->m_LPFilter.m_NumTaps;
      break;
    default:
      return 0;
  }
  n = std::min(n, max_floats);
  if (out)
    memcpy(out, src, n * sizeof(float));
  return n;
}

// ProcessStream with every intermediate tapped.  Same calls, same order as FmDecode.cpp:417-502
// and RDSProcess.cpp:120-180; additionally records the differential-decoded RDS data bits.
__attribute__((visibility("default"))) unsigned ref_fm_process_staged(void* hv,
                                                                      const float* iq,
                                                                      unsigned samples,
                                                                      float* audio,
                                                                      RefTaps* t)
{
  RefFm* h = static_cast<RefFm*>(hv);
  cFmDecoder& d = *h->dec;
  RefTaps dummy;
  memset(&dummy, 0, sizeof(dummy));
  if (!t)
    t = &dummy;

  unsigned dataSize = samples;
  d.m_FineTuner.Process(reinterpret_cast<const ComplexType*>(iq), d.m_BufferIfTuned, dataSize);
  CopyOut(t->tuned, d.m_BufferIfTuned, 2 * (size_t)dataSize);

  d.m_InterfaceLevel =
      0.95f * d.m_InterfaceLevel + 0.05f * d.RMSLevelApprox(d.m_BufferIfTuned, dataSize);

  dataSize = d.m_ReSampleInput.Process(d.m_BufferIfTuned, d.m_BufferDemod, dataSize);
  t->nb = dataSize;
  CopyOut(t->demod_in, d.m_BufferDemod, 2 * (size_t)dataSize);

  d.PhaseLockedLoop(d.m_BufferDemod, d.m_BufferBaseband, dataSize);
  CopyOut(t->baseband, d.m_BufferBaseband, dataSize);

  { // ---- cRDSRxSignalProcessor::Process, staged ----
    cRDSRxSignalProcessor& r = d.m_RDSProcess;
    const unsigned inLength = dataSize;
    for (unsigned i = 0; i < inLength; i++)
      r.m_ProcessArrayIn[i] = d.m_BufferBaseband[i];
    unsigned length = r.m_DownConvert.ProcessData(inLength, r.m_ProcessArrayIn, r.m_RdsRaw);
    t->nr = length;
    CopyOut(t->rds_dec, r.m_RdsRaw, 2 * (size_t)length);
    r.m_RdsLPFilter.Process(r.m_RdsRaw, length);
    CopyOut(t->rds_lp, r.m_RdsRaw, 2 * (size_t)length);
    r.ProcessRdsPll(r.m_RdsRaw, r.m_RdsData, length);
    CopyOut(t->rds_pll, r.m_RdsData, length);
    r.m_RdsMatchedFilter.Process(r.m_RdsData, length);
    CopyOut(t->rds_mf, r.m_RdsData, length);
    for (unsigned i = 0; i < length; i++)
      r.m_RdsMag[i] = r.m_RdsData[i] * r.m_RdsData[i];
    r.m_RdsBitSyncFilter.Process(r.m_RdsMag, length);
    CopyOut(t->rds_sync, r.m_RdsMag, length);
    for (unsigned i = 0; i < length; i++)
    {
      RealType Data = r.m_RdsData[i];
      RealType SyncVal = r.m_RdsMag[i];
      RealType Slope = SyncVal - r.m_RdsLastSync;
      r.m_RdsLastSync = SyncVal;
      if ((Slope < 0.0) && (r.m_RdsLastSyncSlope * Slope) < 0.0)
      {
        int bit = (r.m_RdsLastData >= 0) ? 1 : 0;
        r.m_RdsRaw[i].real(r.m_RdsLastData);
        h->bits.push_back((uint8_t)(bit ^ r.m_RdsLastBit));
        r.ProcessNewRdsBit(bit ^ r.m_RdsLastBit);
        r.m_RdsLastBit = bit;
      }
      else
      {
        r.m_RdsRaw[i].real(0);
      }
      r.m_RdsLastData = Data;
      r.m_RdsLastSyncSlope = Slope;
      r.m_RdsRaw[i].imag(Data);
    }
  }

  RealType baseband_mean, baseband_rms;
  d.SamplesMeanRMS(d.m_BufferBaseband, baseband_mean, baseband_rms, dataSize);
  d.m_BasebandMean = 0.95f * d.m_BasebandMean + 0.05f * baseband_mean;
  d.m_BasebandLevel = 0.95f * d.m_BasebandLevel + 0.05f * baseband_rms;

  unsigned monoSize = d.m_ReSampleMono.Process(d.m_BufferBaseband, d.m_BufferMono, dataSize);
  CopyOut(t->mono_rs, d.m_BufferMono, monoSize);

  d.m_StereoDetected = d.m_PilotPLL.Process(d.m_BufferBaseband, d.m_BufferRawStereo, dataSize);
  t->stereo = d.m_StereoDetected ? 1 : 0;
  CopyOut(t->pilot38, d.m_BufferRawStereo, dataSize);

  for (unsigned i = 0; i < dataSize; ++i)
    d.m_BufferRawStereo[i] *= 2 * d.m_BufferBaseband[i];
  CopyOut(t->rawstereo, d.m_BufferRawStereo, dataSize);

  dataSize = d.m_ReSampleStereo.Process(d.m_BufferRawStereo, d.m_BufferStereo, dataSize);
  t->na = dataSize;
  CopyOut(t->stereo_rs, d.m_BufferStereo, dataSize);

  d.m_LPFilter.ProcessTwo(d.m_BufferStereo, d.m_BufferMono, dataSize);
  if (t->lp)
  {
    CopyOut(t->lp, d.m_BufferStereo, dataSize);
    CopyOut(t->lp + dataSize, d.m_BufferMono, dataSize);
  }
  d.ProcessDeemphasisFilter(d.m_BufferStereo, d.m_BufferMono, dataSize);
  if (t->deemph)
  {
    CopyOut(t->deemph, d.m_BufferStereo, dataSize);
    CopyOut(t->deemph + dataSize, d.m_BufferMono, dataSize);
  }
  d.m_NotchFilter.ProcessTwo(d.m_BufferStereo, d.m_BufferMono, dataSize);
  if (t->notch)
  {
    CopyOut(t->notch, d.m_BufferStereo, dataSize);
    CopyOut(t->notch + dataSize, d.m_BufferMono, dataSize);
  }

  if (d.m_StereoDetected)
  {
    assert(dataSize == monoSize);
    for (unsigned i = 0; i < dataSize; ++i)
    {
      float m = d.m_BufferMono[i];
      float s = d.m_BufferStereo[i];
      audio[2 * i] = (m + s) * 0.5f;
      audio[2 * i + 1] = (m - s) * 0.5f;
    }
  }
  else
  {
    for (unsigned i = 0; i < dataSize; ++i)
    {
      float m = d.m_BufferMono[i] * 0.5f;
      audio[2 * i] = m;
      audio[2 * i + 1] = m;
    }
  }
  return 2 * dataSize;
}

// --- cFineTuner -------------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_finetuner_create(unsigned table_size, int shift)
{
  return new cFineTuner(table_size, shift);
}
__attribute__((visibility("default"))) void ref_finetuner_destroy(void* h)
{
  delete static_cast<cFineTuner*>(h);
}
__attribute__((visibility("default"))) void ref_finetuner_process(void* h,
                                                                  const float* in,
                                                                  float* out,
                                                                  unsigned n)
{
  static_cast<cFineTuner*>(h)->Process(reinterpret_cast<const ComplexType*>(in),
                                       reinterpret_cast<ComplexType*>(out), n);
}

// --- cPilotPhaseLock --------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_pilot_create(float freq, float bw, float minsig)
{
  return new cPilotPhaseLock(freq, bw, minsig);
}
__attribute__((visibility("default"))) void ref_pilot_destroy(void* h)
{
  delete static_cast<cPilotPhaseLock*>(h);
}
__attribute__((visibility("default"))) int ref_pilot_process(void* h,
                                                             const float* in,
                                                             float* out,
                                                             unsigned n)
{
  return static_cast<cPilotPhaseLock*>(h)->Process(in, out, n) ? 1 : 0;
}
__attribute__((visibility("default"))) float ref_pilot_level(void* h)
{
  return static_cast<cPilotPhaseLock*>(h)->GetPilotLevel();
}

// --- cFreqShift -------------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_freqshift_create(float nco_freq, float in_rate)
{
  return new cFreqShift(nco_freq, in_rate);
}
__attribute__((visibility("default"))) void ref_freqshift_destroy(void* h)
{
  delete static_cast<cFreqShift*>(h);
}
__attribute__((visibility("default"))) void ref_freqshift_reset(void* h)
{
  static_cast<cFreqShift*>(h)->Reset();
}
__attribute__((visibility("default"))) void ref_freqshift_process(void* h, float* iq, unsigned n)
{
  static_cast<cFreqShift*>(h)->Process(reinterpret_cast<ComplexType*>(iq), n);
}

// --- cDownsampleFilter ------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_downsample_create(unsigned order,
                                                                   double cutoff,
                                                                   double downsample,
                                                                   int integer_factor)
{
  return new cDownsampleFilter(order, cutoff, downsample, integer_factor != 0);
}
__attribute__((visibility("default"))) void ref_downsample_destroy(void* h)
{
  delete static_cast<cDownsampleFilter*>(h);
}
__attribute__((visibility("default"))) void ref_downsample_reset(void* h)
{
  static_cast<cDownsampleFilter*>(h)->Reset();
}
__attribute__((visibility("default"))) unsigned ref_downsample_process_real(void* h,
                                                                            const float* in,
                                                                            float* out,
                                                                            unsigned n)
{
  return static_cast<cDownsampleFilter*>(h)->Process(in, out, n);
}
__attribute__((visibility("default"))) unsigned ref_downsample_process_complex(void* h,
                                                                               const float* in,
                                                                               float* out,
                                                                               unsigned n)
{
  return static_cast<cDownsampleFilter*>(h)->Process(reinterpret_cast<const ComplexType*>(in),
                                                     reinterpret_cast<ComplexType*>(out), n);
}
__attribute__((visibility("default"))) unsigned ref_downsample_coeff(void* h, float* out)
{
  cDownsampleFilter* f = static_cast<cDownsampleFilter*>(h);
  unsigned n = f->m_stateOrderSize + 2;
  if (out)
    memcpy(out, f->m_coeff, n * sizeof(float));
  return n;
}

// --- CRDSDownConvert --------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_rdsdc_create()
{
  return new CRDSDownConvert();
}
__attribute__((visibility("default"))) void ref_rdsdc_destroy(void* h)
{
  delete static_cast<CRDSDownConvert*>(h);
}
__attribute__((visibility("default"))) void ref_rdsdc_set_frequency(void* h, float f)
{
  static_cast<CRDSDownConvert*>(h)->SetFrequency(f);
}
__attribute__((visibility("default"))) float ref_rdsdc_set_data_rate(void* h,
                                                                     float in_rate,
                                                                     float max_bw)
{
  return static_cast<CRDSDownConvert*>(h)->SetDataRate(in_rate, max_bw);
}
__attribute__((visibility("default"))) float ref_rdsdc_set_wfm_data_rate(void* h,
                                                                         float in_rate,
                                                                         float max_bw)
{
  return static_cast<CRDSDownConvert*>(h)->SetWfmDataRate(in_rate, max_bw);
}
// in-place on `inout` exactly like the reference (pInData is modified); result also in `out`.
__attribute__((visibility("default"))) int ref_rdsdc_process(void* h,
                                                             int n,
                                                             float* inout,
                                                             float* out)
{
  return static_cast<CRDSDownConvert*>(h)->ProcessData(n, reinterpret_cast<ComplexType*>(inout),
                                                       reinterpret_cast<ComplexType*>(out));
}
// stage lengths: returns number of stages, lens[k] = FIR length (11 = fixed 11-tap, 3 = CIC3)
__attribute__((visibility("default"))) int ref_rdsdc_stages(void* h, int* lens, int max)
{
  CRDSDownConvert* d = static_cast<CRDSDownConvert*>(h);
  int n = 0;
  while (n < MAX_DECSTAGES && d->m_pDecimatorPtrs[n])
  {
    if (n < max)
    {
      auto* st = d->m_pDecimatorPtrs[n];
      if (auto* hb = dynamic_cast<CRDSDownConvert::CHalfBandDecimateBy2*>(st))
        lens[n] = hb->m_FirLength;
      else if (dynamic_cast<CRDSDownConvert::CHalfBand11TapDecimateBy2*>(st))
        lens[n] = 11;
      else
        lens[n] = 3;
    }
    ++n;
  }
  return n;
}

// --- cFirFilter -------------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_fir_create()
{
  return new cFirFilter();
}
__attribute__((visibility("default"))) void ref_fir_destroy(void* h)
{
  delete static_cast<cFirFilter*>(h);
}
__attribute__((visibility("default"))) int ref_fir_init_lp(
    void* h, unsigned taps, float scale, float astop, float fpass, float fstop, float fs)
{
  return static_cast<cFirFilter*>(h)->InitLPFilter(taps, scale, astop, fpass, fstop, fs);
}
__attribute__((visibility("default"))) int ref_fir_init_hp(
    void* h, unsigned taps, float scale, float astop, float fpass, float fstop, float fs)
{
  return static_cast<cFirFilter*>(h)->InitHPFilter(taps, scale, astop, fpass, fstop, fs);
}
__attribute__((visibility("default"))) void ref_fir_init_const(void* h,
                                                               unsigned taps,
                                                               const float* coef,
                                                               float fs)
{
  static_cast<cFirFilter*>(h)->InitConstFir(taps, coef, fs);
}
__attribute__((visibility("default"))) void ref_fir_init_const_iq(
    void* h, unsigned taps, const float* icoef, const float* qcoef, float fs)
{
  static_cast<cFirFilter*>(h)->InitConstFir(taps, icoef, qcoef, fs);
}
__attribute__((visibility("default"))) unsigned ref_fir_coef(void* h, float* out)
{
  cFirFilter* f = static_cast<cFirFilter*>(h);
  if (out)
    memcpy(out, f->m_Coef, f->m_NumTaps * sizeof(float));
  return f->m_NumTaps;
}
__attribute__((visibility("default"))) void ref_fir_process_real(void* h, float* buf, unsigned n)
{
  static_cast<cFirFilter*>(h)->Process(buf, n);
}
__attribute__((visibility("default"))) void ref_fir_process_complex(void* h,
                                                                    float* buf,
                                                                    unsigned n)
{
  static_cast<cFirFilter*>(h)->Process(reinterpret_cast<ComplexType*>(buf), n);
}
__attribute__((visibility("default"))) void ref_fir_process_two(void* h,
                                                                float* a,
                                                                float* b,
                                                                unsigned n)
{
  static_cast<cFirFilter*>(h)->ProcessTwo(a, b, n);
}

// --- cIirFilter -------------------------------------------------------------------------------
__attribute__((visibility("default"))) void* ref_iir_create()
{
  return new cIirFilter();
}
__attribute__((visibility("default"))) void ref_iir_destroy(void* h)
{
  delete static_cast<cIirFilter*>(h);
}
__attribute__((visibility("default"))) int ref_iir_init(void* h, int type, float f0, float q, float fs)
{
  return static_cast<cIirFilter*>(h)->Init((eFilterType)type, f0, q, fs) ? 1 : 0;
}
// out: A1 A2 B0 B1 B2
__attribute__((visibility("default"))) void ref_iir_coef(void* h, float* out)
{
  cIirFilter* f = static_cast<cIirFilter*>(h);
  out[0] = f->m_A1;
  out[1] = f->m_A2;
  out[2] = f->m_B0;
  out[3] = f->m_B1;
  out[4] = f->m_B2;
}
__attribute__((visibility("default"))) void ref_iir_process_real(void* h, float* buf, unsigned n)
{
  static_cast<cIirFilter*>(h)->Process(buf, n);
}
__attribute__((visibility("default"))) void ref_iir_process_complex(void* h,
                                                                    float* buf,
                                                                    unsigned n)
{
  static_cast<cIirFilter*>(h)->Process(reinterpret_cast<ComplexType*>(buf), n);
}
__attribute__((visibility("default"))) void ref_iir_process_two(void* h,
                                                                float* a,
                                                                float* b,
                                                                unsigned n)
{
  static_cast<cIirFilter*>(h)->ProcessTwo(a, b, n);
}

// --- cRDSRxSignalProcessor (stand-alone) ------------------------------------------------------
__attribute__((visibility("default"))) void* ref_rds_create(float sample_rate)
{
  RefRds* h = new RefRds;
  h->rds = new cRDSRxSignalProcessor(nullptr, sample_rate);
  h->rds->m_InBitStream = 0;
  h->rds->m_BlockErrors = 0;
  RegisterSink(&h->rds->m_Decoder, &h->groups);
  return h;
}
__attribute__((visibility("default"))) void ref_rds_destroy(void* hv)
{
  RefRds* h = static_cast<RefRds*>(hv);
  UnregisterSink(&h->rds->m_Decoder);
  delete h->rds;
  delete h;
}
__attribute__((visibility("default"))) void ref_rds_reset(void* hv)
{
  static_cast<RefRds*>(hv)->rds->Reset();
}
__attribute__((visibility("default"))) void ref_rds_process(void* hv, const float* in, unsigned n)
{
  static_cast<RefRds*>(hv)->rds->Process(in, n);
}
__attribute__((visibility("default"))) unsigned ref_rds_take_groups(void* hv,
                                                                    uint16_t* out,
                                                                    unsigned max_groups)
{
  RefRds* h = static_cast<RefRds*>(hv);
  unsigned n = std::min<size_t>(h->groups.size() / 4, max_groups);
  if (out && n)
    memcpy(out, h->groups.data(), n * 4 * sizeof(uint16_t));
  h->groups.erase(h->groups.begin(), h->groups.begin() + n * 4);
  return n;
}
// Feed raw (already differentially decoded) bits straight into the block-sync / FEC state machine
// (RDSProcess.cpp:272-431) -- integer KATs.
__attribute__((visibility("default"))) void ref_rds_push_bits(void* hv,
                                                              const uint8_t* bits,
                                                              unsigned n)
{
  RefRds* h = static_cast<RefRds*>(hv);
  for (unsigned i = 0; i < n; ++i)
    h->rds->ProcessNewRdsBit(bits[i] & 1);
}
// CheckBlock on an explicit 26-bit word; returns syndrome, *corrected = word after FEC.
__attribute__((visibility("default"))) uint32_t ref_rds_check_block(
    void* hv, uint32_t word26, uint32_t offset_syndrome, int use_fec, uint32_t* corrected)
{
  RefRds* h = static_cast<RefRds*>(hv);
  uint32_t saved = h->rds->m_InBitStream;
  h->rds->m_InBitStream = word26;
  uint32_t syn = h->rds->CheckBlock(offset_syndrome, use_fec);
  if (corrected)
    *corrected = h->rds->m_InBitStream;
  h->rds->m_InBitStream = saved;
  return syn;
}

} // extern "C"

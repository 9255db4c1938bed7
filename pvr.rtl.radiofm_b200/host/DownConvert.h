// host/DownConvert.h -- cDownsampleFilter (DownConvert.h:21-60) and CRDSDownConvert (DownConvert.h:68-77) with the
// reference's signatures over the C ABI
// (rfm_downconvert_*).  The reference object is default-constructed and configured afterwards; the device object is
// (re)built when SetDataRate / SetWfmDataRate changes the plan, exactly when the reference rebuilds its stage list
// (DownConvert.cpp:331,382).
#pragma once

#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

// cDownsampleFilter (DownConvert.h:21-60) over rfm_downsample_*: the complex + integer and the real + fractional forms
// (the ones cFmDecoder instantiates, FmDecode.cpp:257-273) and real + integer; complex + fractional (an endless loop in
// the reference) returns 0 samples.
class cDownsampleFilter
{
public:
  cDownsampleFilter(unsigned int filter_order, double cutoff, double downsample = 1, bool integer_factor = true,
                    unsigned int max_len = 1u << 16, int cuda_device = -1)
  {
    if (rfm_downsample_create(1, filter_order, cutoff, downsample, integer_factor ? 1 : 0, max_len, cuda_device, &m_f) != RFM_OK)
      throw std::runtime_error(std::string("cDownsampleFilter (B200): ") + rfm_last_error());
  }
  virtual ~cDownsampleFilter() { rfm_downsample_destroy(m_f); }
  cDownsampleFilter(const cDownsampleFilter&) = delete;
  cDownsampleFilter& operator=(const cDownsampleFilter&) = delete;

  void Reset() { rfm_downsample_reset(m_f); } // DownConvert.cpp:90-96
  unsigned int Process(const RealType* samples_in, RealType* samples_out, unsigned int length) // :156-256
  {
    uint32_t n = 0;
    return rfm_downsample_process_real(m_f, samples_in, samples_out, length, &n) == RFM_OK ? n : 0;
  }
  unsigned int Process(const ComplexType* samples_in, ComplexType* samples_out, unsigned int length) // :98-154
  {
    uint32_t n = 0;
    return rfm_downsample_process_complex(m_f, reinterpret_cast<const float*>(samples_in),
                                          reinterpret_cast<float*>(samples_out), length, &n) == RFM_OK ? n : 0;
  }

private:
  rfm_downsample* m_f = nullptr;
};

class CRDSDownConvert
{
public:
  explicit CRDSDownConvert(unsigned int max_len = 1u << 16, int cuda_device = -1) : m_max_len(max_len), m_device(cuda_device) {}
  virtual ~CRDSDownConvert() { rfm_downconvert_destroy(m_dc); }
  CRDSDownConvert(const CRDSDownConvert&) = delete;
  CRDSDownConvert& operator=(const CRDSDownConvert&) = delete;

  void SetFrequency(RealType NcoFreq) // DownConvert.cpp:311-320
  {
    m_NcoFreq = NcoFreq + m_CW_Offset;
    if (m_dc)
      rfm_downconvert_set_frequency(m_dc, &m_NcoFreq);
  }
  void SetCwOffset(RealType offset) { m_CW_Offset = offset; } // DownConvert.h:74
  RealType SetDataRate(RealType InRate, RealType MaxBW) { return Plan(InRate, MaxBW, 0); }    // :327-371
  RealType SetWfmDataRate(RealType InRate, RealType MaxBW) { return Plan(InRate, MaxBW, 1); } // :378-399
  // :412-489.  Returns the number of output samples.  pInData is left untouched (the reference leaves its intermediate
  // results there); inputs the reference would mis-filter (odd / too short for a stage) return 0.
  int ProcessData(int InLength, ComplexType* pInData, ComplexType* pOutData)
  {
    if (!m_dc || InLength <= 0)
      return 0;
    uint32_t n = 0;
    if (rfm_downconvert_process_cf32(m_dc, reinterpret_cast<const float*>(pInData), (uint32_t)InLength,
                                     reinterpret_cast<float*>(pOutData), &n) != RFM_OK)
      return 0;
    return (int)n;
  }

private:
  RealType Plan(RealType InRate, RealType MaxBW, int wfm)
  {
    if (m_dc && InRate == m_InRate && MaxBW == m_MaxBW) // DownConvert.cpp:331,382: unchanged -> keep the stages
      return rfm_downconvert_output_rate(m_dc);
    rfm_downconvert_destroy(m_dc);
    m_dc = nullptr;
    m_InRate = InRate;
    m_MaxBW = MaxBW;
    if (rfm_downconvert_create(1, &m_NcoFreq, InRate, MaxBW, wfm, m_max_len, m_device, &m_dc) != RFM_OK)
      throw std::runtime_error(std::string("CRDSDownConvert (B200): ") + rfm_last_error());
    return rfm_downconvert_output_rate(m_dc);
  }

  unsigned int m_max_len;
  int m_device;
  RealType m_NcoFreq = 0.0f, m_CW_Offset = 0.0f, m_InRate = 0.0f, m_MaxBW = 0.0f;
  rfm_downconvert* m_dc = nullptr;
};

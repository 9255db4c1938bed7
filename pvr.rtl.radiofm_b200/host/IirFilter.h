// host/IirFilter.h -- cIirFilter with the reference's signatures (IirFilter.h:12-36) over the C ABI (rfm_iir_*).
// One object = one biquad on the device (rows = 1); buffers are filtered in place, delays carried between calls.
#pragma once

#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

class cIirFilter
{
public:
  explicit cIirFilter(unsigned int max_len = 1u << 16, int cuda_device = -1)
  {
    if (rfm_iir_create(1, max_len, cuda_device, &m_f) != RFM_OK)
      throw std::runtime_error(std::string("cIirFilter (B200): ") + rfm_last_error());
  }
  virtual ~cIirFilter() { rfm_iir_destroy(m_f); }
  cIirFilter(const cIirFilter&) = delete;
  cIirFilter& operator=(const cIirFilter&) = delete;

  bool Init(eFilterType type, RealType F0Freq, RealType FilterQ, RealType SampleRate) // IirFilter.cpp:11-60
  {
    return rfm_iir_init(m_f, (int)type, F0Freq, FilterQ, SampleRate) == RFM_OK;
  }
  void Process(ComplexType* buffer, unsigned int length) // :62-76
  {
    rfm_iir_process_complex(m_f, reinterpret_cast<float*>(buffer), length);
  }
  void Process(RealType* buffer, unsigned int length) { rfm_iir_process_real(m_f, buffer, length); } // :78-87
  void ProcessTwo(RealType* bufferA, RealType* bufferB, unsigned int length) // :89-105
  {
    rfm_iir_process_two(m_f, bufferA, bufferB, length);
  }

private:
  rfm_iir* m_f = nullptr;
};

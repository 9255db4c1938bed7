// host/FreqShift.h -- cFreqShift with the reference's signatures (FreqShift.h:12-27) over the C ABI.
#pragma once

#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

class cFreqShift
{
public:
  cFreqShift(RealType NcoFreq, RealType InRate, unsigned int max_len = 1u << 20, int cuda_device = -1)
  {
    if (rfm_freqshift_create(1, &NcoFreq, InRate, max_len, cuda_device, &m_fs) != RFM_OK)
      throw std::runtime_error("cFreqShift (B200): no usable CUDA device");
  }
  virtual ~cFreqShift() { rfm_freqshift_destroy(m_fs); }
  cFreqShift(const cFreqShift&) = delete;
  cFreqShift& operator=(const cFreqShift&) = delete;

  void Reset() { rfm_freqshift_reset(m_fs); }
  void Process(ComplexType* pInData, unsigned int InLength)
  {
    rfm_freqshift_process_cf32(m_fs, reinterpret_cast<float*>(pInData), InLength);
  }

private:
  rfm_freqshift* m_fs = nullptr;
};

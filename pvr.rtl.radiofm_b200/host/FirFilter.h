// host/FirFilter.h -- cFirFilter with the reference's signatures (FirFilter.h:17-60) over the C ABI (rfm_fir_*): the
// members the live chain uses -- InitLPFilter, InitConstFir (real taps), Process (real / complex, in place),
// ProcessTwo -- plus InitHPFilter and the complex two-buffer Process.  GenerateHBFilter, the I/Q form of InitConstFir
// and Process(RealType*, ComplexType*) (whose sum is an assignment in the reference, FirFilter.cpp:466-468) are not
// called anywhere on the hot path and are not provided.
#pragma once

#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

#define MAX_NUMCOEF 75 // FirFilter.h:15

class cFirFilter
{
public:
  explicit cFirFilter(unsigned int max_len = 1u << 16, int cuda_device = -1)
  {
    if (rfm_fir_create(1, max_len, cuda_device, &m_f) != RFM_OK)
      throw std::runtime_error(std::string("cFirFilter (B200): ") + rfm_last_error());
  }
  virtual ~cFirFilter() { rfm_fir_destroy(m_f); }
  cFirFilter(const cFirFilter&) = delete;
  cFirFilter& operator=(const cFirFilter&) = delete;

  void InitConstFir(unsigned int NumTaps, const RealType* pCoef, RealType Fsamprate) // FirFilter.cpp:302-320
  {
    rfm_fir_init_const(m_f, NumTaps, pCoef, Fsamprate);
  }
  int InitLPFilter(unsigned int NumTaps, RealType Scale, RealType Astop, RealType Fpass, RealType Fstop,
                   RealType Fsamprate) // FirFilter.cpp:78-148
  {
    uint32_t n = 0;
    rfm_fir_init_lp(m_f, NumTaps, Scale, Astop, Fpass, Fstop, Fsamprate, &n);
    return (int)n;
  }
  int InitHPFilter(unsigned int NumTaps, RealType Scale, RealType Astop, RealType Fpass, RealType Fstop,
                   RealType Fsamprate) // FirFilter.cpp:195-264
  {
    uint32_t n = 0;
    rfm_fir_init_hp(m_f, NumTaps, Scale, Astop, Fpass, Fstop, Fsamprate, &n);
    return (int)n;
  }
  void Process(ComplexType* buffer, unsigned int length) // :330-350
  {
    rfm_fir_process_complex(m_f, reinterpret_cast<float*>(buffer), length);
  }
  void Process(RealType* buffer, unsigned int length) { rfm_fir_process_real(m_f, buffer, length); } // :360-377
  // :421-445 -- the in-place complex form writing to a second buffer (InBuf is left untouched)
  void Process(ComplexType* InBuf, ComplexType* OutBuf, unsigned int length)
  {
    if (OutBuf != InBuf)
      std::memcpy(static_cast<void*>(OutBuf), InBuf, (size_t)length * sizeof(ComplexType));
    Process(OutBuf, length);
  }
  void ProcessTwo(RealType* bufferA, RealType* bufferB, unsigned int length) // :387-413
  {
    rfm_fir_process_two(m_f, bufferA, bufferB, length);
  }

private:
  rfm_fir* m_f = nullptr;
};

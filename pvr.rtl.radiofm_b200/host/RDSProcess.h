// host/RDSProcess.h -- cRDSRxSignalProcessor with the reference's signatures (RDSProcess.h:56-64) over the C ABI
// (rfm_rdsproc_*).  The reference hands every decoded group to its cRDSGroupDecoder from inside Process
// (RDSProcess.cpp:312,355); here the groups decoded during a Process call go, in order and on the calling thread, to
// the sink set with SetGroupSink() before Process returns (the bits themselves to SetBitSink(), if set).
#pragma once

#include <stdint.h>

#include <functional>
#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

class cRDSRxSignalProcessor
{
public:
  cRDSRxSignalProcessor(cRadioReceiver* proc, RealType SampleRate, unsigned int max_len = 1u << 16, int cuda_device = -1)
    : m_proc(proc)
  {
    if (rfm_rdsproc_create(1, SampleRate, max_len, cuda_device, &m_r) != RFM_OK)
      throw std::runtime_error(std::string("cRDSRxSignalProcessor (B200): ") + rfm_last_error());
  }
  virtual ~cRDSRxSignalProcessor() { rfm_rdsproc_destroy(m_r); }
  cRDSRxSignalProcessor(const cRDSRxSignalProcessor&) = delete;
  cRDSRxSignalProcessor& operator=(const cRDSRxSignalProcessor&) = delete;

  void Reset() { rfm_rdsproc_reset(m_r); } // RDSProcess.cpp:90-118

  void Process(const RealType* inputStream, unsigned int items) // RDSProcess.cpp:120-180
  {
    if (rfm_rdsproc_process(m_r, inputStream, items) != RFM_OK)
      return;
    uint32_t n = 0;
    if (m_bit_sink)
    {
      uint8_t bits[256];
      do
      {
        if (rfm_rdsproc_take_bits(m_r, 0, bits, 256, &n) != RFM_OK)
          return;
        for (uint32_t i = 0; i < n; ++i)
          m_bit_sink(bits[i]);
      } while (n == 256);
    }
    uint16_t blk[64][4];
    do
    {
      if (rfm_rdsproc_take_groups(m_r, 0, &blk[0][0], 64, &n) != RFM_OK)
        return;
      for (uint32_t i = 0; i < n && m_group_sink; ++i)
        m_group_sink(blk[i]);
    } while (n == 64);
  }

  void SetGroupSink(std::function<void(uint16_t*)> sink) { m_group_sink = std::move(sink); }
  void SetBitSink(std::function<void(int)> sink) { m_bit_sink = std::move(sink); }
  cRadioReceiver* Receiver() const { return m_proc; }

private:
  cRadioReceiver* m_proc;
  rfm_rdsproc* m_r = nullptr;
  std::function<void(uint16_t*)> m_group_sink;
  std::function<void(int)> m_bit_sink;
};

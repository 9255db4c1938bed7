// rfm_rdssync.h -- RDS block synchronisation + (26,16) shortened-cyclic-code FEC on the host.
// Integer-only part of cRDSRxSignalProcessor (RDSProcess.cpp:13-41, 272-431); SURVEY.md section 8a row
// a12 keeps it on the host: 1187.5 bit/s per stream, branchy, bit-exact by construction.
#pragma once

#include <stdint.h>

#include <vector>

namespace rfm
{

// syndrome of the low 26 bits of *in_bits for the given offset word; with use_fec a correctable burst
// (<= 5 bits) is repaired in place (Meggitt decoder).  Returns 0 when the block is good.
uint32_t RdsCheckBlock(uint32_t* in_bits, uint32_t offset_syndrome, bool use_fec);

class RdsBlockSync
{
public:
  RdsBlockSync() { Reset(); }
  // the part of cRDSRxSignalProcessor::Reset that concerns the decoder state (RDSProcess.cpp:108-118);
  // the shift register and the block-error counter are deliberately left alone, as in the reference.
  void Reset();
  void PushBit(int bit); // ProcessNewRdsBit
  std::vector<uint16_t>& Groups() { return m_groups; } // 4 words per decoded group

private:
  enum State { BITSYNC = 0, BLOCKSYNC = 1, GROUPDECODE = 2, GROUPRESYNC = 3 };
  uint32_t m_in = 0;
  int m_block = 0, m_bitpos = 0, m_state = BITSYNC, m_bgroup = 0, m_errors = 0;
  uint16_t m_data[4] = {0, 0, 0, 0};
  std::vector<uint16_t> m_groups;
};

} // namespace rfm

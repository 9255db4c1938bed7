// rfm_rdssync.cpp -- see rfm_rdssync.h
#include "rfm_rdssync.h"

namespace rfm
{

namespace
{
// offset-word syndromes A, B, C, D / A, B, C', D (RDSProcess.h:49-53, RDSProcess.cpp:13-17)
const uint32_t kOffsetSyndrome[8] = {0x3D8, 0x3D4, 0x25C, 0x258, 0x3D8, 0x3D4, 0x3CC, 0x258};
// parity-check matrix rows for the 16 message bits (RDSProcess.cpp:24-41)
const uint32_t kParity[16] = {0x2DC, 0x16E, 0x0B7, 0x287, 0x39F, 0x313, 0x355, 0x376,
                              0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};
const uint32_t kCrcPoly = 0x5B9; // x^10+x^8+x^7+x^5+x^4+x^3+1
const int kBitsPerBlock = 26;
const int kBlockErrorLimit = 0;  // BLOCK_ERROR_LIMIT, RDSProcess.h:31

// The syndrome is linear in the 26 received bits: bits 25..16 map to themselves (identity part of the check matrix),
// bit 15-i to kParity[i].  One table per byte of the word turns the reference's 16-step loop (RDSProcess.cpp:386-396)
// into four lookups -- the bit-sync search evaluates it for every incoming bit of every stream.
struct SyndromeTables
{
  uint16_t t[4][256];
  SyndromeTables()
  {
    for (int k = 0; k < 4; ++k)
      for (int v = 0; v < 256; ++v)
      {
        uint32_t syn = 0;
        for (int b = 0; b < 8; ++b)
        {
          const int j = 8 * k + b; // bit position in the 26-bit word
          if (!((v >> b) & 1) || j > 25)
            continue;
          syn ^= (j >= 16) ? (1u << (j - 16)) : kParity[15 - j];
        }
        t[k][v] = (uint16_t)syn;
      }
  }
};
const SyndromeTables kSyn;
} // namespace

uint32_t RdsCheckBlock(uint32_t* in_bits, uint32_t offset_syndrome, bool use_fec)
{
  const uint32_t block = *in_bits & 0x3FFFFFF;
  uint32_t syn = (uint32_t)(kSyn.t[0][block & 0xff] ^ kSyn.t[1][(block >> 8) & 0xff] ^ kSyn.t[2][(block >> 16) & 0xff] ^
                            kSyn.t[3][block >> 24]);
  syn ^= offset_syndrome;
  if (syn != 0 && use_fec)
  {
    uint32_t mask = 1u << (kBitsPerBlock - 1);
    for (int i = 0; i < 16; ++i, mask >>= 1)
    {
      const bool msb = (syn & 0x200) != 0;
      const bool trap = msb && (syn & 0x1F) == 0;
      if (trap)
        *in_bits ^= mask; // correct this message bit
      syn <<= 1;
      if (msb && !trap)
        syn ^= kCrcPoly;
    }
    syn &= 0x3FF;
  }
  return syn;
}

void RdsBlockSync::Reset()
{
  m_bitpos = 0;
  m_block = 0;
  m_state = BITSYNC;
  m_bgroup = 0;
}

void RdsBlockSync::PushBit(int bit)
{
  m_in = (m_in << 1) | (uint32_t)(bit & 1);
  if (m_state == BITSYNC)
  {
    // slide bit by bit until a clean block A shows up (no FEC)
    if (RdsCheckBlock(&m_in, kOffsetSyndrome[0], false) == 0)
    {
      m_bitpos = 0;
      m_bgroup = 0;
      m_data[0] = (uint16_t)(m_in >> 10);
      m_block = 1;
      m_state = BLOCKSYNC;
    }
    return;
  }
  if (++m_bitpos < kBitsPerBlock)
    return;
  m_bitpos = 0;
  if (m_state == GROUPRESYNC)
  {
    // skip the rest of a damaged group
    if (++m_block > 3)
    {
      m_block = 0;
      m_state = GROUPDECODE;
    }
    return;
  }
  const bool decoding = (m_state == GROUPDECODE);
  const uint32_t syn = RdsCheckBlock(&m_in, kOffsetSyndrome[m_block + m_bgroup], decoding);
  if (syn != 0)
  {
    if (!decoding)
    {
      m_state = BITSYNC;
      return;
    }
    if (++m_errors > kBlockErrorLimit)
    {
      m_state = BITSYNC;
      return;
    }
    if (++m_block > 3)
      m_block = 0;
    if (m_block != 0)
      m_state = GROUPRESYNC;
    return;
  }
  m_data[m_block] = (uint16_t)(m_in >> 10);
  m_bgroup = (m_block == 1 && (m_data[1] & 0x0800)) ? 4 : 0; // version B groups use C'
  if (m_block >= 3)
  {
    m_block = 0;
    m_errors = 0;
    m_state = GROUPDECODE;
    m_groups.insert(m_groups.end(), m_data, m_data + 4);
  }
  else
  {
    ++m_block;
  }
}

} // namespace rfm

// rfm_rdsgroup.cpp -- see rfm_rdsgroup.h.  Behaviour follows cRDSGroupDecoder (RDSGroupDecoder.cpp, built with its
// default macros: IMPROVE_CHECK* and IMPROVE_CHECK_ALT_FREQ undefined, so the alternative-frequency scan and the
// diagnostics are not part of it) frame for frame and byte for byte; the structure here is its own.
#include "rfm_rdsgroup.h"

#include <string.h>

#include <initializer_list>
#include <string>

#include "../../include/radiofm_b200.h"

namespace rfm
{

namespace
{
inline uint8_t Hi(uint16_t w) { return (uint8_t)(w >> 8); }
inline uint8_t Lo(uint16_t w) { return (uint8_t)(w & 0xff); }

// CRC16-CCITT (x^16 + x^12 + x^5 + 1), start 0xffff, result inverted: RDSGroupDecoder.cpp:965-981, bitwise form
uint16_t FrameCrc(const uint8_t* p, int n)
{
  unsigned crc = 0xffffu;
  for (int i = 0; i < n; ++i)
  {
    crc ^= (unsigned)p[i] << 8;
    for (int b = 0; b < 8; ++b)
      crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xffffu : (crc << 1) & 0xffffu;
  }
  return (uint16_t)(~crc & 0xffffu);
}
} // namespace

size_t UecpStuffFrame(const uint8_t* frame, uint32_t len, std::vector<uint8_t>& out)
{
  const size_t before = out.size();
  out.push_back(0xFE);
  for (uint32_t i = 0; i < len; ++i)
  {
    const uint8_t v = frame[i];
    if (v < 0xFD)
      out.push_back(v);
    else
    {
      out.push_back(0xFD);
      out.push_back((uint8_t)((v & 3) - 1));
    }
  }
  out.push_back(0xFF);
  return out.size() - before;
}

RdsGroupDecoder::RdsGroupDecoder(const RdsGroupSink* sink)
{
  if (sink)
    m_sink = *sink;
  // the reference's constructor initialises nothing (RDSGroupDecoder.cpp:136-138); Reset() first runs from
  // Decode_PI when the first group arrives with a PI different from the zero the storage holds
}

void RdsGroupDecoder::SetSink(const RdsGroupSink* sink) { m_sink = sink ? *sink : RdsGroupSink(); }

void RdsGroupDecoder::Reset()
{
  m_pi = 0;
  m_rt_segments = 0;
  m_rt_count = 0;
  m_rt_first = false;
  m_di = 0;
  m_di_prev = 0xff;   // uint8_t(-1)
  m_ms = 0;
  m_ms_prev = 0xff;
  m_pin = 0xffff;     // uint16_t(-1)
  m_ptyn_set = 0;
  m_ps_set = 0;
  m_ta_tp = -1;
  m_rtplus_ready = false;
  memset(m_rt, 0, sizeof(m_rt));
  memset(m_oda, 0, sizeof(m_oda));
  memset(m_ptyn, 0x20, sizeof(m_ptyn));
  memset(m_ps, 0x20, sizeof(m_ps)); // including the terminator the reference had just written (:153,163)
}

bool RdsGroupDecoder::SettingActive() const
{
  return m_sink.is_setting_active ? m_sink.is_setting_active(m_sink.user) != 0 : false;
}

bool RdsGroupDecoder::NameAccepted(const char* name)
{
  if (m_sink.set_channel_name)
    return m_sink.set_channel_name(m_sink.user, name) != 0;
  memcpy(m_name, name, 8);
  m_name[8] = 0;
  return true; // cRadioReceiver::SetChannelName without a settings dialog, RadioReceiver.cpp:600-612
}

void RdsGroupDecoder::Begin()
{
  m_frame[0] = 0;     // ADD
  m_frame[1] = 0;
  m_frame[2] = m_seq; // SQC: the count BEFORE this frame is sent (:950-961 against :986)
  m_frame[3] = 0;     // MFL, filled in by Send
  m_len = 0;
}

void RdsGroupDecoder::Put(uint8_t v)
{
  if (m_len > 255)
    return; // "value reach end of allowed UECP frame size", :996-1000
  m_frame[4 + m_len++] = v;
}

void RdsGroupDecoder::Send()
{
  if (SettingActive())
    return;
  ++m_seq;
  m_frame[3] = (uint8_t)m_len;
  const uint16_t crc = FrameCrc(m_frame, m_len + 4);
  m_frame[4 + m_len] = Hi(crc);
  m_frame[5 + m_len] = Lo(crc);
  const uint32_t total = (uint32_t)m_len + 6;
  if (m_sink.add_uecp_frame)
  {
    m_sink.add_uecp_frame(m_sink.user, m_frame, total);
    return;
  }
  if (m_out.size() > 16384)
    return; // RadioReceiver.cpp:389-390
  UecpStuffFrame(m_frame, total, m_out);
}

// one whole frame: every argument is truncated to a byte, as AddStuffingValue(uint8_t) does
#define RFM_EMIT(...)                                                  \
  do                                                                   \
  {                                                                    \
    const long long bytes_[] = {__VA_ARGS__};                          \
    Begin();                                                           \
    for (long long v_ : bytes_)                                        \
      Put((uint8_t)v_);                                                \
    Send();                                                            \
  } while (0)

void RdsGroupDecoder::Decode(const uint16_t blk[4])
{
  const unsigned type = (blk[1] >> 11) & 0x1Fu; // group type code + version bit B0
  const bool version_b = (type & 1u) != 0;
  if (blk[0] != m_pi)
    OnPI(blk[0]);
  const int pty = (blk[1] >> 5) & 0x1F;
  if (pty != m_pty)
    OnPTY(pty);

  switch (type >> 1)
  {
    case 0: Type0(blk); return;
    case 1: Type1(blk, version_b); return;
    case 2: Type2(blk, version_b); return;
    case 10:
      if (!version_b)
      {
        Type10A(blk);
        return;
      }
      break;
    case 14: return; // 14A / 14B: EON, nothing forwarded (:863-874)
    case 15: return; // 15A / 15B
    default: break;
  }
  if (type == 0x06) // 3A
  {
    Type3A(blk);
    return;
  }
  if (type == 0x08) // 4A
  {
    Type4A(blk);
    return;
  }
  // every other type belongs to an Open Data Application when a 3A group has registered one for it; of the
  // remaining native handlers only 8A (TMC) forwards anything (:200-270)
  if (m_oda[type] > 0)
    Oda(blk, m_oda[type]);
  else if (type == 0x10)
    Type8A(blk);
}

void RdsGroupDecoder::OnPI(uint16_t pi)
{
  Reset();
  m_pi = pi;
  RFM_EMIT(MEC_PI, 0x00, 0x01, Lo(pi), Hi(pi));
}

void RdsGroupDecoder::OnPTY(int pty)
{
  m_pty = pty; // always 0..31 here
  RFM_EMIT(MEC_PTY, 0x00, 0x01, pty);
}

// 0A / 0B: decoder identification, TA / TP, music / speech, programme service name (:307-408)
void RdsGroupDecoder::Type0(const uint16_t* b)
{
  const unsigned seg = b[1] & 3u;
  {
    const uint8_t bit = (uint8_t)(8u >> seg); // segment 0 carries d3 ... segment 3 carries d0
    if (b[1] & 0x04)
      m_di |= bit;
    else
      m_di &= (uint8_t)~bit;
    ++m_di_seen;
  }
  const int ta_tp = ((b[1] & 0x10) ? 1 : 0) | ((b[1] & 0x400) ? 2 : 0);
  if (ta_tp != m_ta_tp)
  {
    m_ta_tp = ta_tp;
    RFM_EMIT(MEC_TA_TP, 0x00, 0x01, ta_tp);
  }
  if (m_di_seen >= 4 && m_di_prev != m_di)
  {
    m_di_seen = 0;
    m_di_prev = m_di;
    RFM_EMIT(MEC_DI, 0x00, 0x01, m_di & 0xf);
  }
  m_ms = (b[1] & 0x08) ? 1 : 0;
  if (m_ms_prev != m_ms)
  {
    m_ms_prev = m_ms;
    RFM_EMIT(MEC_MS, 0x00, 0x01, m_ms);
  }

  m_ps_work[2 * seg] = (char)Hi(b[3]);
  m_ps_work[2 * seg + 1] = (char)Lo(b[3]);
  m_ps_set |= 1 << seg;
  if (m_ps_set != 0x0F)
    return;
  // all four segments seen: publish when the name changed (or the settings dialog is open).  An unchanged name
  // leaves the flags set, so from then on every single changed segment publishes at once -- as the reference does.
  if (!SettingActive() && memcmp(m_ps, m_ps_work, 8) == 0)
    return;
  if (NameAccepted(m_ps_work))
  {
    Begin();
    Put(MEC_PS);
    Put(0x00);
    Put(0x01);
    for (int i = 0; i < 8; ++i)
      Put((uint8_t)m_ps_work[i]);
    Send();
    memcpy(m_ps, m_ps_work, 8);
  }
  m_ps_set = 0;
}

// 1A / 1B: programme item number, slow labelling codes (:556-587)
void RdsGroupDecoder::Type1(const uint16_t* b, bool version_b)
{
  if (m_pin != b[3])
  {
    m_pin = b[3];
    RFM_EMIT(MEC_PIN, 0x00, 0x01, Hi(m_pin), Lo(m_pin));
  }
  if (!version_b)
    RFM_EMIT(MEC_SLOW_LABEL, 0x00, Hi(b[2]) & 0x7F, Lo(b[2]));
}

// 2A / 2B: RadioText (:592-658)
void RdsGroupDecoder::Type2(const uint16_t* b, bool version_b)
{
  const unsigned seg = b[1] & 0x0fu;
  m_rtplus_ready = false;

  if (seg == 0 && m_rt_first && m_rt_count > 1)
  {
    // a text is complete when segments 0 .. count-1 have all been seen since the last segment 0
    bool complete = true;
    for (int i = 0; i < m_rt_count; ++i)
      if (i >= 32 || !(m_rt_segments & (1u << i)))
      {
        complete = false;
        m_rt_segments = 0;
        m_rt_count = 0;
        break;
      }
    if (complete)
    {
      Begin();
      Put(MEC_RT);
      Put(0x00);
      Put(0x01);
      Put(65);
      Put((uint8_t)m_rt_ab);
      for (int i = 0; i < 64; ++i)
        Put((uint8_t)m_rt[i]);
      Send();
      m_rtplus_ready = true;
    }
  }

  const int ab = (b[1] >> 4) & 1;
  if (m_rt_ab != ab)
  {
    memset(m_rt, 0x20, sizeof(m_rt));
    m_rt_ab = ab;
    m_rt_first = false;
    m_rt_segments = 0;
    m_rt_count = 0;
  }
  if (!version_b)
  {
    m_rt[4 * seg] = (char)Hi(b[2]);
    m_rt[4 * seg + 1] = (char)Lo(b[2]);
    m_rt[4 * seg + 2] = (char)Hi(b[3]);
    m_rt[4 * seg + 3] = (char)Lo(b[3]);
  }
  else
  {
    m_rt[2 * seg] = (char)Hi(b[3]);
    m_rt[2 * seg + 1] = (char)Lo(b[3]);
  }
  m_rt_segments |= 1u << seg;
  ++m_rt_count;
  if (!m_rt_first && seg == 0)
    m_rt_first = true;
}

// 3A: application identification for Open Data (:663-706)
void RdsGroupDecoder::Type3A(const uint16_t* b)
{
  const unsigned app_group = b[1] & 0x1Fu;
  RFM_EMIT(MEC_ODA_CONF, app_group, Hi(b[3]), Lo(b[3]), 0, Hi(b[2]), Lo(b[2]), 0);
  const int aid = b[3];
  m_oda[app_group] = (aid == AID_RTPLUS || aid == AID_TFC) ? aid : 0;
}

// 4A: clock time and date (:711-741); the arithmetic keeps the reference's types (double MJD, truncations to int,
// unsigned year / month / day)
void RdsGroupDecoder::Type4A(const uint16_t* b)
{
  const double mjd = (double)(((b[1] & 0x03) << 15) | ((b[2] >> 1) & 0x7fff));
  const unsigned hours = ((b[2] & 0x01u) << 4) | ((b[3] >> 12) & 0x0fu);
  const unsigned minutes = (b[3] >> 6) & 0x3fu;
  const int offset = b[3] & 0x3f;
  unsigned year = (unsigned)(int)((mjd - 15078.2) / 365.25);
  unsigned month = (unsigned)(int)((mjd - 14956.1 - (int)(year * 365.25)) / 30.6001);
  const unsigned day = (unsigned)(mjd - 14956 - (int)(year * 365.25) - (int)(month * 30.6001));
  const int k = (month == 14 || month == 15) ? 1 : 0;
  year += (unsigned)(k + 1900);
  month -= (unsigned)(1 + k * 12);
  RFM_EMIT(MEC_RTC, (int)(year % 100), (int)month, (int)day, (int)hours, (int)minutes, 0, 0, offset);
}

// 8A: traffic message channel (:789-803)
void RdsGroupDecoder::Type8A(const uint16_t* b)
{
  RFM_EMIT(MEC_TMC, 6, 0, b[1] & 0x1F, Hi(b[2]), Lo(b[2]), Hi(b[3]), Lo(b[3]));
}

// 10A: programme type name (:821-853); a frame goes out with every group once either half has been seen
void RdsGroupDecoder::Type10A(const uint16_t* b)
{
  const unsigned seg = b[1] & 1u;
  const bool ab = ((b[1] >> 4) & 1u) != 0;
  if (m_ptyn_ab != ab)
  {
    memset(m_ptyn, 0x20, 8);
    m_ptyn_ab = ab;
    m_ptyn_set = 0;
  }
  m_ptyn[4 * seg] = (char)Hi(b[2]);
  m_ptyn[4 * seg + 1] = (char)Lo(b[2]);
  m_ptyn[4 * seg + 2] = (char)Hi(b[3]);
  m_ptyn[4 * seg + 3] = (char)Lo(b[3]);
  m_ptyn_set |= 1 << seg;
  Begin();
  Put(MEC_PTYN);
  Put(0x00);
  Put(0x01);
  for (int i = 0; i < 8; ++i)
    Put((uint8_t)m_ptyn[i]);
  Send();
}

// groups claimed by an Open Data Application (:922-960)
void RdsGroupDecoder::Oda(const uint16_t* b, int aid)
{
  if (aid == AID_RTPLUS)
  {
    if (!m_rtplus_ready)
      return;
    RFM_EMIT(MEC_ODA_DATA, 8, 0x4b, 0xd7, Hi(b[1]), Lo(b[1]), Hi(b[2]), Lo(b[2]), Hi(b[3]), Lo(b[3]));
    m_rtplus_ready = false;
  }
  else if (aid == AID_TFC)
    RFM_EMIT(MEC_ODA_DATA, 7, 0xcd, 0x46, Lo(b[1]), Hi(b[2]), Lo(b[2]), Hi(b[3]), Lo(b[3]));
}

#undef RFM_EMIT

} // namespace rfm

// ==================================================================================================
// C ABI
// ==================================================================================================
struct rfm_rdsgroup
{
  rfm::RdsGroupDecoder dec;
};

extern "C"
{

int rfm_rdsgroup_create(const rfm_rdsgroup_callbacks* cb, rfm_rdsgroup** out)
{
  if (!out)
    return RFM_ERR_INVALID;
  rfm_rdsgroup* g = new rfm_rdsgroup();
  if (cb)
  {
    rfm::RdsGroupSink s;
    s.user = cb->user;
    s.add_uecp_frame = cb->add_uecp_frame;
    s.set_channel_name = cb->set_channel_name;
    s.is_setting_active = cb->is_setting_active;
    g->dec.SetSink(&s);
  }
  *out = g;
  return RFM_OK;
}

void rfm_rdsgroup_destroy(rfm_rdsgroup* g) { delete g; }

void rfm_rdsgroup_reset(rfm_rdsgroup* g)
{
  if (g)
    g->dec.Reset();
}

int rfm_rdsgroup_decode(rfm_rdsgroup* g, const uint16_t* blocks, uint32_t n_groups)
{
  if (!g || (!blocks && n_groups))
    return RFM_ERR_INVALID;
  for (uint32_t i = 0; i < n_groups; ++i)
    g->dec.Decode(blocks + 4 * (size_t)i);
  return RFM_OK;
}

int rfm_rdsgroup_take_uecp(rfm_rdsgroup* g, uint8_t* out, uint32_t cap, uint32_t* n)
{
  if (!g || !n)
    return RFM_ERR_INVALID;
  std::vector<uint8_t>& p = g->dec.Pending();
  const uint32_t k = (uint32_t)std::min<size_t>(p.size(), out ? cap : 0);
  if (k)
    memcpy(out, p.data(), k);
  p.erase(p.begin(), p.begin() + k);
  *n = k;
  return RFM_OK;
}

int rfm_rdsgroup_channel_name(const rfm_rdsgroup* g, char out[9])
{
  if (!g || !out)
    return RFM_ERR_INVALID;
  memcpy(out, g->dec.ChannelName(), 9);
  return RFM_OK;
}

uint32_t rfm_uecp_stuff_frame(const uint8_t* frame, uint32_t len, uint8_t* out, uint32_t cap)
{
  std::vector<uint8_t> v;
  rfm::UecpStuffFrame(frame, len, v);
  if (out && v.size() <= cap)
    memcpy(out, v.data(), v.size());
  return (uint32_t)v.size();
}

} // extern "C"

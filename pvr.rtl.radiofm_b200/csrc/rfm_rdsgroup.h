// rfm_rdsgroup.h -- RDS group -> UECP frame formatter on the host.
// Integer-only consumer of the groups the block synchroniser delivers (rfm_rdssync.h): what cRDSGroupDecoder does
// between cRDSRxSignalProcessor and cRadioReceiver::AddUECPDataFrame (RDSGroupDecoder.cpp:136-1001,
// RadioReceiver.cpp:387-414).  11.4 groups/s per stream, branchy, byte-exact by construction -- it stays on the host.
#pragma once

#include <stdint.h>

#include <vector>

namespace rfm
{

// The three calls the reference makes on its cRadioReceiver from inside DecodeRDS (RadioReceiver.h:77,80,115).
struct RdsGroupSink
{
  void* user = nullptr;
  int (*add_uecp_frame)(void* user, const uint8_t* frame, uint32_t len) = nullptr; // AddUECPDataFrame
  int (*set_channel_name)(void* user, const char* name) = nullptr;                 // SetChannelName (nonzero: accepted)
  int (*is_setting_active)(void* user) = nullptr;                                  // IsSettingActive
};

// cRadioReceiver::AddUECPDataFrame's transport framing (RadioReceiver.cpp:387-414): 0xFE, the frame with every byte
// >= 0xFD replaced by (0xFD, (byte & 3) - 1), 0xFF.  Appends to `out`; returns the bytes appended.
size_t UecpStuffFrame(const uint8_t* frame, uint32_t len, std::vector<uint8_t>& out);

class RdsGroupDecoder
{
public:
  explicit RdsGroupDecoder(const RdsGroupSink* sink = nullptr);
  void SetSink(const RdsGroupSink* sink);
  void Reset();                          // cRDSGroupDecoder::Reset, RDSGroupDecoder.cpp:140-164
  void Decode(const uint16_t block[4]);  // cRDSGroupDecoder::DecodeRDS, RDSGroupDecoder.cpp:166-272
  // without an add_uecp_frame callback the frames collect here, framed as the add-on's PID-2 byte stream
  // (frames arriving while more than 16384 bytes are pending are dropped, RadioReceiver.cpp:389-390)
  std::vector<uint8_t>& Pending() { return m_out; }
  const char* ChannelName() const { return m_name; } // last name SetChannelName accepted (when no callback is set)

private:
  // MEC values used below (UECP, SPB 490)
  enum : uint8_t
  {
    MEC_PI = 0x01, MEC_PS = 0x02, MEC_TA_TP = 0x03, MEC_DI = 0x04, MEC_MS = 0x05, MEC_PIN = 0x06, MEC_PTY = 0x07,
    MEC_RT = 0x0A, MEC_RTC = 0x0D, MEC_SLOW_LABEL = 0x1A, MEC_TMC = 0x30, MEC_PTYN = 0x3A, MEC_ODA_CONF = 0x40,
    MEC_ODA_DATA = 0x46
  };
  enum : int { AID_RTPLUS = 0x4bd7, AID_TFC = 0xcd46 };

  void Begin();                 // ClearUECPFrame
  void Put(uint8_t v);          // AddStuffingValue
  void Send();                  // SendUECPFrame
  bool SettingActive() const;
  bool NameAccepted(const char* name);

  void OnPI(uint16_t pi);
  void OnPTY(int pty);
  void Type0(const uint16_t* b);
  void Type1(const uint16_t* b, bool version_b);
  void Type2(const uint16_t* b, bool version_b);
  void Type3A(const uint16_t* b);
  void Type4A(const uint16_t* b);
  void Type8A(const uint16_t* b);
  void Type10A(const uint16_t* b);
  void Oda(const uint16_t* b, int aid);

  RdsGroupSink m_sink;
  std::vector<uint8_t> m_out;
  char m_name[9] = {0};

  // ---- fields the reference's constructor and Reset() never write (RDSGroupDecoder.cpp:136-164 against the uses at
  // :176,322,631,817,955): defined here as zero, which is what the reference has when its object lives in zeroed
  // storage -- the state the parity fixtures are generated in (oracle/ref_uecp_harness.cpp)
  uint8_t m_seq = 0;          // m_UECPDataFrameSeqCnt
  int m_pty = 0;              // m_PTY
  int m_di_seen = 0;          // m_DI_Finished
  int m_rt_ab = 0;            // m_RadioText_ABFlag
  bool m_ptyn_ab = false;     // m_PTYN_ABFlag
  char m_ps_work[9] = {0};    // the function-static ps_text of Decode_Type0___PS_DI_MS (:311), one per decoder here:
                              // a process-wide static would mix the streams of a batch

  // ---- fields Reset() initialises
  uint16_t m_pi = 0, m_pin = 0;
  uint32_t m_rt_segments = 0;
  int m_rt_count = 0;
  bool m_rt_first = false, m_rtplus_ready = false;
  uint8_t m_di = 0, m_di_prev = 0, m_ms = 0, m_ms_prev = 0;
  int m_ta_tp = 0, m_ptyn_set = 0, m_ps_set = 0;
  char m_ps[9] = {0}, m_ptyn[9] = {0};
  char m_rt[66] = {0};
  int m_oda[32] = {0};

  // frame under construction: ADD, ADD, SQC, MFL, message, CRC
  uint8_t m_frame[263] = {0};
  int m_len = 0;
};

} // namespace rfm

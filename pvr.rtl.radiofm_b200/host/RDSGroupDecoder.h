// host/RDSGroupDecoder.h -- cRDSGroupDecoder with the reference's signatures (RDSGroupDecoder.h:24-31) over the C ABI
// (rfm_rdsgroup_*): RDS groups in, UECP frames out through the receiver's AddUECPDataFrame / SetChannelName /
// IsSettingActive (RadioReceiver.h:77,80,115), called synchronously from inside DecodeRDS as in the reference.
// cRadioReceiver only has to be a complete type by the END of the translation unit that constructs a decoder (the
// reference includes RadioReceiver.h from its .cpp files, not from this header): the three forwarding functions are
// templates, instantiated where the constructor is used.
#pragma once

#include <stdint.h>

#include <stdexcept>
#include <string>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"

class cRDSGroupDecoder
{
public:
  cRDSGroupDecoder(cRadioReceiver* proc) : m_RadioProc(proc) { Bind<cRadioReceiver>(); }
  virtual ~cRDSGroupDecoder() { rfm_rdsgroup_destroy(m_g); }
  cRDSGroupDecoder(const cRDSGroupDecoder&) = delete;
  cRDSGroupDecoder& operator=(const cRDSGroupDecoder&) = delete;

  void DecodeRDS(uint16_t* blockData) { rfm_rdsgroup_decode(m_g, blockData, 1); } // RDSGroupDecoder.cpp:166-272
  void Reset() { rfm_rdsgroup_reset(m_g); }                                       // RDSGroupDecoder.cpp:140-164

private:
  template <class Receiver>
  void Bind()
  {
    rfm_rdsgroup_callbacks cb;
    cb.user = m_RadioProc;
    cb.add_uecp_frame = [](void* u, const uint8_t* frame, uint32_t len) -> int {
      return static_cast<Receiver*>(u)->AddUECPDataFrame(const_cast<uint8_t*>(frame), len) ? 1 : 0;
    };
    cb.set_channel_name = [](void* u, const char* name) -> int {
      return static_cast<Receiver*>(u)->SetChannelName(std::string(name)) ? 1 : 0;
    };
    cb.is_setting_active = [](void* u) -> int { return static_cast<Receiver*>(u)->IsSettingActive() ? 1 : 0; };
    if (!m_RadioProc) // the oracle builds its decoders with proc == nullptr: frames collect inside the object
      cb.add_uecp_frame = nullptr, cb.set_channel_name = nullptr, cb.is_setting_active = nullptr;
    if (rfm_rdsgroup_create(&cb, &m_g) != RFM_OK)
      throw std::runtime_error(std::string("cRDSGroupDecoder (B200): ") + rfm_last_error());
  }

  cRadioReceiver* m_RadioProc;
  rfm_rdsgroup* m_g = nullptr;
};

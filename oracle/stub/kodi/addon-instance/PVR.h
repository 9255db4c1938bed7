// Minimal stand-in for <kodi/addon-instance/PVR.h>: the names RadioReceiver.h / RadioReceiver.cpp use, with just
// enough behaviour for the reference's cRadioReceiver to run its receive path (OpenLiveStream -> DemuxRead) inside
// oracle/ref_addon_harness.cpp: value classes with the setters / getters the add-on calls, a demux packet allocator
// on the heap, a codec lookup that knows "pcm_f32le" and "rds".  TEST INFRASTRUCTURE ONLY; written from the names the
// reference uses, not from the dev-kit.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../AddonBase.h"
#include "../General.h"

enum PVR_ERROR { PVR_ERROR_NO_ERROR = 0, PVR_ERROR_UNKNOWN = -1, PVR_ERROR_NOT_IMPLEMENTED = -2,
                 PVR_ERROR_REJECTED = -5, PVR_ERROR_INVALID_PARAMETERS = -7, PVR_ERROR_FAILED = -9 };
enum PVR_CODEC_TYPE { PVR_CODEC_TYPE_UNKNOWN = -1, PVR_CODEC_TYPE_VIDEO = 0, PVR_CODEC_TYPE_AUDIO, PVR_CODEC_TYPE_DATA,
                      PVR_CODEC_TYPE_SUBTITLE, PVR_CODEC_TYPE_RDS };
#define DEMUX_SPECIALID_STREAMINFO -10
#define DEMUX_SPECIALID_STREAMCHANGE -11
#define STREAM_TIME_BASE 1000000
struct DEMUX_PACKET
{
  unsigned char* pData;
  int iSize;
  int iStreamId;
  int64_t demuxerId;
  int iGroupId;
  void* pSideData;
  int iSideDataElems;
  double pts;
  double dts;
  double duration;
  int dispTime;
  bool recoveryPoint;
  void* cryptoInfo;
};

namespace kodi
{
namespace addon
{
class PVRCapabilities
{
public:
#define RFM_STUB_FLAG(name) \
  void Set##name(bool v) { m_##name = v; } \
  bool Get##name() const { return m_##name; } \
  bool m_##name = false;
  RFM_STUB_FLAG(SupportsEPG) RFM_STUB_FLAG(SupportsRecordings) RFM_STUB_FLAG(SupportsRecordingEdl)
  RFM_STUB_FLAG(SupportsRecordingsUndelete) RFM_STUB_FLAG(SupportsTimers) RFM_STUB_FLAG(SupportsTV)
  RFM_STUB_FLAG(SupportsRadio) RFM_STUB_FLAG(SupportsChannelGroups) RFM_STUB_FLAG(HandlesInputStream)
  RFM_STUB_FLAG(HandlesDemuxing) RFM_STUB_FLAG(SupportsChannelScan) RFM_STUB_FLAG(SupportsChannelSettings)
#undef RFM_STUB_FLAG
};

class PVRChannel
{
public:
  void SetUniqueId(unsigned int v) { m_uid = v; }
  unsigned int GetUniqueId() const { return m_uid; }
  void SetIsRadio(bool v) { m_radio = v; }
  void SetChannelNumber(unsigned int v) { m_number = v; }
  int GetChannelNumber() const { return (int)m_number; }
  void SetChannelName(const std::string& v) { m_name = v; }
  std::string GetChannelName() const { return m_name; }
  void SetIconPath(const std::string& v) { m_icon = v; }
  std::string GetIconPath() const { return m_icon; }
  void SetIsHidden(bool v) { m_hidden = v; }

private:
  unsigned int m_uid = 0, m_number = 0;
  bool m_radio = false, m_hidden = false;
  std::string m_name, m_icon;
};

class PVRChannelsResultSet
{
public:
  void Add(const PVRChannel& c) { channels.push_back(c); }
  std::vector<PVRChannel> channels;
};

class PVRSignalStatus
{
public:
  void SetAdapterName(const std::string& v) { adapter_name = v; }
  void SetAdapterStatus(const std::string& v) { adapter_status = v; }
  void SetProviderName(const std::string& v) { provider_name = v; }
  void SetSignal(int v) { signal = v; }
  void SetSNR(int v) { snr = v; }
  std::string adapter_name, adapter_status, provider_name;
  int signal = 0, snr = 0;
};

class PVRCodec
{
public:
  void SetCodecType(PVR_CODEC_TYPE t) { m_type = t; }
  PVR_CODEC_TYPE GetCodecType() const { return m_type; }
  void SetCodecId(unsigned int id) { m_id = id; }
  unsigned int GetCodecId() const { return m_id; }

private:
  PVR_CODEC_TYPE m_type = PVR_CODEC_TYPE_UNKNOWN;
  unsigned int m_id = 0;
};

class PVRStreamProperties
{
public:
#define RFM_STUB_PROP(type, name) \
  void Set##name(type v) { m_##name = v; } \
  type Get##name() const { return m_##name; } \
  type m_##name = (type)0;
  RFM_STUB_PROP(unsigned int, PID) RFM_STUB_PROP(PVR_CODEC_TYPE, CodecType) RFM_STUB_PROP(unsigned int, CodecId)
  RFM_STUB_PROP(int, Channels) RFM_STUB_PROP(int, SampleRate) RFM_STUB_PROP(int, BitsPerSample) RFM_STUB_PROP(int, BitRate)
#undef RFM_STUB_PROP
};

class CInstancePVRClient
{
public:
  virtual ~CInstancePVRClient() = default;
  virtual PVR_ERROR GetCapabilities(PVRCapabilities&) = 0;
  virtual PVR_ERROR GetBackendName(std::string&) = 0;
  virtual PVR_ERROR GetBackendVersion(std::string&) = 0;
  virtual PVR_ERROR GetConnectionString(std::string&) = 0;
  virtual PVR_ERROR GetChannelsAmount(int&) = 0;
  virtual PVR_ERROR GetChannels(bool, PVRChannelsResultSet&) = 0;
  virtual PVR_ERROR DeleteChannel(const PVRChannel&) = 0;
  virtual PVR_ERROR RenameChannel(const PVRChannel&) = 0;
  virtual PVR_ERROR OpenDialogChannelSettings(const PVRChannel&) = 0;
  virtual PVR_ERROR OpenDialogChannelAdd(const PVRChannel&) = 0;
  virtual PVR_ERROR GetSignalStatus(int, PVRSignalStatus&) = 0;
  virtual bool OpenLiveStream(const PVRChannel&) = 0;
  virtual void CloseLiveStream() = 0;
  virtual PVR_ERROR GetStreamProperties(std::vector<PVRStreamProperties>&) = 0;
  virtual DEMUX_PACKET* DemuxRead() = 0;
  virtual void DemuxAbort() = 0;

  // Kodi-side services the add-on calls
  DEMUX_PACKET* AllocateDemuxPacket(int size)
  {
    DEMUX_PACKET* p = (DEMUX_PACKET*)calloc(1, sizeof(DEMUX_PACKET));
    if (p && size > 0)
      p->pData = (unsigned char*)malloc((size_t)size);
    return p;
  }
  void FreeDemuxPacket(DEMUX_PACKET* p)
  {
    if (p)
      free(p->pData);
    free(p);
  }
  PVRCodec GetCodecByName(const std::string& name) const
  {
    PVRCodec c;
    if (name == "pcm_f32le")
    {
      c.SetCodecType(PVR_CODEC_TYPE_AUDIO);
      c.SetCodecId(65557);
    }
    else if (name == "rds")
    {
      c.SetCodecType(PVR_CODEC_TYPE_RDS);
      c.SetCodecId(100000);
    }
    return c;
  }
  void TriggerChannelUpdate() { ++channel_updates; }
  int channel_updates = 0;
};
} // namespace addon
} // namespace kodi

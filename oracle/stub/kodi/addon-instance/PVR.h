// Minimal stand-in for <kodi/addon-instance/PVR.h>: just enough declarations for
// the reference's RadioReceiver.h to parse (its methods are never defined or called:
// the oracle builds cFmDecoder with proc == nullptr).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <string>
#include <vector>
#include "../AddonBase.h"
#include "../General.h"
enum PVR_ERROR { PVR_ERROR_NO_ERROR = 0, PVR_ERROR_UNKNOWN = -1, PVR_ERROR_NOT_IMPLEMENTED = -2,
                 PVR_ERROR_REJECTED = -5, PVR_ERROR_INVALID_PARAMETERS = -7, PVR_ERROR_FAILED = -9 };
struct DEMUX_PACKET;
namespace kodi { namespace addon {
class CAddonBase { public: virtual ~CAddonBase() = default; };
class PVRCapabilities {};
class PVRChannelsResultSet {};
class PVRSignalStatus {};
class PVRStreamProperties {};
class PVRChannel { public: int GetChannelNumber() const { return 0; } };
class CInstancePVRClient {
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
};
}}

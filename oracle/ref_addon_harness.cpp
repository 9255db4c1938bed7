// oracle/ref_addon_harness.cpp -- C-ABI harness around the UNMODIFIED reference add-on receive path:
// cRadioReceiver (RadioReceiver.cpp: OpenLiveStream, the IQ block queue, DemuxRead, AddUECPDataFrame, the signal
// status) + cRtlSdrSource (RTL_SDR_Source.cpp: Configure, the reader thread, the u8 -> float callback) + cFmDecoder,
// cRDSRxSignalProcessor, cRDSGroupDecoder and the DSP classes -- every reference source file of the path, compiled
// where it lies under /root/reference/src (never copied) by oracle/Makefile (`make addon`) into
// oracle/_ref/libradiofm_ref_addon.so.
//
// TEST INFRASTRUCTURE ONLY (same rules as ref_harness.cpp): nothing in the product may include, link or call it.
//
// What stands in for the outside world (all of it written here or in oracle/stub/, from the names the reference uses):
//   * the Kodi dev-kit: oracle/stub/kodi/... (value classes, AllocateDemuxPacket on the heap, GetCodecByName);
//   * TinyXML: oracle/stub/tinyxml.h (in-memory tree and "files"), so the constructor's LoadChannelData /
//     SaveChannelData run as written;
//   * the settings dialog cChannelSettings (ChannelSettings.cpp is GUI code): its members are defined below as no-ops;
//   * librtlsdr: the rtlsdr_* functions below -- a "device" that hands out, in buffers of the requested length, the
//     bytes the test feeds it, and logs every configuration call.
// This pins the sequencing rows the Python restatement (oracle/demux_port.py) could only restate in round 1:
// RadioReceiver.cpp:387-414 (UECP transport framing), :420-460 (block queue), :462-542 (DemuxRead order, pts, duration,
// audio level), :544-582 (signal status), RTL_SDR_Source.cpp:64-146 (Configure, block-length rule), :196-213 (callback).
// (The reader thread's restart branch, RTL_SDR_Source.cpp:226-241, assigns a new std::thread to the joinable m_thread
// from inside that thread -- std::terminate; it is not exercised.)
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

// every std header the reference headers pull in comes first: "#define private public" must not reach them
#define private public
#define protected public
#include "RadioReceiver.h"
#include "ChannelSettings.h"
#include "FmDecode.h"
#undef private
#undef protected
#include <rtl-sdr.h>
#include <tinyxml.h>

// ---- a zeroed heap ---------------------------------------------------------------------------------------------
// cRDSGroupDecoder's constructor and Reset() leave m_PTY, m_DI_Finished, m_RadioText_ABFlag, m_PTYN_ABFlag and
// m_UECPDataFrameSeqCnt uninitialised; OpenLiveStream allocates the decoder with plain `new`, so the first frames (and
// the sequence counter of all of them) depend on what the heap held before.  Every allocation made inside this library
// (-Wl,-Bsymbolic-functions binds the reference objects' operator new to this one) is zeroed: the members start at
// zero, the state ref_uecp_harness.cpp pins with its zeroed placement storage.
void* operator new(size_t n)
{
  void* p = calloc(1, n ? n : 1);
  if (!p)
    throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, size_t) noexcept { free(p); }
void operator delete[](void* p, size_t) noexcept { free(p); }

// ---- the settings dialog: never opened -------------------------------------------------------------------------
cChannelSettings::cChannelSettings() : kodi::gui::CWindow("ChannelTuner.xml", "skin.estuary", true, true) {}
cChannelSettings::~cChannelSettings() {}
PVR_ERROR cChannelSettings::Open(const kodi::addon::PVRChannel&, cRadioReceiver*, bool) { return PVR_ERROR_NO_ERROR; }
void cChannelSettings::UpdateName(std::string name) { m_Name = name; }
bool cChannelSettings::OnClick(int) { return false; }
bool cChannelSettings::OnFocus(int) { return false; }
bool cChannelSettings::OnInit() { return false; }
bool cChannelSettings::OnAction(ADDON_ACTION) { return false; }
void cChannelSettings::Process() {}
void cChannelSettings::UpdateFreq(uint32_t) {}

// ---- the "device" ------------------------------------------------------------------------------------------------
struct rtlsdr_dev
{
  uint32_t rate = 0, freq = 0;
  int gain = 0, gain_mode = 0, agc = 0;
};

namespace
{
struct FakeDevice
{
  std::mutex mu;
  std::condition_variable cv;
  std::vector<uint8_t> bytes; // fed, not yet handed out
  size_t head = 0;
  bool cancel = false;
  bool in_read = false;    // the reader thread is inside rtlsdr_read_async
  bool delivering = false; // ... and inside the callback
  uint32_t buf_len = 0;
  bool short_read_pending = false; // hand out one half-length buffer after the next full one
  std::string log;
  void Log(const char* fmt, ...)
  {
    char line[128];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(line, sizeof(line), fmt, ap);
    va_end(ap);
    log += line;
    log += '\n';
  }
};
FakeDevice g_dev;
const int kGains[] = {0, 9, 14, 27, 37, 77, 87, 125, 144, 157, 166, 197, 207, 229, 254, 280, 297, 328, 338, 364, 372, 386, 402, 421, 434, 439, 445, 480, 496};
} // namespace

extern "C" {
uint32_t rtlsdr_get_device_count(void) { return 1; }
const char* rtlsdr_get_device_name(uint32_t index) { return index == 0 ? "Fake RTL2838 (fed by the test)" : ""; }
int rtlsdr_open(rtlsdr_dev_t** dev, uint32_t index)
{
  if (index != 0)
    return -1;
  *dev = new rtlsdr_dev();
  std::lock_guard<std::mutex> l(g_dev.mu);
  g_dev.Log("open %u", index);
  return 0;
}
int rtlsdr_close(rtlsdr_dev_t* dev)
{
  std::lock_guard<std::mutex> l(g_dev.mu);
  g_dev.Log("close");
  delete dev;
  return 0;
}
#define RFM_FAKE_SET(fn, field, type, fmt)              \
  int fn(rtlsdr_dev_t* dev, type v)                     \
  {                                                     \
    dev->field = v;                                     \
    std::lock_guard<std::mutex> l(g_dev.mu);            \
    g_dev.Log(fmt, v);                                  \
    return 0;                                           \
  }
RFM_FAKE_SET(rtlsdr_set_sample_rate, rate, uint32_t, "sample_rate %u")
RFM_FAKE_SET(rtlsdr_set_center_freq, freq, uint32_t, "center_freq %u")
RFM_FAKE_SET(rtlsdr_set_tuner_gain_mode, gain_mode, int, "gain_mode %d")
RFM_FAKE_SET(rtlsdr_set_tuner_gain, gain, int, "gain %d")
RFM_FAKE_SET(rtlsdr_set_agc_mode, agc, int, "agc %d")
uint32_t rtlsdr_get_sample_rate(rtlsdr_dev_t* dev) { return dev->rate; }
uint32_t rtlsdr_get_center_freq(rtlsdr_dev_t* dev) { return dev->freq; }
int rtlsdr_get_tuner_gain(rtlsdr_dev_t* dev) { return dev->gain; }
int rtlsdr_get_tuner_gains(rtlsdr_dev_t*, int* gains)
{
  const int n = (int)(sizeof(kGains) / sizeof(kGains[0]));
  if (gains)
    memcpy(gains, kGains, sizeof(kGains));
  return n;
}
int rtlsdr_reset_buffer(rtlsdr_dev_t*)
{
  std::lock_guard<std::mutex> l(g_dev.mu);
  g_dev.cancel = false;
  g_dev.Log("reset_buffer");
  return 0;
}
int rtlsdr_cancel_async(rtlsdr_dev_t*)
{
  std::lock_guard<std::mutex> l(g_dev.mu);
  g_dev.cancel = true;
  g_dev.cv.notify_all();
  return 0;
}
int rtlsdr_read_async(rtlsdr_dev_t*, rtlsdr_read_async_cb_t cb, void* ctx, uint32_t buf_num, uint32_t buf_len)
{
  std::unique_lock<std::mutex> l(g_dev.mu);
  g_dev.Log("read_async %u %u", buf_num, buf_len);
  g_dev.buf_len = buf_len;
  g_dev.in_read = true;
  g_dev.cv.notify_all();
  std::vector<uint8_t> buf(buf_len);
  while (!g_dev.cancel)
  {
    if (g_dev.bytes.size() - g_dev.head < buf_len)
    {
      g_dev.cv.wait_for(l, std::chrono::milliseconds(20));
      continue;
    }
    memcpy(buf.data(), g_dev.bytes.data() + g_dev.head, buf_len);
    g_dev.head += buf_len;
    const bool short_read = g_dev.short_read_pending;
    g_dev.short_read_pending = false;
    g_dev.delivering = true;
    l.unlock();
    cb(buf.data(), buf_len, ctx);
    if (short_read)
      cb(buf.data(), buf_len / 2, ctx); // must be dropped (RTL_SDR_Source.cpp:200-204)
    l.lock();
    g_dev.delivering = false;
    g_dev.cv.notify_all();
  }
  g_dev.in_read = false;
  g_dev.cv.notify_all();
  return 0;
}
} // extern "C"

// ---- the harness ---------------------------------------------------------------------------------------------
namespace
{
struct RefAddon
{
  cRadioReceiver* radio = nullptr;
  kodi::addon::PVRChannel channel;
};

// the reader thread has handed the demux queue every whole buffer the test has fed so far
bool WaitSourceIdle()
{
  std::unique_lock<std::mutex> l(g_dev.mu);
  return g_dev.cv.wait_for(l, std::chrono::seconds(20), [] {
    return g_dev.in_read && !g_dev.delivering && g_dev.bytes.size() - g_dev.head < g_dev.buf_len;
  });
}
} // namespace

extern "C" {
__attribute__((visibility("default"))) void* refaddon_create(float channel_freq_hz)
{
  {
    std::lock_guard<std::mutex> l(g_dev.mu);
    g_dev.bytes.clear();
    g_dev.head = 0;
    g_dev.cancel = false;
    g_dev.in_read = g_dev.delivering = false;
    g_dev.buf_len = 0;
    g_dev.short_read_pending = false;
    g_dev.log.clear();
  }
  TiXmlDocument::Store().clear();
  RefAddon* a = new RefAddon();
  a->radio = new cRadioReceiver(); // LoadChannelData(true): no settings file -> SaveChannelData writes the first one
  FMRadioChannel ch;
  ch.iUniqueId = a->radio->CreateNewUniqueId();
  if (ch.iUniqueId == 0) // m_UniqueIdNextNew is only set when a settings file was loaded; 0 is "no id" for the loader
    ch.iUniqueId = a->radio->CreateNewUniqueId();
  ch.fChannelFreq = channel_freq_hz;
  ch.strChannelName = "-";
  a->radio->GetChannelData()->push_back(ch);
  a->radio->SaveChannelData();
  a->channel.SetUniqueId(ch.iUniqueId);
  a->channel.SetChannelName("test");
  return a;
}

// the settings round trip: a second receiver loads what the first one saved; returns its channel count and the
// frequency of channel 0
__attribute__((visibility("default"))) int refaddon_reload_channels(void*, float* freq0)
{
  cRadioReceiver other;
  if (!other.GetChannelData()->empty() && freq0)
    *freq0 = other.GetChannelData()->front().fChannelFreq;
  return (int)other.GetChannelData()->size();
}

__attribute__((visibility("default"))) int refaddon_open(void* h)
{
  RefAddon* a = static_cast<RefAddon*>(h);
  if (!a->radio->OpenLiveStream(a->channel))
    return 0;
  std::unique_lock<std::mutex> l(g_dev.mu); // the reader thread has reached rtlsdr_read_async
  g_dev.cv.wait_for(l, std::chrono::seconds(10), [] { return g_dev.in_read; });
  return 1;
}

// what OpenLiveStream derived: IF rate, tuning offset handed to cFmDecoder, down-sampling factor, source block length
__attribute__((visibility("default"))) void refaddon_params(void* h, double* if_rate, double* tuning_offset,
                                                            unsigned* downsample, unsigned* block_length, double* tuner_freq)
{
  RefAddon* a = static_cast<RefAddon*>(h);
  *if_rate = a->radio->m_IfRate;
  *tuning_offset = a->radio->m_activeChannelFrequency - a->radio->m_activeTunerFreq;
  *downsample = (unsigned)std::max(1, int(a->radio->m_IfRate / 215.0e3));
  *block_length = (unsigned)a->radio->m_RtlSdrReceiver.m_BlockLength;
  *tuner_freq = a->radio->m_activeTunerFreq;
}

__attribute__((visibility("default"))) void refaddon_feed(void*, const uint8_t* iq, size_t nbytes, int short_read_after_next)
{
  std::lock_guard<std::mutex> l(g_dev.mu);
  g_dev.bytes.insert(g_dev.bytes.end(), iq, iq + nbytes);
  if (short_read_after_next)
    g_dev.short_read_pending = true;
  g_dev.cv.notify_all();
}

// One DemuxRead call.  Returns 1 and the packet; 0 if the call would wait for the source (stream change and UECP
// buffer empty, every fed block decoded); -1 if DemuxRead returned nullptr; -2 if `cap` is too small.
__attribute__((visibility("default"))) int refaddon_read(void* h, int* stream_id, int* size, double* pts, double* duration,
                                                         uint8_t* data, int cap)
{
  RefAddon* a = static_cast<RefAddon*>(h);
  cRadioReceiver* r = a->radio;
  if (!r->m_StreamChange && r->m_UECPOutputBuffer.empty())
  {
    if (!WaitSourceIdle())
      return -3;
    if (r->SourceQueuedSamples() == 0)
      return 0;
  }
  DEMUX_PACKET* p = r->DemuxRead();
  if (!p)
    return -1;
  *stream_id = p->iStreamId;
  *size = p->iSize;
  *pts = p->pts;
  *duration = p->duration;
  int rc = 1;
  if (p->iSize > cap)
    rc = -2;
  else if (p->iSize > 0)
    memcpy(data, p->pData, (size_t)p->iSize);
  r->FreeDemuxPacket(p);
  return rc;
}

__attribute__((visibility("default"))) int refaddon_signal(void* h, float* if_level_db, float* audio_level_db, int* stereo,
                                                           int* signal, int* snr, char* status, int status_cap)
{
  RefAddon* a = static_cast<RefAddon*>(h);
  bool st = false;
  if (!a->radio->GetSignalStatus(*if_level_db, *audio_level_db, st))
    return 0;
  *stereo = st ? 1 : 0;
  kodi::addon::PVRSignalStatus s;
  if (a->radio->GetSignalStatus(0, s) != PVR_ERROR_NO_ERROR)
    return 0;
  *signal = s.signal;
  *snr = s.snr;
  snprintf(status, (size_t)status_cap, "%s", s.adapter_status.c_str());
  return 1;
}

__attribute__((visibility("default"))) float refaddon_audio_level(void* h) { return static_cast<RefAddon*>(h)->radio->m_AudioLevel; }
__attribute__((visibility("default"))) unsigned long long refaddon_queued_samples(void* h)
{
  return static_cast<RefAddon*>(h)->radio->SourceQueuedSamples();
}
__attribute__((visibility("default"))) void refaddon_set_stream_change(void* h) { static_cast<RefAddon*>(h)->radio->SetStreamChange(); }

// stream properties as OpenLiveStream published them: pid, codec type, channels, sample rate, bits, bit rate x n
__attribute__((visibility("default"))) int refaddon_stream_properties(void* h, int* out, int cap_streams)
{
  std::vector<kodi::addon::PVRStreamProperties> props;
  if (static_cast<RefAddon*>(h)->radio->GetStreamProperties(props) != PVR_ERROR_NO_ERROR)
    return 0;
  int n = 0;
  for (const auto& p : props)
  {
    if (n >= cap_streams)
      break;
    int* o = out + 6 * n++;
    o[0] = (int)p.GetPID(); o[1] = (int)p.GetCodecType(); o[2] = p.GetChannels();
    o[3] = p.GetSampleRate(); o[4] = p.GetBitsPerSample(); o[5] = p.GetBitRate();
  }
  return n;
}

__attribute__((visibility("default"))) int refaddon_device_log(void*, char* out, int cap)
{
  std::lock_guard<std::mutex> l(g_dev.mu);
  snprintf(out, (size_t)cap, "%s", g_dev.log.c_str());
  return (int)g_dev.log.size();
}

__attribute__((visibility("default"))) void refaddon_close(void* h) { static_cast<RefAddon*>(h)->radio->CloseLiveStream(); }

__attribute__((visibility("default"))) void refaddon_destroy(void* h)
{
  RefAddon* a = static_cast<RefAddon*>(h);
  delete a->radio;
  delete a;
}
} // extern "C"

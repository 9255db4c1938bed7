// rfm_source.cpp -- cRtlSdrSource (RTL_SDR_Source.cpp:27-256): the live source adapter in front of the
// demux loop.  librtlsdr is bound at RUN time (dlopen): the library neither links nor ships it, and a host without it
// gets RFM_ERR_UNSUPPORTED from rfm_rtlsdr_open, nothing else changes.  What the reference does with the device --
// Open / Configure (sample rate, centre frequency, gain mode and gain, AGC, block length rule, reset_buffer), the
// reader thread around rtlsdr_read_async with its five restart attempts, Close (cancel_async + join), the getters --
// is reproduced call for call; the one deliberate difference is the callback: the raw u8 block is queued as it is
// (rfm_demux_source_cb) and converted on the device, where the reference expands it to complex<float> on the host
// (RTL_SDR_Source.cpp:207-211).
#include <dlfcn.h>
#include <limits.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <string>
#include <thread>

#include "../../include/radiofm_b200.h"

namespace
{
constexpr unsigned kRestartTries = 5; // DEVICE_RESTART_TRIES, Definitions.h:13

typedef void (*read_cb_t)(unsigned char*, uint32_t, void*);
struct RtlApi
{
  void* lib = nullptr;
  uint32_t (*get_device_count)() = nullptr;
  const char* (*get_device_name)(uint32_t) = nullptr;
  int (*open)(void**, uint32_t) = nullptr;
  int (*close)(void*) = nullptr;
  int (*set_sample_rate)(void*, uint32_t) = nullptr;
  uint32_t (*get_sample_rate)(void*) = nullptr;
  int (*set_center_freq)(void*, uint32_t) = nullptr;
  uint32_t (*get_center_freq)(void*) = nullptr;
  int (*set_tuner_gain_mode)(void*, int) = nullptr;
  int (*set_tuner_gain)(void*, int) = nullptr;
  int (*get_tuner_gain)(void*) = nullptr;
  int (*set_agc_mode)(void*, int) = nullptr;
  int (*reset_buffer)(void*) = nullptr;
  int (*read_async)(void*, read_cb_t, void*, uint32_t, uint32_t) = nullptr;
  int (*cancel_async)(void*) = nullptr;
};

bool LoadApi(const char* library, RtlApi* a, std::string* why)
{
  const char* names[] = {library, "librtlsdr.so.0", "librtlsdr.so", "librtlsdr.so.2"};
  for (size_t i = library ? 0 : 1; i < (library ? 1u : 4u) && !a->lib; ++i)
    a->lib = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
  if (!a->lib)
  {
    const char* e = dlerror();
    *why = std::string("librtlsdr not found (") + (e ? e : "dlopen failed") + ")";
    return false;
  }
#define RFM_SYM(field, name)                                                       \
  *reinterpret_cast<void**>(&a->field) = dlsym(a->lib, name);                      \
  if (!a->field)                                                                   \
  {                                                                                \
    *why = std::string("librtlsdr lacks ") + name;                                 \
    return false;                                                                  \
  }
  RFM_SYM(get_device_count, "rtlsdr_get_device_count")
  RFM_SYM(get_device_name, "rtlsdr_get_device_name")
  RFM_SYM(open, "rtlsdr_open")
  RFM_SYM(close, "rtlsdr_close")
  RFM_SYM(set_sample_rate, "rtlsdr_set_sample_rate")
  RFM_SYM(get_sample_rate, "rtlsdr_get_sample_rate")
  RFM_SYM(set_center_freq, "rtlsdr_set_center_freq")
  RFM_SYM(get_center_freq, "rtlsdr_get_center_freq")
  RFM_SYM(set_tuner_gain_mode, "rtlsdr_set_tuner_gain_mode")
  RFM_SYM(set_tuner_gain, "rtlsdr_set_tuner_gain")
  RFM_SYM(get_tuner_gain, "rtlsdr_get_tuner_gain")
  RFM_SYM(set_agc_mode, "rtlsdr_set_agc_mode")
  RFM_SYM(reset_buffer, "rtlsdr_reset_buffer")
  RFM_SYM(read_async, "rtlsdr_read_async")
  RFM_SYM(cancel_async, "rtlsdr_cancel_async")
#undef RFM_SYM
  return true;
}
} // namespace

struct rfm_rtlsdr
{
  RtlApi api;
  rfm_demux* demux = nullptr;
  void* dev = nullptr;
  std::thread thread;
  std::atomic<bool> running{false}, cancelled{false};
  uint32_t sample_rate = 0, frequency = 0, block = 0;
  int tuner_gain = INT_MIN, agc = 0;
  std::atomic<unsigned> restarts{0};
  std::string error;
};

namespace
{
// cRtlSdrSource::Configure without the thread start, RTL_SDR_Source.cpp:66-137
bool Configure(rfm_rtlsdr* s)
{
  const RtlApi& a = s->api;
  if (!s->dev)
    return false;
  if (a.set_sample_rate(s->dev, s->sample_rate) < 0)
    return s->error = "rtlsdr_set_sample_rate failed", false;
  if (a.set_center_freq(s->dev, s->frequency) < 0)
    return s->error = "rtlsdr_set_center_freq failed", false;
  if (s->tuner_gain == INT_MIN)
  {
    if (a.set_tuner_gain_mode(s->dev, 0) < 0)
      return s->error = "rtlsdr_set_tuner_gain_mode could not set automatic gain", false;
  }
  else
  {
    if (a.set_tuner_gain_mode(s->dev, 1) < 0)
      return s->error = "rtlsdr_set_tuner_gain_mode could not set manual gain", false;
    if (a.set_tuner_gain(s->dev, s->tuner_gain) < 0)
      return s->error = "rtlsdr_set_tuner_gain failed", false;
  }
  if (a.set_agc_mode(s->dev, s->agc) < 0)
    return s->error = "rtlsdr_set_agc_mode failed", false;
  if (a.reset_buffer(s->dev) < 0)
    return s->error = "rtlsdr_reset_buffer failed", false;
  return true;
}

// cRtlSdrSource::Process, RTL_SDR_Source.cpp:215-245
void Process(rfm_rtlsdr* s)
{
  unsigned restart_try = 0;
  bool ok = true;
  while (s->running && !s->cancelled)
  {
    const int fail = s->api.read_async(s->dev, rfm_demux_source_cb, s->demux, 15, 2 * s->block);
    if (s->cancelled)
      break;
    if (!ok || fail)
    {
      if (restart_try >= kRestartTries)
        break;
      for (int i = 0; i < 100 && !s->cancelled; ++i) // the reference sleeps one second; here in slices, so Close is prompt
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      ++restart_try;
      s->restarts = restart_try;
      ok = Configure(s);
      continue;
    }
  }
  rfm_demux_end(s->demux); // m_Proc->EndDataBuffer()
}
} // namespace

extern "C"
{

int rfm_rtlsdr_device_count(const char* library)
{
  RtlApi a;
  std::string why;
  if (!LoadApi(library, &a, &why))
    return RFM_ERR_UNSUPPORTED;
  const int n = (int)a.get_device_count();
  dlclose(a.lib);
  return n;
}

int rfm_rtlsdr_open(rfm_demux* demux, const char* library, int dev_index, rfm_rtlsdr** out)
{
  if (!demux || !out)
    return RFM_ERR_INVALID;
  *out = nullptr;
  rfm_rtlsdr* s = new rfm_rtlsdr;
  s->demux = demux;
  std::string why;
  if (!LoadApi(library, &s->api, &why))
  {
    if (s->api.lib)
      dlclose(s->api.lib);
    delete s;
    return RFM_ERR_UNSUPPORTED;
  }
  if (s->api.open(&s->dev, (uint32_t)dev_index) < 0 || !s->dev) // cRtlSdrSource::Open, RTL_SDR_Source.cpp:34-51
  {
    dlclose(s->api.lib);
    delete s;
    return RFM_ERR_NO_DEVICE;
  }
  *out = s;
  return RFM_OK;
}

int rfm_rtlsdr_configure(rfm_rtlsdr* s, uint32_t sample_rate, uint32_t frequency, int tuner_gain, int block_length,
                         int agcmode)
{
  if (!s || s->running)
    return RFM_ERR_INVALID;
  s->sample_rate = sample_rate;
  s->frequency = frequency;
  s->tuner_gain = tuner_gain;
  s->agc = agcmode ? 1 : 0;
  s->block = rfm_source_block_length(block_length < 0 ? 0u : (uint32_t)block_length); // RTL_SDR_Source.cpp:124-126
  if (rfm_demux_set_source_block_length(s->demux, s->block) != RFM_OK)
    return s->error = "block length exceeds the decoder's max_block_len", RFM_ERR_INVALID;
  if (!Configure(s))
    return RFM_ERR_INVALID;
  s->cancelled = false;
  s->running = true;
  s->thread = std::thread(Process, s);
  return RFM_OK;
}

void rfm_rtlsdr_close(rfm_rtlsdr* s) // cRtlSdrSource::Close + the destructor, RTL_SDR_Source.cpp:27-32,53-64
{
  if (!s)
    return;
  s->cancelled = true;
  if (s->dev)
    s->api.cancel_async(s->dev);
  s->running = false;
  if (s->thread.joinable())
    s->thread.join();
  if (s->dev)
    s->api.close(s->dev);
  if (s->api.lib)
    dlclose(s->api.lib);
  delete s;
}

uint32_t rfm_rtlsdr_get_sample_rate(rfm_rtlsdr* s) { return s && s->dev ? s->api.get_sample_rate(s->dev) : 0; }
uint32_t rfm_rtlsdr_get_frequency(rfm_rtlsdr* s) { return s && s->dev ? s->api.get_center_freq(s->dev) : 0; }
void rfm_rtlsdr_set_frequency(rfm_rtlsdr* s, uint32_t freq)
{
  if (s && s->dev)
    s->api.set_center_freq(s->dev, freq);
}
int rfm_rtlsdr_get_tuner_gain(rfm_rtlsdr* s) { return s && s->dev ? s->api.get_tuner_gain(s->dev) : 0; }
uint32_t rfm_rtlsdr_block_length(const rfm_rtlsdr* s) { return s ? s->block : 0; }
uint32_t rfm_rtlsdr_restarts(const rfm_rtlsdr* s) { return s ? s->restarts.load() : 0; }
const char* rfm_rtlsdr_error(const rfm_rtlsdr* s) { return s ? s->error.c_str() : ""; }

} // extern "C"

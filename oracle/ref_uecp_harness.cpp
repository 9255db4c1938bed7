// oracle/ref_uecp_harness.cpp -- C-ABI harness around the UNMODIFIED reference cRDSGroupDecoder.
//
// TEST INFRASTRUCTURE ONLY (same rules as ref_harness.cpp): nothing in the product may include, link or call it.
//
// oracle/Makefile compiles /root/reference/src/RDSGroupDecoder.cpp in place (never copied) and links it with this
// file into oracle/_ref/libradiofm_ref_uecp.so.  It is a separate library because ref_harness.cpp replaces
// cRDSGroupDecoder by a recording stand-in of the same name.
//
// The group decoder talks to its cRadioReceiver through three members (RadioReceiver.h:77,80,115).  RadioReceiver.cpp
// itself cannot be compiled (Kodi PVR dev-kit, librtlsdr), so
//   * cRadioReceiver::AddUECPDataFrame and ::SetChannelName are DEFINED here as recorders (SetChannelName keeps the
//     reference's return rule, RadioReceiver.cpp:600-612: false while a settings dialog is registered);
//   * IsSettingActive() is the reference's own inline (m_SettingsDialog != nullptr) reading a zeroed stand-in object
//     whose m_SettingsDialog this harness sets; no cRadioReceiver is ever constructed.
// The transport framing of RadioReceiver.cpp:387-414 is therefore NOT covered by THIS library (ref_addon_harness.cpp,
// which compiles RadioReceiver.cpp itself against a fuller stand-in dev-kit, covers it; here: "parity unpinned" for
// those lines: they are restated in oracle/uecp_port.py and checked against hand-made vectors).
// The decoder object is placement-constructed in zeroed storage: the members its constructor and Reset() leave
// uninitialised (m_PTY, m_DI_Finished, m_RadioText_ABFlag, m_PTYN_ABFlag, m_UECPDataFrameSeqCnt) start at zero.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

// every std header the reference headers pull in comes first: "#define private public" must not reach them
#include <limits.h>
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <complex>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "RadioReceiver.h"
#include "RDSGroupDecoder.h"
#undef private
#undef protected

namespace
{
struct RefUecp
{
  cRadioReceiver* radio = nullptr; // zeroed storage, never constructed
  cRDSGroupDecoder* dec = nullptr;
  void* dec_mem = nullptr;
  std::vector<uint8_t> frames; // u16 length (LE) + raw frame bytes, per AddUECPDataFrame call
  std::vector<char> names;     // 9 bytes per SetChannelName call
};
std::mutex g_mutex;
std::map<const cRadioReceiver*, RefUecp*> g_by_radio;

RefUecp* Find(const cRadioReceiver* r)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  auto it = g_by_radio.find(r);
  return it == g_by_radio.end() ? nullptr : it->second;
}
} // namespace

bool cRadioReceiver::AddUECPDataFrame(uint8_t* UECPDataFrame, unsigned int length)
{
  if (RefUecp* h = Find(this))
  {
    h->frames.push_back((uint8_t)(length & 0xff));
    h->frames.push_back((uint8_t)(length >> 8));
    h->frames.insert(h->frames.end(), UECPDataFrame, UECPDataFrame + length);
  }
  return true;
}

bool cRadioReceiver::SetChannelName(std::string name)
{
  if (RefUecp* h = Find(this))
  {
    char b[9] = {0};
    memcpy(b, name.data(), name.size() < 8 ? name.size() : 8);
    h->names.insert(h->names.end(), b, b + 9);
  }
  return m_SettingsDialog == nullptr;
}

extern "C"
{

void* ref_uecp_create(void)
{
  RefUecp* h = new RefUecp();
  h->radio = static_cast<cRadioReceiver*>(calloc(1, sizeof(cRadioReceiver)));
  h->dec_mem = calloc(1, sizeof(cRDSGroupDecoder));
  h->dec = new (h->dec_mem) cRDSGroupDecoder(h->radio);
  std::lock_guard<std::mutex> lock(g_mutex);
  g_by_radio[h->radio] = h;
  return h;
}

void ref_uecp_destroy(void* p)
{
  RefUecp* h = static_cast<RefUecp*>(p);
  if (!h)
    return;
  {
    std::lock_guard<std::mutex> lock(g_mutex);
    g_by_radio.erase(h->radio);
  }
  h->dec->~cRDSGroupDecoder();
  free(h->dec_mem);
  free(h->radio);
  delete h;
}

void ref_uecp_reset(void* p) { static_cast<RefUecp*>(p)->dec->Reset(); }

void ref_uecp_set_setting_active(void* p, int active)
{
  static_cast<RefUecp*>(p)->radio->m_SettingsDialog = active ? reinterpret_cast<cChannelSettings*>(1) : nullptr;
}

void ref_uecp_decode(void* p, const uint16_t* blocks, unsigned n_groups)
{
  RefUecp* h = static_cast<RefUecp*>(p);
  for (unsigned i = 0; i < n_groups; ++i)
  {
    uint16_t b[4] = {blocks[4 * i], blocks[4 * i + 1], blocks[4 * i + 2], blocks[4 * i + 3]};
    h->dec->DecodeRDS(b);
  }
}

unsigned ref_uecp_take_frames(void* p, uint8_t* out, unsigned cap)
{
  RefUecp* h = static_cast<RefUecp*>(p);
  const unsigned n = (unsigned)h->frames.size();
  if (out && n <= cap)
  {
    memcpy(out, h->frames.data(), n);
    h->frames.clear();
  }
  return n;
}

unsigned ref_uecp_take_names(void* p, char* out, unsigned cap)
{
  RefUecp* h = static_cast<RefUecp*>(p);
  const unsigned n = (unsigned)h->names.size();
  if (out && n <= cap)
  {
    memcpy(out, h->names.data(), n);
    h->names.clear();
  }
  return n;
}

} // extern "C"

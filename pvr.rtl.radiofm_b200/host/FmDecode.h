// host/FmDecode.h -- cFmDecoder with the reference's signatures (FmDecode.h:99-165), forwarding to the B200 chain
// through the C ABI (include/radiofm_b200.h).  A maintainer of the add-on replaces `#include "FmDecode.h"` by this
// header and links libradiofm_b200.so; cRadioReceiver::OpenLiveStream / DemuxRead (RadioReceiver.cpp:296-300,
// 515-525) keep calling the same constructor, ProcessStream, Reset and getters.  No CPU fallback: construction
// throws std::runtime_error when no sm_100 device / library is usable.
//
// RDS: the reference hands decoded groups to its cRDSGroupDecoder from inside ProcessStream
// (cFmDecoder -> cRDSRxSignalProcessor m_RDSProcess -> cRDSGroupDecoder m_Decoder(proc): RDSProcess.cpp:312,355,
// RDSGroupDecoder.cpp:979-991 -> cRadioReceiver::AddUECPDataFrame, RadioReceiver.cpp:387).  So does this class: built
// with a receiver (proc != nullptr) it owns a cRDSGroupDecoder (host/RDSGroupDecoder.h) bound to it, and every group
// decoded during a ProcessStream call goes through DecodeRDS -- in order, on the calling thread, before ProcessStream
// returns -- which calls proc->AddUECPDataFrame / SetChannelName / IsSettingActive exactly as the reference does.
// RadioReceiver.cpp needs NO change beyond the include.  (cRadioReceiver must be a complete type by the end of the
// translation unit, as in RadioReceiver.cpp.)  SetRdsGroupSink() replaces that route by a callback of your own:
//     dec.SetRdsGroupSink([&](uint16_t* blk) { ... });
#pragma once

#include <stdint.h>

#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/radiofm_b200.h"
#include "Definitions.h"
#include "RDSGroupDecoder.h"

#define DEFAULT_BANDWIDTH_PCM 15000 // FmDecode.h:22

class cFmDecoder
{
public:
  cFmDecoder(cRadioReceiver* proc, double sample_rate_if, double tuning_offset, double sample_rate_pcm,
             double bandwidth_pcm = DEFAULT_BANDWIDTH_PCM, unsigned int downsample = 1, bool USver = false,
             int cuda_device = -1)
    : m_proc(proc)
  {
    rfm_config cfg;
    rfm_config_default(&cfg);
    cfg.sample_rate_if = sample_rate_if;
    cfg.tuning_offset = tuning_offset;
    cfg.sample_rate_pcm = sample_rate_pcm;
    cfg.bandwidth_pcm = bandwidth_pcm;
    cfg.downsample = downsample;
    cfg.us_deemphasis = USver ? 1 : 0;
    cfg.n_streams = 1;
    cfg.max_block_len = 65536; // FmDecode.cpp:277-282
    cfg.device = cuda_device;
    if (rfm_decoder_create(&cfg, &m_dec) != RFM_OK)
      throw std::runtime_error(std::string("cFmDecoder (B200): ") + rfm_last_error());
    if (proc) // m_RDSProcess(proc, ...) -> m_Decoder(proc), FmDecode.cpp:247, RDSProcess.cpp:43
      m_groups.reset(new cRDSGroupDecoder(proc));
  }
  virtual ~cFmDecoder() { rfm_decoder_destroy(m_dec); }
  cFmDecoder(const cFmDecoder&) = delete;
  cFmDecoder& operator=(const cFmDecoder&) = delete;

  void Reset() // FmDecode.cpp:326-338 -> cRDSRxSignalProcessor::Reset -> m_Decoder.Reset(), RDSProcess.cpp:92
  {
    rfm_decoder_reset(m_dec);
    if (m_groups)
      m_groups->Reset();
  }

  // FmDecode.h:135 -- returns the number of floats written (2 x frames, interleaved L,R)
  unsigned int ProcessStream(const ComplexType* samples_in, unsigned int samples, float* audio)
  {
    uint32_t n_out = 0;
    const int rc = rfm_decoder_process_cf32(m_dec, reinterpret_cast<const float*>(samples_in), samples, audio,
                                            2 * (size_t)samples, &n_out);
    if (rc != RFM_OK)
      return 0; // the reference has no error channel; rfm_last_error() has the text
    DeliverGroups();
    return n_out;
  }
  // the same, straight from the RTL-SDR byte stream (skips the u8 -> float expansion of RTL_SDR_Source.cpp:207-211)
  unsigned int ProcessStreamU8(const uint8_t* iq_u8, unsigned int samples, float* audio)
  {
    uint32_t n_out = 0;
    if (rfm_decoder_process_u8(m_dec, iq_u8, samples, audio, 2 * (size_t)samples, &n_out) != RFM_OK)
      return 0;
    DeliverGroups();
    return n_out;
  }

  bool StereoDetected() const { return Status().stereo_detected != 0; }       // FmDecode.h:140
  RealType GetTuningOffset() const { return Status().tuning_offset; }         // FmDecode.h:146-150
  RealType GetInterfaceLevel() const { return Status().interface_level; }     // FmDecode.h:155
  RealType GetBasebandLevel() const { return Status().baseband_level; }       // FmDecode.h:160
  RealType GetPilotLevel() const { return Status().pilot_level; }             // FmDecode.h:165

  void SetRdsGroupSink(std::function<void(uint16_t*)> sink) { m_sink = std::move(sink); }
  cRadioReceiver* Receiver() const { return m_proc; }

private:
  rfm_stream_status Status() const
  {
    rfm_stream_status s = {};
    rfm_decoder_get_status(m_dec, 0, &s);
    return s;
  }
  void DeliverGroups()
  {
    uint16_t blk[64][4];
    uint32_t n = 0;
    do
    {
      if (rfm_decoder_rds_take_groups(m_dec, 0, &blk[0][0], 64, &n) != RFM_OK)
        return;
      for (uint32_t i = 0; i < n; ++i)
      {
        if (m_sink)
          m_sink(blk[i]);
        else if (m_groups)
          m_groups->DecodeRDS(blk[i]);
      }
    } while (n == 64);
  }

  cRadioReceiver* m_proc;
  rfm_decoder* m_dec = nullptr;
  std::function<void(uint16_t*)> m_sink;
  std::unique_ptr<cRDSGroupDecoder> m_groups; // the reference's m_RDSProcess.m_Decoder
};

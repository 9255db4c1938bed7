// tests/cpp/host_api_driver.cpp -- drives the drop-in C++ classes (pvr.rtl.radiofm_b200/host/) exactly the way
// cRadioReceiver::DemuxRead drives the reference's (RadioReceiver.cpp:515-525): u8 -> complex<float> conversion
// as in RTL_SDR_Source.cpp:207-211, one ProcessStream call per block, audio appended to a file, RDS groups
// collected through the sink.  tests/test_gpu_host_api.py compares the files with the oracle.
//   usage: host_api_driver <iq_u8.bin> <fs> <tuning_offset> <downsample> <block> <audio_out.bin> <groups_out.bin> <shift_hz>
//                          [<uecp_out.bin> [<reset_after_block>]]
// With a 9th argument the decoder is built WITH a receiver and NO sink -- the unchanged-caller case: the class's own
// cRDSGroupDecoder delivers the UECP frames to cRadioReceiver::AddUECPDataFrame from inside ProcessStream; the raw
// frames (u16 length + bytes) go to <uecp_out.bin>.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <string>

#include "FmDecode.h"
#include "FreqShift.h"

class cRadioReceiver // stand-in for the PVR client: the three members cRDSGroupDecoder calls (RadioReceiver.h:77,80,115)
{
public:
  bool AddUECPDataFrame(uint8_t* frame, unsigned len)
  {
    const uint16_t l = (uint16_t)len;
    frames.insert(frames.end(), reinterpret_cast<const uint8_t*>(&l), reinterpret_cast<const uint8_t*>(&l) + 2);
    frames.insert(frames.end(), frame, frame + len);
    return true;
  }
  bool SetChannelName(std::string) { return true; }
  bool IsSettingActive() { return false; }
  std::vector<uint8_t> frames;
};

int main(int argc, char** argv)
{
  if (argc < 9)
    return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f)
    return 3;
  const double fs = atof(argv[2]), off = atof(argv[3]);
  const unsigned ds = (unsigned)atoi(argv[4]), blk = (unsigned)atoi(argv[5]);
  const float shift = (float)atof(argv[8]);
  std::vector<uint8_t> raw(2 * (size_t)blk);
  std::vector<ComplexType> iq(blk);
  std::vector<float> audio(2 * (size_t)blk);
  std::vector<uint16_t> groups;
  try
  {
    cRadioReceiver radio;
    const bool unchanged_caller = argc > 9;
    const int reset_after = argc > 10 ? atoi(argv[10]) : -1;
    cFmDecoder dec(unchanged_caller ? &radio : nullptr, fs, off, 48000.0, DEFAULT_BANDWIDTH_PCM, ds);
    cFreqShift fsh(shift, (RealType)fs, blk);
    if (!unchanged_caller)
      dec.SetRdsGroupSink([&](uint16_t* b) { groups.insert(groups.end(), b, b + 4); });
    FILE* fa = fopen(argv[6], "wb");
    unsigned stereo_blocks = 0, blocks = 0;
    while (fread(raw.data(), 2, blk, f) == blk)
    {
      for (unsigned i = 0; i < blk; ++i) // RTL_SDR_Source.cpp:207-211
        iq[i] = ComplexType(raw[2 * i] / (255.0 / 2.0) - 1.0, raw[2 * i + 1] / (255.0 / 2.0) - 1.0);
      if (shift != 0.0f)
      { // shift up and back down: exercises cFreqShift::Process; the pair is NOT an identity in float
        fsh.Process(iq.data(), blk);
      }
      const unsigned n = dec.ProcessStream(iq.data(), blk, audio.data());
      fwrite(audio.data(), sizeof(float), n, fa);
      stereo_blocks += dec.StereoDetected();
      ++blocks;
      if ((int)blocks == reset_after)
        dec.Reset();
    }
    if (unchanged_caller)
    {
      FILE* fu = fopen(argv[9], "wb");
      fwrite(radio.frames.data(), 1, radio.frames.size(), fu);
      fclose(fu);
    }
    fclose(fa);
    FILE* fg = fopen(argv[7], "wb");
    fwrite(groups.data(), sizeof(uint16_t), groups.size(), fg);
    fclose(fg);
    printf("blocks %u stereo_blocks %u groups %zu if_level %.9g bb_level %.9g pilot %.9g tuning %.9g\n", blocks,
           stereo_blocks, groups.size() / 4, dec.GetInterfaceLevel(), dec.GetBasebandLevel(), dec.GetPilotLevel(),
           dec.GetTuningOffset());
  }
  catch (const std::exception& e)
  {
    fprintf(stderr, "%s\n", e.what());
    return 4;
  }
  fclose(f);
  return 0;
}

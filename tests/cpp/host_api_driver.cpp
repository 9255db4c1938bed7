// tests/cpp/host_api_driver.cpp -- drives the drop-in C++ classes (pvr.rtl.radiofm_b200/host/) exactly the way
// cRadioReceiver::DemuxRead drives the reference's (RadioReceiver.cpp:515-525): u8 -> complex<float> conversion
// as in RTL_SDR_Source.cpp:207-211, one ProcessStream call per block, audio appended to a file, RDS groups
// collected through the sink.  tests/test_gpu_host_api.py compares the files with the oracle.
//   usage: host_api_driver <iq_u8.bin> <fs> <tuning_offset> <downsample> <block> <audio_out.bin> <groups_out.bin> <shift_hz>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "FmDecode.h"
#include "FreqShift.h"

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
    cFmDecoder dec(nullptr, fs, off, 48000.0, DEFAULT_BANDWIDTH_PCM, ds);
    cFreqShift fsh(shift, (RealType)fs, blk);
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

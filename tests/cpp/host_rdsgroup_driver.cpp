// Drives the drop-in cRDSGroupDecoder (host/RDSGroupDecoder.h) the way cRDSRxSignalProcessor does
// (RDSProcess.cpp:312,355): reads groups (u16 x 4 each), writes every raw frame handed to
// cRadioReceiver::AddUECPDataFrame (u16 length + bytes) and every name handed to SetChannelName (8 bytes).
// usage: host_rdsgroup_driver groups.bin frames.bin names.bin
#include <stdio.h>

#include <vector>

#include "RDSGroupDecoder.h" // first: cRadioReceiver is still incomplete here, as in the reference's own header

class cRadioReceiver // stand-in for the PVR client: only the three members the group decoder calls
{
public:
  bool AddUECPDataFrame(uint8_t* frame, unsigned int length)
  {
    frames.push_back((uint8_t)(length & 0xff));
    frames.push_back((uint8_t)(length >> 8));
    frames.insert(frames.end(), frame, frame + length);
    return true;
  }
  bool SetChannelName(std::string name)
  {
    name.resize(8, '\0');
    names.insert(names.end(), name.begin(), name.end());
    return true;
  }
  bool IsSettingActive() { return false; }
  std::vector<uint8_t> frames;
  std::vector<char> names;
};

int main(int argc, char** argv)
{
  if (argc < 4)
    return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f)
    return 3;
  std::vector<uint16_t> g;
  uint16_t blk[4];
  while (fread(blk, sizeof(uint16_t), 4, f) == 4)
    g.insert(g.end(), blk, blk + 4);
  fclose(f);
  cRadioReceiver radio;
  cRDSGroupDecoder dec(&radio);
  for (size_t i = 0; i + 4 <= g.size(); i += 4)
    dec.DecodeRDS(&g[i]);
  f = fopen(argv[2], "wb");
  fwrite(radio.frames.data(), 1, radio.frames.size(), f);
  fclose(f);
  f = fopen(argv[3], "wb");
  fwrite(radio.names.data(), 1, radio.names.size(), f);
  fclose(f);
  printf("groups %zu frame_bytes %zu names %zu\n", g.size() / 4, radio.frames.size(), radio.names.size() / 8);
  return 0;
}

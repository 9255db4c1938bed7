// tests/cpp/host_primitives_driver.cpp -- drives the drop-in primitive classes (pvr.rtl.radiofm_b200/host/
// IirFilter.h, FirFilter.h, DownConvert.h, RDSProcess.h) through the reference's own member signatures, in the call
// order the reference uses them.  tests/test_gpu_host_api.py compares the output files with the oracle.
//   usage: host_primitives_driver <in.f32> <n> <block> <out_prefix>
// in.f32: n float32 samples (real signal; consecutive pairs double as complex samples)
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "DownConvert.h"
#include "FirFilter.h"
#include "IirFilter.h"
#include "RDSProcess.h"

static void dump(const std::string& path, const void* p, size_t bytes)
{
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(p, 1, bytes, f);
  fclose(f);
}

int main(int argc, char** argv)
{
  if (argc < 5)
    return 2;
  const unsigned n = (unsigned)atoi(argv[2]), blk = (unsigned)atoi(argv[3]);
  const std::string pre = argv[4];
  std::vector<float> x(n);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(x.data(), sizeof(float), n, f) != n)
    return 3;
  fclose(f);
  try
  {
    { // cIirFilter: the 19 kHz notch of the audio tail (FmDecode.cpp:285,471) on two channels, then a real run
      cIirFilter notch;
      if (!notch.Init(ftBR, 19000.0f, 5.0f, 48000.0f) || notch.Init(ftConst, 1.0f, 1.0f, 1.0f))
        return 5;
      notch.Init(ftBR, 19000.0f, 5.0f, 48000.0f);
      std::vector<float> a(x.begin(), x.begin() + n / 2), b(x.begin() + n / 2, x.end());
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        notch.ProcessTwo(a.data() + i, b.data() + i, blk);
      dump(pre + "_iir_a.f32", a.data(), a.size() * 4);
      dump(pre + "_iir_b.f32", b.data(), b.size() * 4);
      cIirFilter bp;
      bp.Init(ftBP, 1187.5f, 500.0f, 31250.0f); // RDSProcess.cpp:105
      std::vector<float> r(x);
      for (unsigned i = 0; i + blk <= n; i += blk)
        bp.Process(r.data() + i, blk);
      dump(pre + "_iir_r.f32", r.data(), r.size() * 4);
    }
    { // cFirFilter: the audio low-pass (FmDecode.cpp:286,469) on two channels; the RDS low-pass on complex samples
      cFirFilter lp;
      const int nt = lp.InitLPFilter(0, 1.0f, 60.0f, 15000.0f, 21000.0f, 48000.0f);
      std::vector<float> a(x.begin(), x.begin() + n / 2), b(x.begin() + n / 2, x.end());
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        lp.ProcessTwo(a.data() + i, b.data() + i, blk);
      dump(pre + "_fir_a.f32", a.data(), a.size() * 4);
      dump(pre + "_fir_b.f32", b.data(), b.size() * 4);
      cFirFilter rlp;
      rlp.InitLPFilter(0, 1.0f, 40.0f, 2400.0f, 1.3f * 2400.0f, 31250.0f); // RDSProcess.cpp:99
      std::vector<ComplexType> z(n / 2);
      for (unsigned i = 0; i < n / 2; ++i)
        z[i] = ComplexType(x[2 * i], x[2 * i + 1]);
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        rlp.Process(z.data() + i, blk);
      dump(pre + "_fir_z.f32", z.data(), z.size() * 8);
      // the two-buffer complex form (FirFilter.cpp:421-445) and a high-pass design (:195-264)
      cFirFilter hp;
      const int nh = hp.InitHPFilter(0, 1.0f, 50.0f, 6000.0f, 3000.0f, 48000.0f);
      std::vector<ComplexType> zi(n / 2), zo(n / 2);
      for (unsigned i = 0; i < n / 2; ++i)
        zi[i] = ComplexType(x[2 * i], x[2 * i + 1]);
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        hp.Process(zi.data() + i, zo.data() + i, blk);
      dump(pre + "_fir_hp.f32", zo.data(), zo.size() * 8);
      printf("fir_hp_taps %d\n", nh);
      printf("fir_taps %d\n", nt);
    }
    { // CRDSDownConvert as cRDSRxSignalProcessor sets it up (RDSProcess.cpp:46-48)
      CRDSDownConvert dc;
      const RealType rate = dc.SetDataRate(250000.0f, 8000.0f);
      dc.SetFrequency(-57000.0f);
      std::vector<ComplexType> z(n / 2), y(n / 2);
      for (unsigned i = 0; i < n / 2; ++i)
        z[i] = ComplexType(x[2 * i], x[2 * i + 1]);
      size_t m = 0;
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        m += (size_t)dc.ProcessData((int)blk, z.data() + i, y.data() + m);
      dump(pre + "_dc.f32", y.data(), m * 8);
      printf("dc_rate %.9g dc_out %zu\n", rate, m);
    }
    { // cDownsampleFilter as cFmDecoder builds it at 1.0 MS/s (FmDecode.cpp:257-267): complex / 4, real -> 48 kHz
      cDownsampleFilter in(32, 0.6 / 4, 4, true), mono(250, 15000.0 / 250000.0, 250000.0 / 48000.0, false);
      std::vector<ComplexType> z(n / 2), y(n / 2);
      for (unsigned i = 0; i < n / 2; ++i)
        z[i] = ComplexType(x[2 * i], x[2 * i + 1]);
      size_t m = 0;
      for (unsigned i = 0; i + blk <= n / 2; i += blk)
        m += in.Process(z.data() + i, y.data() + m, blk);
      dump(pre + "_ds_c.f32", y.data(), m * 8);
      std::vector<float> a(n);
      size_t k = 0;
      for (unsigned i = 0; i + blk <= n; i += blk)
        k += mono.Process(x.data() + i, a.data() + k, blk);
      dump(pre + "_ds_r.f32", a.data(), k * 4);
      printf("ds_c %zu ds_r %zu\n", m, k);
    }
    { // cRDSRxSignalProcessor on the real signal
      cRDSRxSignalProcessor rds(nullptr, 250000.0f);
      std::vector<uint8_t> bits;
      std::vector<uint16_t> groups;
      rds.SetBitSink([&](int b) { bits.push_back((uint8_t)b); });
      rds.SetGroupSink([&](uint16_t* g) { groups.insert(groups.end(), g, g + 4); });
      for (unsigned i = 0; i + blk <= n; i += blk)
        rds.Process(x.data() + i, blk);
      dump(pre + "_rds_bits.u8", bits.data(), bits.size());
      dump(pre + "_rds_groups.u16", groups.data(), groups.size() * 2);
      printf("rds_bits %zu rds_groups %zu\n", bits.size(), groups.size() / 4);
    }
  }
  catch (const std::exception& e)
  {
    fprintf(stderr, "%s\n", e.what());
    return 4;
  }
  return 0;
}

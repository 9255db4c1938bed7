// tools/exhaustive_math.cpp -- TEST INFRASTRUCTURE.  Exhaustive / massive CPU validation of rfm_math.cuh (the header
// is __host__ __device__; the product only ever runs it on the GPU):
//   sincos : EVERY float in [0, 16) (and the negatives), rfm_sincos vs float(sin/cos(double)) [= x87 fsincos -> float]
//   atan2f : rfm_atan2f (restructured) vs glibc atan2f, N random pairs per scale + every float a in atanf via (a, 1)
// g++ -O2 -std=c++17 -ffp-contract=off -fopenmp -I pvr.rtl.radiofm_b200/csrc tools/exhaustive_math.cpp -o /tmp/exh -lm
#include "rfm_math.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <random>
#ifdef _OPENMP
#include <omp.h>
#endif
using namespace rfm;
int main(int argc, char** argv)
{
  unsigned long long bad_s = 0, bad_c = 0, n = 0;
  const uint32_t top = f2u(16.0f);
#pragma omp parallel for reduction(+ : bad_s, bad_c, n) schedule(dynamic, 1 << 20)
  for (uint32_t u = 0; u < top; ++u)
  {
    for (int sg = 0; sg < 2; ++sg)
    {
      const float x = u2f(u | (sg ? 0x80000000u : 0u));
      float s, c;
      rfm_sincos(x, &s, &c);
      bad_s += f2u(s) != f2u((float)sin((double)x));
      bad_c += f2u(c) != f2u((float)cos((double)x));
#if defined(__x86_64__)
      double ds, dc; // the instruction the reference itself executes (FmDecode.cpp:167,386)
      __asm__("fsincos" : "=t"(dc), "=u"(ds) : "0"((double)x));
      bad_s += f2u(s) != f2u((float)ds);
      bad_c += f2u(c) != f2u((float)dc);
#endif
      ++n;
    }
  }
  printf("sincos: %llu floats in (-16, 16): sin mismatches %llu, cos mismatches %llu\n", n, bad_s, bad_c);
  unsigned long long bad_a = 0, na = 0;
#pragma omp parallel for reduction(+ : bad_a, na) schedule(dynamic, 1 << 20)
  for (uint32_t u = 0; u < 0x7f800000u; ++u)
  {
    const float a = u2f(u);
    for (int k = 0; k < 4; ++k)
    {
      const float y = (k & 1) ? -a : a, x = (k & 2) ? -1.0f : 1.0f;
      bad_a += f2u(rfm_atan2f(y, x)) != f2u(atan2f(y, x));
      ++na;
    }
  }
  printf("atan2f(+-a, +-1) every finite a: %llu cases, mismatches %llu\n", na, bad_a);
  // phase wraps: every float of the domain
  unsigned long long bad_w = 0, nw = 0;
#pragma omp parallel for reduction(+ : bad_w, nw) schedule(dynamic, 1 << 20)
  for (uint32_t u = f2u(RFM_2PI_HI); u < f2u(12.5f); ++u)
  {
    const float p = u2f(u);
    bad_w += f2u(rfm_sub_2pi(p)) != f2u((float)fmod((double)p, RFM_K_2PI));
    bad_w += f2u(rfm_sub_2pi(p)) != f2u((float)((double)p - RFM_K_2PI));
    ++nw;
  }
#pragma omp parallel for reduction(+ : bad_w, nw) schedule(dynamic, 1 << 20)
  for (uint32_t u = 1; u <= f2u(6.0f); ++u)
  {
    const float p = -u2f(u);
    bad_w += f2u(rfm_add_2pi(p)) != f2u((float)((double)p + RFM_K_2PI));
    ++nw;
  }
  // wrap functions against the reference's double expressions on a dense sweep incl. the boundaries
#pragma omp parallel for reduction(+ : bad_w, nw) schedule(dynamic, 1 << 20)
  for (uint32_t u = 0; u < f2u(12.4f); ++u)
  {
    for (int sg = 0; sg < 2; ++sg)
    {
      float p = sg ? -u2f(u) : u2f(u);
      if (p < -6.0f) continue;
      float a = p;
      if ((double)a >= RFM_K_2PI) a = (float)fmod((double)a, RFM_K_2PI);
      while (a < 0.0f) a = (float)(a + RFM_K_2PI);
      bad_w += f2u(rfm_wrap_demod(p)) != f2u(a);
      float b = p;
      if (b > RFM_K_2PI) b = (float)(b - RFM_K_2PI);
      if (p > 0.0f) bad_w += f2u(rfm_wrap_pilot(p)) != f2u(b);
      ++nw;
    }
  }
  printf("phase wraps: %llu cases, mismatches %llu\n", nw, bad_w);
  const unsigned long long N = argc > 1 ? strtoull(argv[1], 0, 10) : 400000000ull;
  unsigned long long bad_r = 0;
#pragma omp parallel reduction(+ : bad_r)
  {
    std::mt19937_64 g(1234 + 77 * (unsigned)
#ifdef _OPENMP
                      omp_get_thread_num()
#else
                      0
#endif
    );
    std::uniform_int_distribution<uint32_t> U;
#pragma omp for
    for (long long i = 0; i < (long long)N; ++i)
    {
      float y = u2f(U(g)), x = u2f(U(g));
      if (i & 1) { // similar magnitudes (the interesting ranges)
        x = u2f((f2u(y) & 0x7f800000u) | (U(g) & 0x807fffffu));
        if ((i & 6) == 2) x = u2f(f2u(x) + ((U(g) % 5) << 23) - (2u << 23));
      }
      const float a = rfm_atan2f(y, x), b = atan2f(y, x);
      bad_r += f2u(a) != f2u(b) && !(a != a && b != b);
    }
  }
  printf("atan2f random: %llu pairs, mismatches %llu\n", N, bad_r);
  return (bad_a || bad_r || bad_s || bad_c || bad_w) ? 1 : 0;
}

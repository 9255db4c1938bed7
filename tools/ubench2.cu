// tools/ubench2.cu -- per-step latency of the lane recurrences (one warp, 32 different streams)
#include <cstdio>
#include <cuda_runtime.h>
#include "rfm_steps.cuh"
using namespace rfm;
#define N 2048
template <bool FAST> __global__ void k_demod(const float2* z, DemodConst k, long long* cyc, float* sink)
{
  __shared__ float2 zs[32][65];
  DemodState st = {0.f, 0.18f};
  float acc = 0.f;
  long long total = 0;
  for (int t0 = 0; t0 < N; t0 += 64) {
    for (int r = 0; r < 32; ++r) { zs[r][threadIdx.x] = z[r * N + t0 + threadIdx.x]; zs[r][threadIdx.x + 32] = z[r * N + t0 + 32 + threadIdx.x]; }
    __syncwarp();
    long long a = clock64();
    for (int i = 0; i < 64; ++i) { float2 x = zs[threadIdx.x][i]; if (FAST) { bool bad = false; demod_step_fast(st, x.x, x.y, k, bad); acc += bad ? 1.f : 0.f; } else demod_step(st, x.x, x.y, k); acc += st.incr; }
    total += clock64() - a;
    __syncwarp();
  }
  if (threadIdx.x == 0) *cyc = total;
  sink[threadIdx.x] = acc + st.phase;
}
template <bool FAST, bool DC = true> __global__ void k_pilot(const float* bbin, PilotConstDev k, long long* cyc, float* sink)
{
  __shared__ float zs[32][65];
  PilotState st = {0.f, 0.5472f, 0, 0, 0, 0, 0, 1000.f};
  float acc = 0.f, dc = 0.f; long long total = 0;
  for (int t0 = 0; t0 < N; t0 += 64) {
    for (int r = 0; r < 32; ++r) { zs[r][threadIdx.x] = bbin[r * N + t0 + threadIdx.x]; zs[r][threadIdx.x + 32] = bbin[r * N + t0 + 32 + threadIdx.x]; }
    __syncwarp();
    long long a = clock64();
    for (int i = 0; i < 64; ++i) { if (FAST) { bool bad = false; float dcv = DC ? demod_output(zs[threadIdx.x][i], dc, 0.57f) : zs[threadIdx.x][i]; acc += pilot_step_fast(st, dcv, k, rfm_sincos_regs(), bad); acc += bad ? 1.f : 0.f; } else acc += pilot_step(st, zs[threadIdx.x][i], k); }
    total += clock64() - a;
    __syncwarp();
  }
  if (threadIdx.x == 0) *cyc = total;
  sink[threadIdx.x] = acc + st.phase;
}
__global__ void k_dc(const float* bbin, long long* cyc, float* sink)
{
  float dc = 0.f, acc = 0.f;
  long long a = clock64();
  for (int i = 0; i < N; ++i) acc += demod_output(bbin[threadIdx.x * N + i] , dc, 0.55f);
  if (threadIdx.x == 0) *cyc = clock64() - a;
  sink[threadIdx.x] = acc;
}
int main()
{
  // 32 streams: FM signals with different tones, Fb = 218181.8
  static float2 hz[32 * N]; static float hb[32 * N];
  const double fb = 218181.8;
  for (int r = 0; r < 32; ++r) {
    double ph = 0.3 * r;
    for (int i = 0; i < N; ++i) {
      double m = 0.4 * sin(2 * M_PI * (300.0 + 100 * r) * i / fb) + 0.1 * sin(2 * M_PI * 19000.0 * i / fb);
      ph += 2 * M_PI * (6250.0 + 75000.0 * m) / fb;
      double q = 1.0 / 127.5;
      hz[r * N + i] = make_float2((float)(floor(0.8 * cos(ph) / q + 0.5) * q * 1.3), (float)(floor(0.8 * sin(ph) / q + 0.5) * q * 1.3));
      hb[r * N + i] = (float)m;
    }
  }
  float2* dz; float* db; long long* dc; float* sink; long long h;
  cudaMalloc(&dz, sizeof(hz)); cudaMalloc(&db, sizeof(hb)); cudaMalloc(&dc, 8); cudaMalloc(&sink, 128);
  cudaMemcpy(dz, hz, sizeof(hz), cudaMemcpyHostToDevice); cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice);
  const float fac = (float)(2 * M_PI / fb);
  DemodConst k; k.gain = 0.57f; k.hi = 0.95f * 0.5f * (float)fb * fac; k.lo = -k.hi; k.alpha = 0.125f * 0.85f * (float)fb * fac; k.beta = k.alpha * k.alpha / 2.0f;
  PilotConstDev pk = {0.5457f, 0.5486f, 9.5e-6f, -1.9919f, 0.99195f, 0.000892f, -0.000892f * 0.99983f, 0.04f, 87272};
  for (int it = 0; it < 2; ++it) { k_demod<false><<<1, 32>>>(dz, k, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("demod_step        %7.1f cycles/sample\n", (double)h / N);
  for (int it = 0; it < 2; ++it) { k_demod<true><<<1, 32>>>(dz, k, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("demod_step_fast   %7.1f cycles/sample\n", (double)h / N);
  for (int it = 0; it < 2; ++it) { k_pilot<false><<<1, 32>>>(db, pk, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("pilot_step        %7.1f cycles/sample\n", (double)h / N);
  for (int it = 0; it < 2; ++it) { k_pilot<true><<<1, 32>>>(db, pk, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("dc+pilot_step_fast%7.1f cycles/sample\n", (double)h / N);
  for (int it = 0; it < 2; ++it) { k_pilot<true, false><<<1, 32>>>(db, pk, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("pilot_step_fast   %7.1f cycles/sample (no dc)\n", (double)h / N);
  for (int it = 0; it < 2; ++it) { k_dc<<<1, 32>>>(db, dc, sink); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost); printf("demod_output %7.1f cycles/sample\n", (double)h / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

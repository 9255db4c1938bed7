// tools/microbench/f32x2.cu -- does the packed fp32 pair arithmetic of sm_100a (fma.rn.f32x2 -> FFMA2) halve the issue
// cost of an EXACT complex-by-real tap (round(a*b) then round(acc + .), no contraction)?  Four variants of the same
// multiply-accumulate stream, timed with CUDA events, plus a bit-for-bit check of the packed forms against FMUL / FADD.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu && ./f32x2
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// exact product / exact sum as FFMA2: a*b + (-0) rounds a*b once (and keeps the sign of a zero product); a*1 + c rounds a+c once
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2 even under --fmad false, and folds the literal forms of
// the two lines above the same way -- so the constants come in as kernel arguments it cannot see through
struct K2 { u64 negzero, one; };
__device__ __forceinline__ u64 mul2x(u64 a, u64 b, const K2& k) { return fma2(a, b, k.negzero); }
__device__ __forceinline__ u64 add2x(u64 a, u64 c, const K2& k) { return fma2(a, k.one, c); }

constexpr int NACC = 8, ITERS = 2048;

template <int V, int EXTRA>
__global__ void __launch_bounds__(256) k_bench(const float* __restrict__ in, float* __restrict__ out, int iters, K2 kk)
{
  unsigned z = threadIdx.x;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  float xr[NACC], xi[NACC], ar[NACC], ai[NACC];
  u64 x2[NACC], a2[NACC];
  for (int k = 0; k < NACC; ++k) {
    xr[k] = in[(t + k) & 1023]; xi[k] = in[(t + k + 7) & 1023]; ar[k] = 0.f; ai[k] = 0.f;
    x2[k] = pack2(xr[k], xi[k]); a2[k] = 0;
  }
  float h = in[t & 1023];
  for (int it = 0; it < iters; ++it) {
    const u64 h2 = pack2(h, h);
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
      if (V == 0) { ar[k] = __fadd_rn(ar[k], __fmul_rn(xr[k], h)); ai[k] = __fadd_rn(ai[k], __fmul_rn(xi[k], h)); }
      if (V == 1) { a2[k] = add2x(mul2x(x2[k], h2, kk), a2[k], kk); }
#pragma unroll
      for (int e = 0; e < EXTRA; ++e) z = (z ^ (unsigned)(it + k)) + (z >> 3);        // alu-pipe filler: LOP3 / SHF / IADD3
      if (V == 2) { a2[k] = fma2(x2[k], h2, a2[k]); }
      if (V == 3) { ar[k] = __fmaf_rn(xr[k], h, ar[k]); ai[k] = __fmaf_rn(xi[k], h, ai[k]); }
    }
    h = __fmul_rn(h, 0.99999f);
  }
  float s = 0.f;
  for (int k = 0; k < NACC; ++k) {
    if (V == 1 || V == 2) unpack2(a2[k], ar[k], ai[k]);
    s += ar[k] + ai[k];
  }
  out[t] = s + (float)z;
}

__global__ void k_check(const uint32_t* a, const uint32_t* b, uint32_t* bad, int n, K2 kk)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (2 * i + 1 >= n) return;
  const float a0 = __uint_as_float(a[2 * i]), a1 = __uint_as_float(a[2 * i + 1]);
  const float b0 = __uint_as_float(b[2 * i]), b1 = __uint_as_float(b[2 * i + 1]);
  float m0, m1, s0, s1;
  unpack2(mul2x(pack2(a0, a1), pack2(b0, b1), kk), m0, m1);
  unpack2(add2x(pack2(a0, a1), pack2(b0, b1), kk), s0, s1);
  const float em0 = __fmul_rn(a0, b0), em1 = __fmul_rn(a1, b1), es0 = __fadd_rn(a0, b0), es1 = __fadd_rn(a1, b1);
  auto same = [](float x, float y) { return (x != x && y != y) || __float_as_uint(x) == __float_as_uint(y); };
  if (!same(m0, em0) || !same(m1, em1)) atomicAdd(&bad[0], 1u);
  if (!same(s0, es0) || !same(s1, es1)) atomicAdd(&bad[1], 1u);
}

static const K2 KK = {0x8000000080000000ull, 0x3f8000003f800000ull};
template <int V, int EXTRA> static float run(const float* in, float* out, int blocks)
{
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_bench<V, EXTRA><<<blocks, 256>>>(in, out, ITERS, KK); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) k_bench<V, EXTRA><<<blocks, 256>>>(in, out, ITERS, KK);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8;
  float *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, (size_t)blocks * 256 * 4);
  float hin[1024]; for (int i = 0; i < 1024; ++i) hin[i] = 0.5f + (float)(i % 37) / 64.f;
  cudaMemcpy(in, hin, 4096, cudaMemcpyHostToDevice);
  const double taps = (double)blocks * 256 * NACC * ITERS;       // complex-by-real taps per launch
  const char* names[4] = {"FMUL+FADD scalar (exact)", "2 x FFMA2 packed (exact)", "1 x FFMA2 fused (inexact)", "FFMA scalar (inexact)"};
  float ms[3][4] = {{run<0, 0>(in, out, blocks), run<1, 0>(in, out, blocks), run<2, 0>(in, out, blocks), run<3, 0>(in, out, blocks)},
                    {run<0, 1>(in, out, blocks), run<1, 1>(in, out, blocks), run<2, 1>(in, out, blocks), run<3, 1>(in, out, blocks)},
                    {run<0, 2>(in, out, blocks), run<1, 2>(in, out, blocks), run<2, 2>(in, out, blocks), run<3, 2>(in, out, blocks)}};
  for (int e = 0; e < 3; ++e)
    for (int v = 0; v < 4; ++v)
      printf("filler %d x3 alu ops per tap  %-28s %.3f ms  %.2f taps/clk/SM (at %d MHz nominal)\n", e, names[v], ms[e][v],
             taps / (ms[e][v] * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  // exactness of the packed product and sum: random bit patterns (NaN, Inf, subnormals included) and near-cancelling sums
  const int n = 1 << 24;
  uint32_t *ha = (uint32_t*)malloc(4 * n), *hb = (uint32_t*)malloc(4 * n), *da, *db, *dbad;
  cudaMalloc(&da, 4 * n); cudaMalloc(&db, 4 * n); cudaMalloc(&dbad, 8);
  uint64_t s = 88172645463325252ull; unsigned tot[2] = {0, 0};
  for (int round = 0; round < 8; ++round) {
    for (int i = 0; i < n; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17; ha[i] = (uint32_t)(s >> 16);
      s ^= s << 13; s ^= s >> 7; s ^= s << 17; hb[i] = (uint32_t)(s >> 16);
      if (round & 1) hb[i] = (ha[i] ^ 0x80000000u) + (uint32_t)((s >> 50) & 0xff) - 128u;   // b ~ -a: cancellation, subnormal sums
      if ((round & 2) && (i & 3) == 0) { ha[i] &= 0x807fffffu; }                              // subnormal operands
    }
    cudaMemcpy(da, ha, 4 * n, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, 4 * n, cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 8);
    k_check<<<(n / 2 + 255) / 256, 256>>>(da, db, dbad, n, KK);
    unsigned bad[2]; cudaMemcpy(bad, dbad, 8, cudaMemcpyDeviceToHost); tot[0] += bad[0]; tot[1] += bad[1];
  }
  printf("packed product mismatches %u, packed sum mismatches %u over %d pairs\n", tot[0], tot[1], 8 * (n / 2) * 2);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

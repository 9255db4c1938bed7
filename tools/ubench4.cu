// tools/ubench4.cu -- issue throughput of un-fused complex x real MACs: scalar FMUL/FADD (4 instr per complex tap) versus
// the sm_100 packed forms FMUL2/FADD2/FFMA2 (2 instr per complex tap), plus a bit-exactness check of the packed forms
// against __fmul_rn/__fadd_rn (incl. denormals, signed zeros, inf/nan).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#define ACC 8
#define ITER 2048
// MODE 0: scalar FMUL,FMUL,FADD,FADD   1: FFMA2(x,c,-0)+FADD2   2: FMUL2 + FFMA2(t,1,acc)   3: fused FFMA2 (reference ceiling)
// 4: scalar FFMA x2 (fused ceiling)
template <int MODE>
__global__ void __launch_bounds__(256) tput(const float* seed, float* sink, u64 nz, u64 one, long long* cyc)
{
  float xr[ACC], xi[ACC], ar[ACC], ai[ACC];
  u64 x2[ACC], a2[ACC];
  float c = seed[0];
#pragma unroll
  for (int k = 0; k < ACC; ++k) {
    xr[k] = seed[1 + k] + threadIdx.x * 1e-6f; xi[k] = seed[9 + k]; ar[k] = 0.f; ai[k] = 0.f;
    x2[k] = pk(xr[k], xi[k]); a2[k] = 0;
  }
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int k = 0; k < ACC; ++k) {
      if (MODE == 0) { ar[k] = __fadd_rn(ar[k], __fmul_rn(xr[k], c)); ai[k] = __fadd_rn(ai[k], __fmul_rn(xi[k], c)); }
      if (MODE == 1) a2[k] = add2(a2[k], fma2(x2[k], pk(c, c), nz));
      if (MODE == 2) a2[k] = fma2(mul2(x2[k], pk(c, c)), one, a2[k]);
      if (MODE == 3) a2[k] = fma2(x2[k], pk(c, c), a2[k]);
      if (MODE == 4) { ar[k] = __fmaf_rn(xr[k], c, ar[k]); ai[k] = __fmaf_rn(xi[k], c, ai[k]); }
    }
    c = __fadd_rn(c, 1e-7f);   // keeps the products loop-variant (1 extra instr per 8 taps)
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < ACC; ++k) { float2 v = upk(a2[k]); s += ar[k] + ai[k] + v.x + v.y; }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__global__ void exact(const float* a, const float* b, const float* c, unsigned* bad, int n, u64 nz, u64 one)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = a[i], y = b[i], z = c[i], w = a[(i + 7) % n];
  float pr = __fmul_rn(x, y), pi = __fmul_rn(w, y);
  float sr = __fadd_rn(z, pr), si = __fadd_rn(x, pi);
  float2 m1 = upk(fma2(pk(x, w), pk(y, y), nz));
  float2 m2 = upk(mul2(pk(x, w), pk(y, y)));
  float2 s1 = upk(add2(pk(z, x), pk(pr, pi)));
  float2 s2 = upk(fma2(pk(pr, pi), one, pk(z, x)));
  auto ne = [](float p, float q) { return __float_as_uint(p) != __float_as_uint(q) && !(p != p && q != q); };
  if (ne(m1.x, pr) || ne(m1.y, pi)) atomicAdd(bad + 0, 1u);
  if (ne(m2.x, pr) || ne(m2.y, pi)) atomicAdd(bad + 1, 1u);
  if (ne(s1.x, sr) || ne(s1.y, si)) atomicAdd(bad + 2, 1u);
  if (ne(s2.x, sr) || ne(s2.y, si)) atomicAdd(bad + 3, 1u);
}

template <int MODE>
void run(const char* name, int sms)
{
  float h[32]; for (int i = 0; i < 32; ++i) h[i] = 0.5f + 0.01f * i;
  float *seed, *sink; long long* cyc;
  const int blocks = sms * 4, thr = 256;
  cudaMalloc(&seed, sizeof h); cudaMalloc(&sink, blocks * thr * 4); cudaMalloc(&cyc, 8);
  cudaMemcpy(seed, h, sizeof h, cudaMemcpyHostToDevice);
  float nzf = -0.0f, onef = 1.0f; uint32_t nzb, oneb; memcpy(&nzb, &nzf, 4); memcpy(&oneb, &onef, 4);
  u64 nz = ((u64)nzb << 32) | nzb, one = ((u64)oneb << 32) | oneb;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  tput<MODE><<<blocks, thr>>>(seed, sink, nz, one, cyc);
  cudaEventRecord(e0);
  tput<MODE><<<blocks, thr>>>(seed, sink, nz, one, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  const double taps_per_sm = 4.0 * thr * (double)ACC * ITER;   // complex taps per SM (4 resident CTAs)
  printf("%-44s %8.3f ms  %7.1f complex taps/clk/SM (block 0: %lld cycles)  %s\n", name, ms, taps_per_sm / (double)hc, hc,
         cudaGetErrorString(cudaGetLastError()));
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("scalar FMUL FMUL FADD FADD", p.multiProcessorCount);
  run<1>("FFMA2(x,c,-0) + FADD2", p.multiProcessorCount);
  run<2>("FMUL2 + FFMA2(t,1,acc)", p.multiProcessorCount);
  run<3>("fused FFMA2 (not usable: ceiling)", p.multiProcessorCount);
  run<4>("fused scalar FFMA x2 (not usable: ceiling)", p.multiProcessorCount);
  // exactness
  const int n = 1 << 24;
  float *ha = (float*)malloc(n * 4), *hb = (float*)malloc(n * 4), *hc = (float*)malloc(n * 4);
  srand(7);
  auto rnd = [](int mode) {
    uint32_t u = ((uint32_t)rand() << 16) ^ (uint32_t)rand() ^ ((uint32_t)rand() << 31);
    if (mode == 1) u &= 0x807fffffu;                       // denormals / zeros
    if (mode == 2) u = (u & 0x80ffffffu) | 0x3f000000u;    // around 1
    float f; memcpy(&f, &u, 4); return f;
  };
  for (int i = 0; i < n; ++i) { int m = i % 4 == 0 ? 1 : (i % 4 == 1 ? 0 : 2); ha[i] = rnd(m); hb[i] = rnd((i / 5) % 3); hc[i] = rnd(2); }
  float *da, *db, *dc; unsigned* bad; unsigned hbad[4];
  cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dc, n * 4); cudaMalloc(&bad, 16); cudaMemset(bad, 0, 16);
  cudaMemcpy(da, ha, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(dc, hc, n * 4, cudaMemcpyHostToDevice);
  float nzf = -0.0f, onef = 1.0f; uint32_t nzb, oneb; memcpy(&nzb, &nzf, 4); memcpy(&oneb, &onef, 4);
  exact<<<n / 256, 256>>>(da, db, dc, bad, n, ((u64)nzb << 32) | nzb, ((u64)oneb << 32) | oneb);
  cudaMemcpy(hbad, bad, 16, cudaMemcpyDeviceToHost);
  printf("exactness over %d cases: fma2(x,c,-0) vs fmul: %u bad; mul2: %u bad; add2 vs fadd: %u bad; fma2(t,1,acc): %u bad\n", n, hbad[0], hbad[1], hbad[2], hbad[3]);
  return 0;
}

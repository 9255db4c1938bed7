// tools/ubench.cu -- dependent-chain latency microbenchmarks on the B200 (one warp, clock64 around N
// dependent operations).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../pvr.rtl.radiofm_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "rfm_math.cuh"
using namespace rfm;

#define N 512
template <typename F>
__global__ void chain(F f, float seed, long long* cyc, float* sink)
{
  float x = seed + threadIdx.x * 1e-3f;
  // warm
  for (int i = 0; i < 16; ++i) x = f(x);
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) x = f(x);
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  sink[threadIdx.x] = x;
}

template <typename F>
void run(const char* name, F f, float seed = 0.7f)
{
  long long* d; float* s; long long h;
  cudaMalloc(&d, 8); cudaMalloc(&s, 128);
  chain<<<1, 32>>>(f, seed, d, s);
  chain<<<1, 32>>>(f, seed, d, s);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.1f cycles/op\n", name, (double)h / N);
  cudaFree(d); cudaFree(s);
}

int main()
{
  run("fmul", [] __device__(float x) { return __fmul_rn(x, 1.0000001f); });
  run("fadd", [] __device__(float x) { return __fadd_rn(x, 1e-7f); });
  run("ffma", [] __device__(float x) { return __fmaf_rn(x, 1.0000001f, 1e-7f); });
  run("fmnmx pair", [] __device__(float x) { return fmaxf(0.1f, fminf(x + 1e-7f, 5.f)); });
  run("fdiv_rn", [] __device__(float x) { return __fdiv_rn(1.3f, x) ; }, 1.1f);
  run("frcp_rn", [] __device__(float x) { return __frcp_rn(x) ; }, 1.1f);
  run("mufu.rcp (fast)", [] __device__(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }, 1.1f);
  run("fsqrt_rn", [] __device__(float x) { return __fsqrt_rn(x) + 0.5f; }, 1.1f);
  run("rintf", [] __device__(float x) { return rintf(x * 1.37f) * 0.25f + 0.3f; });
  run("f2d+d2f", [] __device__(float x) { return __double2float_rn((double)x); });
  run("dadd (via cvt)", [] __device__(float x) { return __double2float_rn(__dadd_rn((double)x, 1e-9)); });
  run("dmul+dadd (via cvt)", [] __device__(float x) { return __double2float_rn(__dadd_rn(__dmul_rn((double)x, 1.0000001), 1e-9)); });
  run("dfma x4 (via cvt)", [] __device__(float x) { double d = x; d = __fma_rn(d, 1.0000001, 1e-9); d = __fma_rn(d, 1.0000001, 1e-9); d = __fma_rn(d, 1.0000001, 1e-9); d = __fma_rn(d, 1.0000001, 1e-9); return __double2float_rn(d); });
  run("dfma x8 (via cvt)", [] __device__(float x) { double d = x;
#pragma unroll
    for (int i = 0; i < 8; ++i) d = __fma_rn(d, 1.0000001, 1e-9); return __double2float_rn(d); });
  run("rfm_sincos (s+c)", [] __device__(float x) { float s, c; rfm_sincos(x, &s, &c); return __fadd_rn(__fmul_rn(s, 3.0f), __fadd_rn(c, 3.1f)); }, 1.0f);
  run("rfm_atan2f(y=x,0.8)", [] __device__(float x) { return __fadd_rn(rfm_atan2f(x, 0.8f), 0.3f); }, 0.5f);
  run("rfm_atan_core", [] __device__(float x) { return __fadd_rn(rfm_atan_core(x), 0.3f); }, 0.5f);
  run("sincos->s only", [] __device__(float x) { float s, c; rfm_sincos(x, &s, &c); return __fadd_rn(s, 1.5f); }, 1.0f);
  run("wrap_pilot(x+0.547)", [] __device__(float x) { return rfm_wrap_pilot(__fadd_rn(x, 0.547f)); }, 1.0f);
  run("wrap_demod(x+1.3)", [] __device__(float x) { return rfm_wrap_demod(__fadd_rn(x, 1.3f)); }, 1.0f);
  run("d2f only (x const dbl add)", [] __device__(float x) { double d = (double)x; float a = __double2float_rn(d + 1.0); float b = __double2float_rn(d + 2.0); return a + b; }, 1.0f);
  run("sel chain x4", [] __device__(float x) { float a = x > 0.5f ? x : -x; a = a > 0.7f ? a * 0.5f : a; a = a < 0.2f ? a + 0.3f : a; return a; }, 1.0f);
  run("sincosf (cuda libm)", [] __device__(float x) { float s, c; sincosf(x, &s, &c); return s * 3.0f + c + 3.1f; }, 1.0f);
  run("sin+cos double libm", [] __device__(float x) { double s, c; sincos((double)x, &s, &c); return (float)(s * 3.0 + c + 3.1); }, 1.0f);
  run("lds roundtrip", [] __device__(float x) { __shared__ float sm[64]; sm[threadIdx.x] = x; __syncwarp(); float y = sm[threadIdx.x ^ 1]; __syncwarp(); return y + 1e-7f; });
  return 0;
}

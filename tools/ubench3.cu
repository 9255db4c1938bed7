// tools/ubench3.cu -- where do the cycles of rfm_sincos go?  (dependent chains, one warp)
#include <cstdio>
#include <cuda_runtime.h>
#include "rfm_math.cuh"
using namespace rfm;
#define N 512
template <typename F>
__global__ void chain(F f, float seed, long long* cyc, float* sink)
{
  float x = seed + threadIdx.x * 1e-3f;
  for (int i = 0; i < 16; ++i) x = f(x);
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = f(x);
  long long t1 = clock64();
  if (threadIdx.x == 0) *cyc = t1 - t0;
  sink[threadIdx.x] = x;
}
template <typename F>
void run(const char* name, F f, float seed = 1.0f)
{
  long long* d; float* s; long long h;
  cudaMalloc(&d, 8); cudaMalloc(&s, 128);
  chain<<<1, 32>>>(f, seed, d, s);
  chain<<<1, 32>>>(f, seed, d, s);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %7.1f cycles/op\n", name, (double)h / N);
  cudaFree(d); cudaFree(s);
}
template <int MODE>
__device__ __forceinline__ float sc(float phase)
{
  const float magic = 12582912.0f;
  const float t = __fmaf_rn(phase, 6.36619772367581382433e-01f, magic);
  const float kf = __fsub_rn(t, magic);
  const int q = (int)__float_as_uint(t);
  const float r1 = __fmaf_rn(-kf, 1.57079637050628662109375f, phase);
  if (MODE == 0) return __fadd_rn(r1, 1.0f);                       // float reduction only
  const double kd = (double)kf;
  const double r = __fma_rn(-kd, -4.37113900018624283e-08, (double)r1);
  if (MODE == 1) return __double2float_rn(r) + 1.0f;                // + conversions + 1 dfma
  const double z = r * r;
  const double z2 = z * z;
  const double rz = r * z;
  const double s01 = __fma_rn(z, 8.33333333332248946124e-03, -1.66666666666666324348e-01);
  const double s23 = __fma_rn(z, 2.75573137070700676789e-06, -1.98412698298579493134e-04);
  const double s45 = __fma_rn(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  const double z4 = z2 * z2;
  const double sa = __fma_rn(z2, s23, s01);
  const double sp = __fma_rn(z4, s45, sa);
  const double sn = __fma_rn(rz, sp, r);
  if (MODE == 2) return __double2float_rn(sn) + 1.0f;               // sin poly only, no quadrant
  const double c01 = __fma_rn(z, -1.38888888888741095749e-03, 4.16666666666666019037e-02);
  const double c23 = __fma_rn(z, -2.75573143513906633035e-07, 2.48015872894767294178e-05);
  const double c45 = __fma_rn(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  const double ca = __fma_rn(z2, c23, c01);
  const double cp = __fma_rn(z4, c45, ca);
  const double h = __fma_rn(-0.5, z, 1.0);
  const double cs = __fma_rn(z2, cp, h);
  const float sf = __double2float_rn(sn), cf = __double2float_rn(cs);
  if (MODE == 3) return sf + cf;                                    // both polys, no quadrant logic
  float so = (q & 1) ? cf : sf;
  float co = (q & 1) ? sf : cf;
  if (q & 2) so = -so;
  if ((q + 1) & 2) co = -co;
  if (MODE == 4) return so + 1.5f;                                  // + quadrant selects, sin used
  return so * 0.5f + co + 2.0f;                                     // both used
}
int main()
{
  run("reduction float only (3 fp32 + add)", [] __device__(float x) { return sc<0>(x); });
  run("+ 2 F2F in, DFMA, F2F out", [] __device__(float x) { return sc<1>(x); });
  run("+ sin Estrin (no quadrant)", [] __device__(float x) { return sc<2>(x); });
  run("+ cos Estrin, both converted", [] __device__(float x) { return sc<3>(x); });
  run("+ quadrant selects (sin used)", [] __device__(float x) { return sc<4>(x); });
  run("+ both used", [] __device__(float x) { return sc<5>(x); });
  run("rfm_sincos (range branch), s only", [] __device__(float x) { float s, c; rfm_sincos(x, &s, &c); return __fadd_rn(s, 1.5f); });
  run("F2F.F64.F32 + F2F.F32.F64", [] __device__(float x) { return __double2float_rn((double)x); });
  run("2x (F2F.F64.F32 + F2F.F32.F64)", [] __device__(float x) { return __double2float_rn((double)__double2float_rn((double)x)); });
  run("dmul chain x6 between cvt", [] __device__(float x) { double d = x; d = d * 1.0000001; d = d * 1.0000001; d = d * 1.0000001; d = d * 1.0000001; d = d * 1.0000001; d = d * 1.0000001; return __double2float_rn(d); });
  run("predicated branch (never taken) + fadd", [] __device__(float x) { if (x > 1e30f) x = sqrtf(x) * 3.f; return __fadd_rn(x, 1e-7f); });
  run("branch taken by half the lanes + fadd", [] __device__(float x) { if (threadIdx.x & 1) x = __fmul_rn(x, 1.0000001f); else x = __fadd_rn(x, 1e-7f); return x; });
  run("fsetp+fsel", [] __device__(float x) { return x > 0.5f ? __fadd_rn(x, 1e-7f) : __fadd_rn(x, 2e-7f); });
  run("fdiv_rn", [] __device__(float x) { return __fdiv_rn(1.3f, x); }, 1.1f);
  run("div: rcp.approx+newton (no checks)", [] __device__(float x) { float a = 1.3f, r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); float e = __fmaf_rn(-x, r, 1.0f); r = __fmaf_rn(r, e, r); float qq = __fmul_rn(a, r); float rem = __fmaf_rn(-x, qq, a); return __fmaf_rn(rem, r, qq); }, 1.1f);
  run("mufu.rcp + fadd", [] __device__(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return __fadd_rn(r, 0.4f); }, 1.1f);
  return 0;
}

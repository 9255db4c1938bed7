// rfm_probe.cu -- device-side evaluation of the scalar building blocks, for the parity tests (C ABI:
// rfm_math_probe / rfm_div_selftest).  Lets tests/ compare what the GPU computes for rfm_math.cuh / rfm_steps.cuh
// with libm and the oracle element by element, and pins the branch-free division against __fdiv_rn on the device.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/radiofm_b200.h"
#include "rfm_dsp.cuh"
#include "rfm_steps.cuh"

using namespace rfm;

namespace
{
__global__ void k_probe(int op, const float* a, const float* b, float* out, unsigned n)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const float x = a[i], y = b ? b[i] : 0.0f;
  float o0 = 0.0f, o1 = 0.0f;
  bool bad = false;
  switch (op)
  {
    case 0: rfm_sincos(x, &o0, &o1); break;
    case 1: rfm_sincos_core(x, &o0, &o1); break;
    case 2: rfm_sincos_generic(x, &o0, &o1); break;
    case 3: o0 = rfm_atan2f(x, y); break;                       // (y = a, x = b)
    case 4: o0 = rfm_atan2f_fast(x, y, bad); o1 = bad ? 1.0f : 0.0f; break;
    case 5: o0 = rfm_atan2f_generic(x, y); break;
    case 6: o0 = rfm_div_fast(x, y); o1 = rfm_div_unsafe(x, y) ? 1.0f : 0.0f; break;
    case 7: o0 = __fdiv_rn(x, y); break;
    case 8: o0 = rfm_wrap_demod_fast(x, bad); o1 = bad ? 1.0f : 0.0f; break;
    case 9: o0 = rfm_wrap_pilot_fast(x, bad); o1 = bad ? 1.0f : 0.0f; break;
    case 10: o0 = rfm_wrap_demod(x); o1 = rfm_wrap_pilot(x); break;
    case 11: o0 = arctan2_approx(x, y); break;
    case 12: o0 = rfm_fmodf_small(x, y); break;
    case 13: o0 = rfm_osc_gain(x); o1 = rfm_osc_gain_exact(x); break;
    default: break;
  }
  out[2 * i] = o0;
  out[2 * i + 1] = o1;
}

__device__ __forceinline__ uint32_t mix32(uint64_t z)
{
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (uint32_t)((z ^ (z >> 31)) >> 16);
}

// pairs per thread; operands: random mantissas, exponents inside the window checked by rfm_div_unsafe
__global__ void k_div_selftest(uint64_t seed, unsigned per_thread, unsigned long long* mismatches,
                               unsigned long long* tested)
{
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long bad = 0, cnt = 0;
  for (unsigned k = 0; k < per_thread; ++k)
  {
    const uint64_t c = seed + (tid * per_thread + k) * 0x9e3779b97f4a7c15ull;
    uint32_t ua = mix32(c), ub = mix32(c ^ 0xdeadbeefcafef00dull);
    if (k & 1)
    { // neighbouring exponents, dense mantissa patterns (hard cases for the final rounding)
      ub = (ub & 0x807fffffu) | (ua & 0x7f800000u);
      if (k & 2)
        ub = (ub & 0xff800000u) | ((ua & 0x007fffffu) ^ (1u << (k % 23)));
    }
    const float a = __uint_as_float(ua), b = __uint_as_float(ub);
    if (rfm_div_unsafe(a, b))
      continue;
    ++cnt;
    bad += __float_as_uint(rfm_div_fast(a, b)) != __float_as_uint(__fdiv_rn(a, b));
  }
  atomicAdd(mismatches, bad);
  atomicAdd(tested, cnt);
}
} // namespace

extern "C"
{

int rfm_math_probe(int op, const float* a, const float* b, float* out2, uint32_t n)
{
  if (!a || !out2)
    return RFM_ERR_INVALID;
  float *da = nullptr, *db = nullptr, *dout = nullptr;
  int rc = RFM_OK;
  if (cudaMalloc(&da, (size_t)n * 4) != cudaSuccess || cudaMalloc(&dout, (size_t)n * 8) != cudaSuccess ||
      (b && cudaMalloc(&db, (size_t)n * 4) != cudaSuccess))
    rc = RFM_ERR_CUDA;
  if (rc == RFM_OK)
  {
    cudaMemcpy(da, a, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (b)
      cudaMemcpy(db, b, (size_t)n * 4, cudaMemcpyHostToDevice);
    k_probe<<<(n + 255) / 256, 256>>>(op, da, db, dout, n);
    if (cudaMemcpy(out2, dout, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = RFM_ERR_CUDA;
  }
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return rc;
}

int rfm_div_selftest(uint64_t seed, uint64_t pairs, uint64_t* mismatches, uint64_t* tested)
{
  if (!mismatches || !tested)
    return RFM_ERR_INVALID;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess)
    return RFM_ERR_CUDA;
  cudaMemset(d, 0, 16);
  const unsigned threads = 256, blocks = 148 * 8, per_thread = (unsigned)(pairs / ((uint64_t)threads * blocks) + 1);
  k_div_selftest<<<blocks, threads>>>(seed, per_thread, d, d + 1);
  unsigned long long h[2] = {0, 0};
  const cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess)
    return RFM_ERR_CUDA;
  *mismatches = h[0];
  *tested = h[1];
  return RFM_OK;
}

} // extern "C"

// Dynamic quantisation kernels feeding the 8-bit tcgen05 GEMMs: per-row (per-token) abs-max scaling to int8 or fp8
// e4m3 in one pass (row kept in registers), optional SmoothQuant per-channel divisor folded in.
#include <cuda_fp8.h>

#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

// q[r, :] = round(x[r, :] / smooth[:] / scale[r]),  scale[r] = absmax(x[r, :] / smooth) / QMAX
template <typename T, bool kFp8>
__global__ void quantize_rows_kernel(const T* __restrict__ x, const float* __restrict__ smooth, uint8_t* __restrict__ q, float* __restrict__ scale,
                                     int cols) {
  __shared__ float scratch[33];
  const int row = blockIdx.x;
  const int nvec = cols >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * cols);
  constexpr int kV = 4;
  float v[kV][8];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      unpack8<T>(ld_stream(xr + vi), v[i]);
      if (smooth) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] /= smooth[vi * 8 + j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(v[i][j]));
    }
  }
  amax = block_max(amax, scratch);
  const float qmax = kFp8 ? 448.f : 127.f;
  const float s = amax > 0.f ? amax / qmax : 1.f;
  const float inv = 1.f / s;
  if (threadIdx.x == 0) scale[row] = s;
  uint2* qr = reinterpret_cast<uint2*>(q + (size_t)row * cols);
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      uint8_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[i][j] * inv;
        if (kFp8) o[j] = (uint8_t)__nv_cvt_float_to_fp8(t, __NV_SATFINITE, __NV_E4M3);
        else o[j] = (uint8_t)(int8_t)max(-127, min(127, __float2int_rn(t)));
      }
      qr[vi] = *reinterpret_cast<const uint2*>(o);
    }
  }
}

cudaError_t quantize_rows(const void* x, const float* smooth, void* q, float* scale, int rows, int cols, int dtype, bool fp8, cudaStream_t st) {
  if (cols % 8 || cols > 32768) return cudaErrorInvalidValue;
  if (!rows) return cudaSuccess;
  int threads = ((cols / 8 + 3) / 4 + 31) / 32 * 32;
  if (threads < 32) threads = 32;
#define PFX_Q(T, F) quantize_rows_kernel<T, F><<<rows, threads, 0, st>>>((const T*)x, smooth, (uint8_t*)q, scale, cols)
  if (dtype == 1) { if (fp8) PFX_Q(__nv_bfloat16, true); else PFX_Q(__nv_bfloat16, false); }
  else { if (fp8) PFX_Q(__half, true); else PFX_Q(__half, false); }
#undef PFX_Q
  return cudaGetLastError();
}

}  // namespace pfx

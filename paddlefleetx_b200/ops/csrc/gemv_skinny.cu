// Skinny GEMM for token-by-token decoding: y[m, n] = sum_k x[m, k] * W[n, k] (+ bias[n]),  m <= 8.
//
// With one to eight activation rows the tensor-core GEMM is idle 94+ % of a 128-row tile and, worse, covers the
// machine with only N/256 CTAs; the op is a pure weight stream.  This kernel is built for that: every warp owns
// kCols weight rows, streams them once with 16-byte no-allocate loads (two k-steps in flight per lane), keeps the
// few activation rows L1-resident, accumulates in fp32 and finishes with a shuffle reduction.  Decode GEMMs of the
// 6.7B model (N = 4096..16384, K = 4096..16384) then run at HBM speed instead of tile-quantised tensor-core speed.
// Reference: the decode path of GPTForGeneration (hybrid_model.py:1202-1339) issues cuBLAS GEMMs for these shapes.
#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

namespace {

template <typename T, int kRows, int kCols>
__global__ void __launch_bounds__(256) gemv_skinny_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ bias,
                                                          T* __restrict__ y, int N, int K) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n0 = warp * kCols;
  if (n0 >= N) return;
  float acc[kCols][kRows];
#pragma unroll
  for (int c = 0; c < kCols; ++c)
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[c][r] = 0.f;
  const uint4* wp[kCols];
#pragma unroll
  for (int c = 0; c < kCols; ++c) wp[c] = reinterpret_cast<const uint4*>(w + (size_t)min(n0 + c, N - 1) * K);
  const uint4* xp = reinterpret_cast<const uint4*>(x);
  const int kvec = K >> 3;                       // 8 elements per 16-byte vector
  for (int v0 = lane; v0 < kvec; v0 += 64) {
    const int v1 = v0 + 32;
    const bool has1 = v1 < kvec;
    uint4 wa[kCols], wb[kCols];
#pragma unroll
    for (int c = 0; c < kCols; ++c) {
      wa[c] = ld_stream(wp[c] + v0);
      wb[c] = has1 ? ld_stream(wp[c] + v1) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 1 && !has1) break;
      const int v = half ? v1 : v0;
      float xf[kRows][8];
#pragma unroll
      for (int r = 0; r < kRows; ++r) unpack8<T>(__ldg(xp + (size_t)r * kvec + v), xf[r]);
#pragma unroll
      for (int c = 0; c < kCols; ++c) {
        float wf[8];
        unpack8<T>(half ? wb[c] : wa[c], wf);
#pragma unroll
        for (int r = 0; r < kRows; ++r)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][r] = fmaf(wf[j], xf[r][j], acc[c][r]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < kCols; ++c)
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float v = acc[c][r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[c][r] = v;
    }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < kCols; ++c) {
      const int n = n0 + c;
      if (n >= N) break;
      const float b = bias ? to_f32<T>(bias[n]) : 0.f;
#pragma unroll
      for (int r = 0; r < kRows; ++r) y[(size_t)r * N + n] = from_f32<T>(acc[c][r] + b);
    }
  }
}

template <typename T, int kRows>
cudaError_t launch_rows(const T* x, const T* w, const T* bias, T* y, int N, int K, int num_sms, cudaStream_t st) {
  // enough warps to cover the machine twice over; fewer columns per warp when N is small
  const bool narrow = (N / 4 + 7) / 8 < 2 * num_sms;
  if (narrow) {
    const int warps = (N + 1) / 2;
    gemv_skinny_kernel<T, kRows, 2><<<(warps + 7) / 8, 256, 0, st>>>(x, w, bias, y, N, K);
  } else {
    const int warps = (N + 3) / 4;
    gemv_skinny_kernel<T, kRows, 4><<<(warps + 7) / 8, 256, 0, st>>>(x, w, bias, y, N, K);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t launch(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int num_sms, cudaStream_t st) {
  const T *xx = (const T*)x, *ww = (const T*)w, *bb = (const T*)bias;
  T* yy = (T*)y;
  switch (M) {
    case 1: return launch_rows<T, 1>(xx, ww, bb, yy, N, K, num_sms, st);
    case 2: return launch_rows<T, 2>(xx, ww, bb, yy, N, K, num_sms, st);
    case 3: return launch_rows<T, 3>(xx, ww, bb, yy, N, K, num_sms, st);
    case 4: return launch_rows<T, 4>(xx, ww, bb, yy, N, K, num_sms, st);
    case 5: return launch_rows<T, 5>(xx, ww, bb, yy, N, K, num_sms, st);
    case 6: return launch_rows<T, 6>(xx, ww, bb, yy, N, K, num_sms, st);
    case 7: return launch_rows<T, 7>(xx, ww, bb, yy, N, K, num_sms, st);
    case 8: return launch_rows<T, 8>(xx, ww, bb, yy, N, K, num_sms, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace

cudaError_t gemv_skinny(const void* x, const void* w, const void* bias, void* y, int M, int N, int K, int dtype, int num_sms, cudaStream_t st) {
  if (M < 1 || M > 8 || K % 8) return cudaErrorInvalidValue;
  if (dtype == 1) return launch<__nv_bfloat16>(x, w, bias, y, M, N, K, num_sms, st);
  if (dtype == 0) return launch<__half>(x, w, bias, y, M, N, K, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace pfx

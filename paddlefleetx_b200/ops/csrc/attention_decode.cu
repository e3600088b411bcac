// Single-query attention over a static KV cache (the decode step of generation).
//
//   out[b, h, :] = softmax(scale * q[b, h, :] . K[b, 0:L, h, :] + mask[b, 0:L]) @ V[b, 0:L, h, :]
//
// q: [B, 1, H, D], K / V: [B, Lmax, H, D] (the framework's cache layout, read in place — no gather, no transpose), mask: additive
// [B, Lmax] (0 = attend, large negative = padded / not yet written), D in {64, 128}.  One CTA per (head, batch): its warps take
// keys round-robin, every lane owns D/32 contiguous channels (one 4- or 8-byte load per key), the q.k dot product is a warp
// shuffle reduction and each warp keeps a private online-softmax state (m, l, acc); the warps' partial states are merged through
// shared memory at the end.  The op is a pure stream over the cache (2 * L * D * 2 bytes per head), so several keys are in flight
// per warp.  Replaces the library SDPA call on the decode path (reference: hybrid_model.py:303-346 core attention with cache).
#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

namespace {

constexpr int kDecWarps = 8;

template <typename T, int kPerLane>      // kPerLane = D / 32 channels per lane (2 or 4)
__global__ void __launch_bounds__(kDecWarps * 32) attention_decode_kernel(const T* __restrict__ q, T* __restrict__ k, T* __restrict__ v,
                                                                           const T* __restrict__ mask, T* __restrict__ out, int L, int Lmax,
                                                                           int H, float scale, const int64_t* __restrict__ write_idx) {
  constexpr int D = kPerLane * 32;
  __shared__ float s_m[kDecWarps], s_l[kDecWarps];
  __shared__ float s_acc[kDecWarps][D];
  const int h = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // packed mode (write_idx != null): q points at the fused QKV projection [B, 1, H, 3, D]; this CTA first appends its head's new
  // key / value at cache position *write_idx (the cache write used to be two separate index_copy launches), then attends
  const int q_head_stride = write_idx ? 3 * D : D;
  if (write_idx != nullptr) {
    const int64_t pos = write_idx[0];
    if (w < 2 && pos >= 0 && pos < Lmax) {
      const T* src = q + ((size_t)b * H + h) * 3 * D + (1 + w) * D;
      T* dst = (w == 0 ? k : v) + ((size_t)b * Lmax + pos) * H * D + (size_t)h * D;
      for (int i = lane; i < D; i += 32) dst[i] = src[i];
    }
    __syncthreads();
  }
  float qv[kPerLane];
  {
    const T* qp = q + ((size_t)b * H + h) * q_head_stride + lane * kPerLane;
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) qv[i] = to_f32<T>(qp[i]) * scale;
  }
  const size_t row_stride = (size_t)H * D;
  const T* kb = k + (size_t)b * Lmax * row_stride + (size_t)h * D + lane * kPerLane;
  const T* vb = v + (size_t)b * Lmax * row_stride + (size_t)h * D + lane * kPerLane;
  const T* mb = mask ? mask + (size_t)b * Lmax : nullptr;
  float m = -INFINITY, l = 0.f, acc[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) acc[i] = 0.f;
  constexpr int kU = 4;                                     // keys in flight per warp
  for (int t0 = w * kU; t0 < L; t0 += kDecWarps * kU) {
    float kk[kU][kPerLane], vv[kU][kPerLane];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = t0 + u;
      if (t < L) {
        if constexpr (kPerLane == 4) {
          const uint2 rk = *reinterpret_cast<const uint2*>(kb + (size_t)t * row_stride);
          const uint2 rv = *reinterpret_cast<const uint2*>(vb + (size_t)t * row_stride);
          const T* pk = reinterpret_cast<const T*>(&rk);
          const T* pv = reinterpret_cast<const T*>(&rv);
#pragma unroll
          for (int i = 0; i < 4; ++i) { kk[u][i] = to_f32<T>(pk[i]); vv[u][i] = to_f32<T>(pv[i]); }
        } else {
          const uint32_t rk = *reinterpret_cast<const uint32_t*>(kb + (size_t)t * row_stride);
          const uint32_t rv = *reinterpret_cast<const uint32_t*>(vb + (size_t)t * row_stride);
          const T* pk = reinterpret_cast<const T*>(&rk);
          const T* pv = reinterpret_cast<const T*>(&rv);
#pragma unroll
          for (int i = 0; i < 2; ++i) { kk[u][i] = to_f32<T>(pk[i]); vv[u][i] = to_f32<T>(pv[i]); }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = t0 + u;
      if (t >= L) break;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < kPerLane; ++i) s = fmaf(qv[i], kk[u][i], s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (mb) s += to_f32<T>(mb[t]);
      const float m_new = fmaxf(m, s);
      const float corr = __expf(m - m_new), p = __expf(s - m_new);
      l = l * corr + p;
#pragma unroll
      for (int i = 0; i < kPerLane; ++i) acc[i] = acc[i] * corr + p * vv[u][i];
      m = m_new;
    }
  }
  if (lane == 0) { s_m[w] = m; s_l[w] = l; }
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) s_acc[w][lane * kPerLane + i] = acc[i];
  __syncthreads();
  // merge the warps' online-softmax states; thread c finishes channel c
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float mm = -INFINITY;
#pragma unroll
    for (int i = 0; i < kDecWarps; ++i) mm = fmaxf(mm, s_m[i]);
    float ll = 0.f, a = 0.f;
#pragma unroll
    for (int i = 0; i < kDecWarps; ++i) {
      const float f = (s_m[i] == -INFINITY) ? 0.f : __expf(s_m[i] - mm);
      ll += s_l[i] * f;
      a += s_acc[i][c] * f;
    }
    out[((size_t)b * H + h) * D + c] = from_f32<T>(ll > 0.f ? a / ll : 0.f);
  }
}

}  // namespace

cudaError_t attention_decode(const void* q, void* k, void* v, const void* mask, void* out, int B, int H, int D, int L, int Lmax,
                             float scale, int dtype, cudaStream_t st, const int64_t* write_idx) {
  if (!B || !H) return cudaSuccess;
  if ((D != 64 && D != 128) || L < 1 || L > Lmax) return cudaErrorInvalidValue;
  const dim3 grid(H, B), block(kDecWarps * 32);
#define PFX_AD(T, P) attention_decode_kernel<T, P><<<grid, block, 0, st>>>((const T*)q, (T*)k, (T*)v, (const T*)mask, (T*)out, L, Lmax, H, scale, write_idx)
  if (dtype == 1) { if (D == 128) PFX_AD(__nv_bfloat16, 4); else PFX_AD(__nv_bfloat16, 2); }
  else if (dtype == 0) { if (D == 128) PFX_AD(__half, 4); else PFX_AD(__half, 2); }
  else return cudaErrorInvalidValue;
#undef PFX_AD
  return cudaGetLastError();
}

}  // namespace pfx

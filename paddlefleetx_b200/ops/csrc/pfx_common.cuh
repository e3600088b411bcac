// Shared device helpers for the bandwidth-bound kernels: 128-bit vector IO, warp/block reductions,
// stateless Philox4x32-10.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pfx {

template <typename T> struct Vec8 { uint4 raw; };

template <typename T> __device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 v = __bfloat1622float2(p[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
template <> __device__ __forceinline__ void unpack8<__half>(const uint4& raw, float (&f)[8]) {
  const __half2* p = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 v = __half22float2(p[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
template <typename T> __device__ __forceinline__ uint4 pack8(const float (&f)[8]);
template <> __device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&f)[8]) {
  uint4 raw; __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return raw;
}
template <> __device__ __forceinline__ uint4 pack8<__half>(const float (&f)[8]) {
  uint4 raw; __half2* p = reinterpret_cast<__half2*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return raw;
}

// Generic scalar converters (used by the fp32 instantiations and tails)
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum broadcast to all threads; `scratch` needs 33 floats
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float t = (lane < nw) ? scratch[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  float t = (lane < nw) ? scratch[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// ---------------------------------------------------------------- Philox4x32-10 (counter-based)
struct Philox {
  __device__ __forceinline__ static uint4 gen(uint64_t seed, uint64_t counter) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
  // 8 Bernoulli(keep) decisions for elements [8*vec_idx, 8*vec_idx+8): one Philox call, 16 bits each
  __device__ __forceinline__ static void keep8(uint64_t seed, uint64_t offset, uint64_t vec_idx, uint32_t thresh16, bool (&keep)[8]) {
    const uint4 r = gen(seed, offset + vec_idx);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      keep[2 * i] = (w[i] & 0xFFFFu) >= thresh16;
      keep[2 * i + 1] = (w[i] >> 16) >= thresh16;
    }
  }
  __device__ __forceinline__ static float uniform(uint64_t seed, uint64_t counter) {
    return (gen(seed, counter).x >> 8) * (1.0f / 16777216.0f);
  }
};

// 4-element packets: 16-byte accesses for fp32, 8-byte for bf16/fp16
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&f)[4]) {
  if constexpr (sizeof(T) == 4) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    f[0] = r.x; f[1] = r.y; f[2] = r.z; f[3] = r.w;
  } else {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const T* h = reinterpret_cast<const T*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = to_f32<T>(h[j]);
  }
}
template <typename T> __device__ __forceinline__ void store4(T* p, const float (&f)[4]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  } else {
    uint2 r;
    T* h = reinterpret_cast<T*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = from_f32<T>(f[j]);
    *reinterpret_cast<uint2*>(p) = r;
  }
}


}  // namespace pfx

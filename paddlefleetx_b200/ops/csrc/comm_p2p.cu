// Collectives written directly against peer memory (CUDA-IPC mapped buffers over NVLink 5 / NVSwitch).
// On an NVSwitch box every peer is one hop away at full bandwidth, so the kernels are flat (no rings,
// no trees):
//   * p2p_barrier        — release/acquire flag exchange through per-rank signal pads
//   * p2p_reduce_scatter — ZeRO gradient path: each rank PULLS its shard from every peer's bucket, reduces in
//                          fp32 in registers and writes (or accumulates into) its fp32/bf16 main-grad shard:
//                          cast + scale + accumulate + reduce-scatter in one pass, deterministic order
//   * p2p_all_gather     — each rank PUSHES its shard into every peer's buffer (posted stores)
//   * adamw_p2p_broadcast— the ZeRO-1/2 "update then broadcast params" step as ONE kernel: the owner updates
//                          its fp32 master shard and stores the new bf16 weights straight into all peers'
//                          parameter buffers
// Reference: C-ZeRO1 / C-DP call sites in SURVEY §2.6 (NCCL reduce / broadcast issued by Paddle's sharding
// optimizer as separate launches).
#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

constexpr int kMaxPeers = 16;
struct PeerPtrs { void* p[kMaxPeers]; };

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

// signal pad layout: pad[slot * kMaxPeers + src_rank].  put = CAS 0->1 on the peer, wait = CAS 1->0 locally.
__global__ void p2p_barrier_kernel(PeerPtrs pads, int rank, int world, uint32_t slot) {
  const int peer = threadIdx.x;
  if (peer >= world || peer == rank) return;
  uint32_t* remote = reinterpret_cast<uint32_t*>(pads.p[peer]) + slot * kMaxPeers + rank;
  uint32_t* local = reinterpret_cast<uint32_t*>(pads.p[rank]) + slot * kMaxPeers + peer;
  long long spins = 0;
  while (cas_release_sys(remote, 0u, 1u) != 0u) { if (++spins > (1ll << 31)) __trap(); }
  spins = 0;
  while (cas_acquire_sys(local, 1u, 0u) != 1u) { if (++spins > (1ll << 31)) __trap(); }
}

cudaError_t p2p_barrier(uint32_t** signal_pads, int rank, int world, uint32_t slot, cudaStream_t st) {
  if (world > kMaxPeers) return cudaErrorInvalidValue;
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) pp.p[i] = signal_pads[i];
  p2p_barrier_kernel<<<1, 32, 0, st>>>(pp, rank, world, slot);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ pull reduce-scatter
template <typename TIn, typename TOut, int kWorld>
__global__ void __launch_bounds__(512) p2p_reduce_scatter_kernel(PeerPtrs bufs, TOut* __restrict__ out, size_t shard_elems, int rank, bool accumulate,
                                                                 float scale) {
  constexpr int kUnroll = (kWorld <= 2) ? 4 : 2;     // independent 16 B loads per peer in flight per thread
  const size_t nvec = shard_elems >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4* src[kWorld];
#pragma unroll
  for (int p = 0; p < kWorld; ++p)
    src[p] = reinterpret_cast<const uint4*>(reinterpret_cast<const TIn*>(bufs.p[(rank + p) % kWorld]) + (size_t)rank * shard_elems);
  for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kUnroll) {
    uint4 raw[kUnroll][kWorld];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i < nvec) {
#pragma unroll
        for (int p = 0; p < kWorld; ++p) raw[u][p] = ld_stream(src[p] + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t i = i0 + u * stride;
      if (i >= nvec) continue;
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int p = 0; p < kWorld; ++p) {   // fixed (rank-relative) order -> bitwise identical sums on every run
        float v[8];
        unpack8<TIn>(raw[u][p], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
      TOut* o = out + i * 8;
      if (accumulate) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = from_f32<TOut>(acc[j] * scale + to_f32<TOut>(o[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = from_f32<TOut>(acc[j] * scale);
      }
    }
  }
}

template <typename TIn, typename TOut>
static cudaError_t rs_launch(const PeerPtrs& pp, void* out, size_t shard, int rank, int world, bool acc, float scale, int grid, cudaStream_t st) {
  switch (world) {
    case 2: p2p_reduce_scatter_kernel<TIn, TOut, 2><<<grid, 512, 0, st>>>(pp, (TOut*)out, shard, rank, acc, scale); break;
    case 4: p2p_reduce_scatter_kernel<TIn, TOut, 4><<<grid, 512, 0, st>>>(pp, (TOut*)out, shard, rank, acc, scale); break;
    case 8: p2p_reduce_scatter_kernel<TIn, TOut, 8><<<grid, 512, 0, st>>>(pp, (TOut*)out, shard, rank, acc, scale); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t p2p_reduce_scatter(void** peer_bufs, void* out, size_t shard_elems, int rank, int world, int in_dtype, int out_dtype,
                               bool accumulate, float scale, int num_ctas, cudaStream_t st) {
  if (shard_elems % 8 || world > kMaxPeers) return cudaErrorInvalidValue;
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) pp.p[i] = peer_bufs[i];
  if (in_dtype == 1 && out_dtype == 3) return rs_launch<__nv_bfloat16, float>(pp, out, shard_elems, rank, world, accumulate, scale, num_ctas, st);
  if (in_dtype == 1 && out_dtype == 1) return rs_launch<__nv_bfloat16, __nv_bfloat16>(pp, out, shard_elems, rank, world, accumulate, scale, num_ctas, st);
  if (in_dtype == 0 && out_dtype == 3) return rs_launch<__half, float>(pp, out, shard_elems, rank, world, accumulate, scale, num_ctas, st);
  if (in_dtype == 0 && out_dtype == 0) return rs_launch<__half, __half>(pp, out, shard_elems, rank, world, accumulate, scale, num_ctas, st);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ local slot reduce (second half of GEMM -> reduce-scatter)
// out[i] = sum_s staging[s][i] (+ bias[col]); every slot was written by a different rank's GEMM epilogue.
template <typename T>
__global__ void slot_reduce_kernel(const T* __restrict__ staging, T* __restrict__ out, const T* __restrict__ bias, size_t slot_elems, int world,
                                   int cols) {
  const size_t nvec = slot_elems >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = 0; s < world; ++s) {
      float v[8];
      unpack8<T>(ld_stream(reinterpret_cast<const uint4*>(staging + (size_t)s * slot_elems) + i), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    if (bias) {
      float b[8];
      unpack8<T>(__ldg(reinterpret_cast<const uint4*>(bias) + (i % (size_t)(cols >> 3))), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += b[j];
    }
    st_stream(reinterpret_cast<uint4*>(out) + i, pack8<T>(acc));
  }
}

cudaError_t slot_reduce(const void* staging, void* out, const void* bias, size_t slot_elems, int world, int cols, int dtype, int num_sms,
                        cudaStream_t st) {
  if (slot_elems % 8 || cols % 8) return cudaErrorInvalidValue;
  const int threads = 256;
  size_t g = (slot_elems / 8 + threads - 1) / threads;
  const size_t cap = (size_t)num_sms * 8;
  const int grid = (int)(g < cap ? (g ? g : 1) : cap);
  if (dtype == 1) slot_reduce_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>((const __nv_bfloat16*)staging, (__nv_bfloat16*)out, (const __nv_bfloat16*)bias, slot_elems, world, cols);
  else slot_reduce_kernel<__half><<<grid, threads, 0, st>>>((const __half*)staging, (__half*)out, (const __half*)bias, slot_elems, world, cols);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ push all-gather
__global__ void p2p_all_gather_kernel(PeerPtrs bufs, const uint4* __restrict__ src, size_t nvec, size_t dst_vec_offset, int rank, int world) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 v = ld_stream(src + i);
    for (int p = 0; p < world; ++p) {
      const int dst = (rank + p) % world;
      st_stream(reinterpret_cast<uint4*>(bufs.p[dst]) + dst_vec_offset + i, v);
    }
  }
}

cudaError_t p2p_all_gather(void** peer_bufs, const void* src, size_t shard_elems, int rank, int world, int dtype, int num_ctas, cudaStream_t st) {
  const size_t esize = dtype == 3 ? 4 : 2;
  const size_t bytes = shard_elems * esize;
  if (bytes % 16 || world > kMaxPeers) return cudaErrorInvalidValue;
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) pp.p[i] = peer_bufs[i];
  p2p_all_gather_kernel<<<num_ctas, 512, 0, st>>>(pp, (const uint4*)src, bytes / 16, (size_t)rank * (bytes / 16), rank, world);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ AdamW + parameter broadcast in one kernel
template <typename TG, typename TP>
__global__ void adamw_p2p_kernel(PeerPtrs params, size_t shard_offset, float* __restrict__ master, const TG* __restrict__ grad,
                                 float* __restrict__ m, float* __restrict__ v, size_t n, float lr, float beta1, float beta2, float eps,
                                 float wd, float bc1, float bc2, const float* __restrict__ gscale, const float* __restrict__ found_inf,
                                 int world) {
  if (found_inf && found_inf[0] != 0.f) return;
  const float gs = gscale ? gscale[0] : 1.f;
  const float step_size = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd, omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  const size_t nvec = n >> 3;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float g[8], w8[8], mm[8], vv[8];
    // all loads first (7 independent 16-byte / 8-byte requests per thread in flight), then math, then stores
    load4<TG>(grad + i * 8, *reinterpret_cast<float(*)[4]>(&g[0]));
    load4<TG>(grad + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&g[4]));
    load4<float>(master + i * 8, *reinterpret_cast<float(*)[4]>(&w8[0]));
    load4<float>(master + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&w8[4]));
    load4<float>(m + i * 8, *reinterpret_cast<float(*)[4]>(&mm[0]));
    load4<float>(m + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&mm[4]));
    load4<float>(v + i * 8, *reinterpret_cast<float(*)[4]>(&vv[0]));
    load4<float>(v + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&vv[4]));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = g[j] * gs;
      mm[j] = beta1 * mm[j] + omb1 * gj;
      vv[j] = beta2 * vv[j] + omb2 * gj * gj;
      w8[j] = w8[j] * decay - step_size * mm[j] / (sqrtf(vv[j]) * inv_sqrt_bc2 + eps);
    }
    store4<float>(m + i * 8, *reinterpret_cast<float(*)[4]>(&mm[0]));
    store4<float>(m + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&mm[4]));
    store4<float>(v + i * 8, *reinterpret_cast<float(*)[4]>(&vv[0]));
    store4<float>(v + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&vv[4]));
    store4<float>(master + i * 8, *reinterpret_cast<float(*)[4]>(&w8[0]));
    store4<float>(master + i * 8 + 4, *reinterpret_cast<float(*)[4]>(&w8[4]));
    const uint4 packed = pack8<TP>(w8);
    for (int p = 0; p < world; ++p)
      st_stream(reinterpret_cast<uint4*>(reinterpret_cast<TP*>(params.p[p]) + shard_offset) + i, packed);
  }
}

cudaError_t adamw_p2p_broadcast(void** peer_param_bufs, size_t shard_offset, float* master, const void* grad, float* m, float* v, size_t n,
                                float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2, const float* gscale,
                                const float* found_inf, int grad_dtype, int lp_dtype, int world, int num_ctas, cudaStream_t st) {
  if (n % 8 || shard_offset % 8 || world > kMaxPeers) return cudaErrorInvalidValue;
  if (!n) return cudaSuccess;
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) pp.p[i] = peer_param_bufs[i];
#define PFX_AP(TG, TP) adamw_p2p_kernel<TG, TP><<<num_ctas, 256, 0, st>>>(pp, shard_offset, master, (const TG*)grad, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2, gscale, found_inf, world)
  if (grad_dtype == 3 && lp_dtype == 1) PFX_AP(float, __nv_bfloat16);
  else if (grad_dtype == 3 && lp_dtype == 0) PFX_AP(float, __half);
  else if (grad_dtype == 1 && lp_dtype == 1) PFX_AP(__nv_bfloat16, __nv_bfloat16);
  else if (grad_dtype == 0 && lp_dtype == 0) PFX_AP(__half, __half);
  else return cudaErrorInvalidValue;
#undef PFX_AP
  return cudaGetLastError();
}

}  // namespace pfx

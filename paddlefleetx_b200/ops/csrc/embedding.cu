// Embedding lookup and its gradient (reference call sites GPTEmbeddings / VocabParallelEmbedding, hybrid_model.py:699-736).
//
//   forward   out[t, :] = W[ids[t] - vocab_start, :]  (zero row for ids owned by another vocab shard)  (+ P[pos[t], :])
//             word + position look-up and their sum in ONE pass (no intermediate tensors, no separate add kernel)
//   backward  deterministic scatter-add without atomics: the (row, token) pairs are sorted by a single-CTA bitonic network in shared
//             memory (up to 16 K tokens per call — one micro-batch), every run of equal rows is summed in fp32 in token order by the
//             CTAs that own its first element, and the sum is stored or accumulated ONCE into the weight-gradient row.  With the flat
//             optimizer's main_grad behind the weight (tied LM head already wrote it) only the rows that occur are touched: no dense
//             [vocab, hidden] gradient is materialised.
#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

namespace {

template <typename T>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ w, const int64_t* __restrict__ pos,
                                                        const T* __restrict__ pw, T* __restrict__ out, int hidden, int64_t vocab_start, int64_t rows) {
  const int64_t t = blockIdx.x;
  const int64_t r = ids[t] - vocab_start;
  const bool own = r >= 0 && r < rows;
  const uint4* src = reinterpret_cast<const uint4*>(w + (own ? r : 0) * hidden);
  const uint4* psrc = pw ? reinterpret_cast<const uint4*>(pw + pos[t] * hidden) : nullptr;
  uint4* dst = reinterpret_cast<uint4*>(out + t * hidden);
  const int nvec = hidden >> 3;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (own) unpack8<T>(__ldg(src + i), a);
    if (psrc) {
      float b[8];
      unpack8<T>(__ldg(psrc + i), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += b[j];
    }
    st_stream(dst + i, pack8<T>(a));
  }
}

// keys = (row << 32 | token) for owned rows, all-ones for foreign / padding entries; sorted ascending in shared memory
__global__ void __launch_bounds__(1024) embed_sort_kernel(const int64_t* __restrict__ ids, int n, int n_pow2, int64_t vocab_start, int64_t rows,
                                                          unsigned long long* __restrict__ sorted) {
  extern __shared__ unsigned long long keys[];
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      const int64_t r = ids[i] - vocab_start;
      if (r >= 0 && r < rows) k = ((unsigned long long)r << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) sorted[i] = keys[i];
}

// grid (n_pow2, column chunks): the CTA at the head of a run of equal rows sums the run (token order) and writes the gradient row once
template <typename T, typename TG>
__global__ void __launch_bounds__(128) embed_bwd_kernel(const unsigned long long* __restrict__ sorted, int n_pow2, const T* __restrict__ dout,
                                                        TG* __restrict__ dw, int hidden, bool accumulate) {
  const int i = blockIdx.x;
  const unsigned long long key = sorted[i];
  if (key == ~0ull) return;
  const unsigned row = (unsigned)(key >> 32);
  if (i > 0 && (unsigned)(sorted[i - 1] >> 32) == row && sorted[i - 1] != ~0ull) return;       // not the head of its run
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c >= hidden) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int e = i; e < n_pow2; ++e) {
    const unsigned long long ke = sorted[e];
    if (ke == ~0ull || (unsigned)(ke >> 32) != row) break;
    float v[8];
    unpack8<T>(ld_stream(reinterpret_cast<const uint4*>(dout + (size_t)(unsigned)ke * hidden + c)), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  TG* dst = dw + (size_t)row * hidden + c;
  if constexpr (sizeof(TG) == 4) {
    float4 o0 = make_float4(acc[0], acc[1], acc[2], acc[3]), o1 = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (accumulate) {
      const float4 p0 = *reinterpret_cast<const float4*>(dst), p1 = *reinterpret_cast<const float4*>(dst + 4);
      o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w; o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
    }
    *reinterpret_cast<float4*>(dst) = o0;
    *reinterpret_cast<float4*>(dst + 4) = o1;
  } else {
    if (accumulate) {
      float p[8];
      unpack8<TG>(*reinterpret_cast<const uint4*>(dst), p);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += p[j];
    }
    *reinterpret_cast<uint4*>(dst) = pack8<TG>(acc);
  }
}

}  // namespace

cudaError_t embedding_fwd(const int64_t* ids, const void* w, const int64_t* pos, const void* pw, void* out, int64_t tokens, int hidden,
                          int64_t vocab_start, int64_t rows, int dtype, cudaStream_t st) {
  if (hidden % 8 || tokens <= 0) return tokens == 0 ? cudaSuccess : cudaErrorInvalidValue;
  if (dtype == 1) embed_fwd_kernel<__nv_bfloat16><<<(unsigned)tokens, 256, 0, st>>>(ids, (const __nv_bfloat16*)w, pos, (const __nv_bfloat16*)pw, (__nv_bfloat16*)out, hidden, vocab_start, rows);
  else if (dtype == 0) embed_fwd_kernel<__half><<<(unsigned)tokens, 256, 0, st>>>(ids, (const __half*)w, pos, (const __half*)pw, (__half*)out, hidden, vocab_start, rows);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

int embedding_bwd_max_tokens() { return 16384; }

cudaError_t embedding_bwd(const int64_t* ids, const void* dout, void* dw, unsigned long long* workspace, int64_t tokens, int hidden,
                          int64_t vocab_start, int64_t rows, int dtype, int grad_dtype, bool accumulate, cudaStream_t st) {
  if (hidden % 8 || tokens > embedding_bwd_max_tokens() || tokens < 0) return cudaErrorInvalidValue;
  if (tokens == 0) return cudaSuccess;
  int n_pow2 = 2;
  while (n_pow2 < tokens) n_pow2 <<= 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(embed_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, embedding_bwd_max_tokens() * 8);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  embed_sort_kernel<<<1, 1024, (size_t)n_pow2 * 8, st>>>(ids, (int)tokens, n_pow2, vocab_start, rows, workspace);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const dim3 grid((unsigned)n_pow2, (unsigned)((hidden / 8 + 127) / 128));
#define PFX_EB(T, TG) embed_bwd_kernel<T, TG><<<grid, 128, 0, st>>>(workspace, n_pow2, (const T*)dout, (TG*)dw, hidden, accumulate)
  if (dtype == 1 && grad_dtype == 1) PFX_EB(__nv_bfloat16, __nv_bfloat16);
  else if (dtype == 1 && grad_dtype == 3) PFX_EB(__nv_bfloat16, float);
  else if (dtype == 0 && grad_dtype == 0) PFX_EB(__half, __half);
  else if (dtype == 0 && grad_dtype == 3) PFX_EB(__half, float);
  else return cudaErrorInvalidValue;
#undef PFX_EB
  return cudaGetLastError();
}

}  // namespace pfx

// Expert-parallel MoE data movement written against peer memory (CUDA-IPC over NVLink 5 / NVSwitch).
//
// The reference (ppfleetx/models/language_model/moe/moe_layer.py:33-235, SURVEY C-EP) moves tokens with two
// variable-size NCCL all-to-alls per direction (global_scatter / global_gather) whose split sizes are read back
// to the host, plus index_select / scatter passes around them.  On an NVSwitch box every expert owner is one hop
// away, so here the all-to-all IS the gather/scatter:
//
//   moe_route     local, deterministic: rank of every (token, k) slot inside its expert + per-expert counts
//   moe_dispatch  ONE kernel: push my count row to every peer, wait for theirs, derive the compact, 128-row
//                 aligned expert-major layout of every destination on the device, then store each token row
//                 (optionally scaled: the combine backward) straight into the owner's receive buffer; ends with a
//                 flag barrier so the expert GEMMs that follow on the stream see complete inputs
//   moe_combine   ONE kernel: flag barrier, then pull the k expert-output rows of every token from their owners,
//                 gate-weight and sum them in fp32 (optionally keeping the pulled rows for the gate gradient)
//
// No host round trip is needed for the exchange itself; the host reads the small segment table only to size the
// expert loop.  Backward reuses the same two kernels with roles swapped (combine-bwd = dispatch of w*dy,
// dispatch-bwd = unweighted combine).
#include <cstdio>

#include "pfx_common.cuh"
#include "pfx_kernels.h"

namespace pfx {

namespace {

constexpr int kMoeMaxPeers = 16;
constexpr int kMoeMaxExperts = 1024;      // world * experts-per-rank
struct MoePeers { void* p[kMoeMaxPeers]; };

__device__ __forceinline__ uint32_t moe_ld_acquire(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void moe_st_release(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void moe_wait_flag(const uint32_t* p, uint32_t epoch) {
  long long t0 = clock64();
  while ((int32_t)(moe_ld_acquire(p) - epoch) < 0) {      // epochs only grow
    if (clock64() - t0 > 4000000000ll) {
      printf("pfx moe: flag wait timed out (block %d, epoch %u, have %u)\n", blockIdx.x, epoch, moe_ld_acquire(p));
      __trap();
    }
  }
}

// ------------------------------------------------------------------ routing (local)
// grid = total experts; CTA g walks the slots in order and numbers those routed to expert g.
__global__ void __launch_bounds__(256) moe_route_kernel(const int64_t* __restrict__ gate_idx, int num_slots, int* __restrict__ slot_rank,
                                                        int* __restrict__ counts) {
  __shared__ int warp_tot[8];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int base = 0;
  for (int s0 = 0; s0 < num_slots; s0 += 256) {
    const int s = s0 + tid;
    const bool mine = s < num_slots && gate_idx[s] == (int64_t)g;
    const unsigned b = __ballot_sync(0xffffffffu, mine);
    const int prefix = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int c = warp_tot[w];
      if (w < warp) woff += c;
      tot += c;
    }
    if (mine) slot_rank[s] = base + woff + prefix;
    base += tot;
    __syncthreads();
  }
  if (tid == 0) counts[g] = base;
}

// ------------------------------------------------------------------ dispatch (push)
struct DispatchArgs {
  const void* src;            // [num_src_rows, H]
  const float* scale;         // [num_slots] or null
  const int64_t* gate_idx;    // [num_slots], -1 = dropped
  const int* slot_rank;       // [num_slots]
  const int* counts;          // [world * e_local] my outgoing counts
  int* slot_loc;              // out [num_slots]: (dst << 24) | row, -1 = dropped
  int* seg;                   // out [2 * e_local + 2]: start[e], count[e], total aligned rows, overflow flag
  MoePeers recv;              // receive buffers [cap_rows, H]
  MoePeers cnt;               // count matrices [world][world * e_local] (row = source rank)
  MoePeers flags;             // uint32 [2][16]: arrive[src], done[src]
  unsigned* block_counter;
  int num_slots, src_div, H, e_local, world, rank, align, cap_rows;
  uint32_t epoch;
};

template <typename T>
__global__ void __launch_bounds__(256) moe_dispatch_kernel(DispatchArgs a) {
  extern __shared__ int sm[];
  const int e_total = a.world * a.e_local;
  int* cnt_s = sm;                       // [world][e_total]
  int* base_s = sm + a.world * e_total;  // [e_total] first row of my block inside expert g's segment on its owner
  __shared__ int is_last;
  const int tid = threadIdx.x;
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.flags.p[a.rank]);

  // phase 0: count exchange (CTA 0 publishes, everybody waits)
  if (blockIdx.x == 0) {
    for (int i = tid; i < a.world * e_total; i += blockDim.x) {
      const int peer = i / e_total, g = i % e_total;
      reinterpret_cast<int*>(a.cnt.p[peer])[a.rank * e_total + g] = a.counts[g];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < a.world) moe_st_release(reinterpret_cast<uint32_t*>(a.flags.p[tid]) + a.rank, a.epoch);
  }
  if (tid < a.world) moe_wait_flag(my_flags + tid, a.epoch);
  __syncthreads();
  const int* my_cnt = reinterpret_cast<const int*>(a.cnt.p[a.rank]);
  for (int i = tid; i < a.world * e_total; i += blockDim.x) cnt_s[i] = __ldcv(my_cnt + i);
  __syncthreads();

  // phase 1: layout of every destination (expert-major, segments aligned to `align` rows)
  for (int g = tid; g < e_total; g += blockDim.x) {
    const int dst = g / a.e_local, e = g % a.e_local;
    int start = 0;
    for (int e2 = 0; e2 < e; ++e2) {
      int tot = 0;
      for (int s = 0; s < a.world; ++s) tot += cnt_s[s * e_total + dst * a.e_local + e2];
      start += (tot + a.align - 1) / a.align * a.align;
    }
    int before = 0;
    for (int s = 0; s < a.rank; ++s) before += cnt_s[s * e_total + g];
    base_s[g] = start + before;
  }
  __syncthreads();
  // my own experts: segment table + zero the alignment padding of my receive buffer
  {
    int start = 0;
    T* mine = reinterpret_cast<T*>(a.recv.p[a.rank]);
    const int vec_per_row = a.H / 8;
    for (int e = 0; e < a.e_local; ++e) {
      int tot = 0;
      for (int s = 0; s < a.world; ++s) tot += cnt_s[s * e_total + a.rank * a.e_local + e];
      const int padded = (tot + a.align - 1) / a.align * a.align;
      if (blockIdx.x == 0 && tid == 0) { a.seg[e] = start; a.seg[a.e_local + e] = tot; }
      const int pad_lo = min(start + tot, a.cap_rows), pad_hi = min(start + padded, a.cap_rows);
      const long long nvec = (long long)(pad_hi - pad_lo) * vec_per_row;
      uint4* z = reinterpret_cast<uint4*>(mine + (size_t)pad_lo * a.H);
      for (long long i = (long long)blockIdx.x * blockDim.x + tid; i < nvec; i += (long long)gridDim.x * blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
      start += padded;
    }
    if (blockIdx.x == 0 && tid == 0) { a.seg[2 * a.e_local] = start; a.seg[2 * a.e_local + 1] = start > a.cap_rows ? 1 : 0; }
  }

  // phase 2: one warp per slot, 16-byte posted stores into the owner's buffer
  const int lane = tid & 31, warps_per_cta = blockDim.x >> 5;
  const int vec_per_row = a.H / 8;
  for (int s = blockIdx.x * warps_per_cta + (tid >> 5); s < a.num_slots; s += gridDim.x * warps_per_cta) {
    const int64_t g = a.gate_idx[s];
    int loc = -1;
    if (g >= 0) {
      const int row = base_s[g] + a.slot_rank[s];
      const int dst = (int)g / a.e_local;
      if (row < a.cap_rows) {
        loc = (dst << 24) | row;
        const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.src) + (size_t)(s / a.src_div) * a.H);
        uint4* dp = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.recv.p[dst]) + (size_t)row * a.H);
        if (a.scale == nullptr) {
          for (int v = lane; v < vec_per_row; v += 32) dp[v] = ld_stream(sp + v);
        } else {
          const float sc = a.scale[s];
          for (int v = lane; v < vec_per_row; v += 32) {
            float f[8];
            unpack8<T>(ld_stream(sp + v), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] *= sc;
            dp[v] = pack8<T>(f);
          }
        }
      }
    }
    if (lane == 0) a.slot_loc[s] = loc;
  }

  // phase 3: the last CTA to finish tells every peer "my rows have landed" and waits for theirs
  __threadfence_system();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(a.block_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    if (tid == 0) *a.block_counter = 0;
    if (tid < a.world) {
      moe_st_release(reinterpret_cast<uint32_t*>(a.flags.p[tid]) + kMoeMaxPeers + a.rank, a.epoch);
      moe_wait_flag(my_flags + kMoeMaxPeers + tid, a.epoch);
    }
  }
}

// ------------------------------------------------------------------ combine (pull)
struct CombineArgs {
  MoePeers src;               // peers' expert-output staging [cap_rows, H]
  const int* slot_loc;        // [T * topk]
  const float* weights;       // [T * topk] or null (= 1)
  void* out;                  // [T, H]
  void* rows;                 // [T * topk, H] or null: pulled rows kept for the gate gradient
  MoePeers flags;             // uint32 [3][16]: third row = ready[src]
  int T, topk, H, world, rank;
  uint32_t epoch;
};

template <typename T>
__global__ void __launch_bounds__(256) moe_combine_kernel(CombineArgs a) {
  const int tid = threadIdx.x;
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(a.flags.p[a.rank]);
  if (blockIdx.x == 0 && tid < a.world) moe_st_release(reinterpret_cast<uint32_t*>(a.flags.p[tid]) + 2 * kMoeMaxPeers + a.rank, a.epoch);
  if (tid < a.world) moe_wait_flag(my_flags + 2 * kMoeMaxPeers + tid, a.epoch);
  __syncthreads();
  const int lane = tid & 31, warps_per_cta = blockDim.x >> 5;
  const int vec_per_row = a.H / 8;
  for (int t = blockIdx.x * warps_per_cta + (tid >> 5); t < a.T; t += gridDim.x * warps_per_cta) {
    for (int v = lane; v < vec_per_row; v += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k = 0; k < a.topk; ++k) {
        const int s = t * a.topk + k;
        const int loc = a.slot_loc[s];
        uint4 raw = make_uint4(0, 0, 0, 0);
        if (loc >= 0) {
          const T* base = reinterpret_cast<const T*>(a.src.p[loc >> 24]) + (size_t)(loc & 0xFFFFFF) * a.H;
          raw = ld_stream(reinterpret_cast<const uint4*>(base) + v);
        }
        if (a.rows != nullptr) reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.rows) + (size_t)s * a.H)[v] = raw;
        if (loc >= 0) {
          float f[8];
          unpack8<T>(raw, f);
          const float w = a.weights ? a.weights[s] : 1.0f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += w * f[j];
        }
      }
      reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.out) + (size_t)t * a.H)[v] = pack8<T>(acc);
    }
  }
}

// Tile table of the grouped expert GEMMs, derived on the device from the dispatch kernel's segment table (start[e], count[e], total, overflow):
//   tile_group[t]  expert that owns rows [128 t, 128 t + 128) of the expert-major receive buffer, -1 = no expert (tile skipped)
//   seg2[2e..]     (row start, row count padded to the alignment) of expert e, clamped to the buffer: the K ranges of the grouped wgrad
// A receive-buffer overflow is OR-ed into a sticky flag that the host polls asynchronously (one step late), never on the critical path.
__global__ void __launch_bounds__(256) moe_tile_table_kernel(const int* __restrict__ seg, int e_local, int align, int cap_rows, int n_tiles,
                                                             int* __restrict__ tile_group, int* __restrict__ seg2, int* __restrict__ sticky) {
  __shared__ int s_start[kMoeMaxExperts], s_end[kMoeMaxExperts];
  for (int e = threadIdx.x; e < e_local; e += blockDim.x) {
    const int start = min(seg[e], cap_rows);
    int padded = (seg[e_local + e] + align - 1) / align * align;
    if (start + padded > cap_rows) padded = (cap_rows - start) / align * align;      // overflow: whole blocks that still fit
    s_start[e] = start; s_end[e] = start + padded;
    seg2[2 * e] = start; seg2[2 * e + 1] = padded;
  }
  if (threadIdx.x == 0 && seg[2 * e_local + 1] != 0) atomicOr(sticky, 1);
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) {
    const int row = t * 128;
    int g = -1;
    for (int e = 0; e < e_local; ++e) if (row >= s_start[e] && row + 128 <= s_end[e]) g = e;
    tile_group[t] = g;
  }
}

// Per-expert column sums (bias gradients of the grouped expert linears): out[e, c] = sum over the rows of segment e of x[r, c].
// grid (column chunks of 256, experts); 32 column vectors x 8 row lanes per CTA, fp32 accumulation, one store per column.
template <typename T>
__global__ void __launch_bounds__(256) grouped_colsum_kernel(const T* __restrict__ x, const int* __restrict__ seg2, int N, T* __restrict__ out) {
  const int e = blockIdx.y;
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cv) * 8;
  const int r0 = seg2[2 * e], r1 = r0 + seg2[2 * e + 1];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < N) {
    for (int r = r0 + rl; r < r1; r += 8) {
      float v[8];
      unpack8<T>(ld_stream(reinterpret_cast<const uint4*>(x + (size_t)r * N + c)), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  __shared__ float part[8][32][9];
#pragma unroll
  for (int j = 0; j < 8; ++j) part[rl][cv][j] = acc[j];
  __syncthreads();
  if (rl == 0 && c < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += part[q][cv][j];
      acc[j] = t;
    }
    *reinterpret_cast<uint4*>(out + (size_t)e * N + c) = pack8<T>(acc);
  }
}

}  // namespace

cudaError_t moe_route(const int64_t* gate_idx, int num_slots, int total_experts, int* slot_rank, int* counts, cudaStream_t st) {
  if (total_experts > kMoeMaxExperts) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(slot_rank, 0xFF, sizeof(int) * (size_t)num_slots, st);
  if (e != cudaSuccess) return e;
  moe_route_kernel<<<total_experts, 256, 0, st>>>(gate_idx, num_slots, slot_rank, counts);
  return cudaGetLastError();
}

cudaError_t moe_dispatch(const void* src, const float* scale, const int64_t* gate_idx, const int* slot_rank, const int* counts, int* slot_loc,
                         int* seg, void** peer_recv, void** peer_cnt, void** peer_flags, unsigned* block_counter, int num_slots, int src_div,
                         int H, int e_local, int world, int rank, int align, int cap_rows, uint32_t epoch, int dtype, int num_ctas,
                         cudaStream_t st) {
  if (world > kMoeMaxPeers || world * e_local > kMoeMaxExperts || H % 8 || cap_rows >= (1 << 24)) return cudaErrorInvalidValue;
  DispatchArgs a{};
  a.src = src; a.scale = scale; a.gate_idx = gate_idx; a.slot_rank = slot_rank; a.counts = counts; a.slot_loc = slot_loc; a.seg = seg;
  for (int i = 0; i < world; ++i) { a.recv.p[i] = peer_recv[i]; a.cnt.p[i] = peer_cnt[i]; a.flags.p[i] = peer_flags[i]; }
  a.block_counter = block_counter;
  a.num_slots = num_slots; a.src_div = src_div; a.H = H; a.e_local = e_local; a.world = world; a.rank = rank; a.align = align;
  a.cap_rows = cap_rows; a.epoch = epoch;
  const int e_total = world * e_local;
  const size_t smem = sizeof(int) * ((size_t)world * e_total + e_total);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (dtype == 1) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(moe_dispatch_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    moe_dispatch_kernel<__nv_bfloat16><<<num_ctas, 256, smem, st>>>(a);
  } else if (dtype == 0) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(moe_dispatch_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    moe_dispatch_kernel<__half><<<num_ctas, 256, smem, st>>>(a);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t moe_combine(void** peer_src, const int* slot_loc, const float* weights, void* out, void* rows, void** peer_flags, int T, int topk,
                        int H, int world, int rank, uint32_t epoch, int dtype, int num_ctas, cudaStream_t st) {
  if (world > kMoeMaxPeers || H % 8) return cudaErrorInvalidValue;
  CombineArgs a{};
  for (int i = 0; i < world; ++i) { a.src.p[i] = peer_src[i]; a.flags.p[i] = peer_flags[i]; }
  a.slot_loc = slot_loc; a.weights = weights; a.out = out; a.rows = rows; a.T = T; a.topk = topk; a.H = H; a.world = world; a.rank = rank;
  a.epoch = epoch;
  if (dtype == 1) moe_combine_kernel<__nv_bfloat16><<<num_ctas, 256, 0, st>>>(a);
  else if (dtype == 0) moe_combine_kernel<__half><<<num_ctas, 256, 0, st>>>(a);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t moe_tile_table(const int* seg, int e_local, int align, int cap_rows, int* tile_group, int* seg2, int* sticky, cudaStream_t st) {
  if (e_local > kMoeMaxExperts || align % 128 || cap_rows % 128) return cudaErrorInvalidValue;
  moe_tile_table_kernel<<<1, 256, 0, st>>>(seg, e_local, align, cap_rows, cap_rows / 128, tile_group, seg2, sticky);
  return cudaGetLastError();
}

cudaError_t grouped_colsum(const void* x, const int* seg2, int groups, int N, void* out, int dtype, cudaStream_t st) {
  if (N % 8 || groups < 1) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)((N / 8 + 31) / 32), (unsigned)groups);
  if (dtype == 1) grouped_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, seg2, N, (__nv_bfloat16*)out);
  else if (dtype == 0) grouped_colsum_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, seg2, N, (__half*)out);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace pfx

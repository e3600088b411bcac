// Flash-style attention forward on tcgen05 / TMEM / TMA (sm_100a), bf16, head_dim 64 or 128, causal or full.
//
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] . K[b, :, h, :]^T (+ causal mask)) @ V[b, :, h, :]
//
// Tensors stay in the framework layout [B, S, H, D] (no transposes): a 2-D tensor map over [B*S, H*D] addresses the
// (128 rows x D) tile of one head directly.  One CTA owns a 128-query tile of one (batch, head):
//   warp 0   TMA producer: Q once, then K / V tiles of 128 keys through a 2-stage ring
//   warp 1   MMA issuer:   S = Q K^T  (UMMA 128x128x16, both operands K-major)   -> TMEM columns [0, 128)
//                          PV = P V   (UMMA 128xDx16, P K-major from smem, V MN-major) -> TMEM columns [128, 128 + D)
//   warps 2-9 softmax:     two threads per query row (each reads its TMEM lane; one takes score columns 0-63 and output channels
//                          0..D/2, the other the rest): scores drained from TMEM in one pass, row max exchanged through shared
//                          memory, exp2 with the running max, P written as bf16 in the 128-byte-swizzled K-major layout the second
//                          MMA consumes; O stays in TMEM (the PV MMA accumulates across tiles) and is rescaled only when the running
//                          maximum has moved by more than 2^8 (lazy rescale), so no softmax thread waits for a PV MMA per tile
// S for tile j+1 is issued as soon as the softmax warps have drained S_j from TMEM, so the tensor core computes the next
// scores while the softmax of the current tile is in its exp / store phase.  Causal tiles above the diagonal are never loaded.
// Output O (bf16) and the row-wise log-sum-exp (fp32, [B, H, S]) are written straight from registers.
//
// Used on the no-grad paths (prefill of generation, evaluation, vision / text encoders in inference).  Training keeps the
// library (cuDNN) kernels: a matching backward is future work.
// Reference call site: flash_attention in hybrid_model.py:284-301 (FlashAttention-2 library on Ampere mma.sync).
#include <cstdio>

#include "pfx_ptx.cuh"
#include "pfx_gemm.h"
#include "pfx_kernels.h"
#include <cudaTypedefs.h>

namespace pfx {

namespace {

constexpr int kFaThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two column halves x four TMEM lane quarters)
constexpr int kFaTile = 128;          // queries per CTA, keys per KV tile

template <int kD>
struct FaSmem {
  static constexpr int kQBytes = kFaTile * kD * 2;
  static constexpr int kKBytes = kFaTile * kD * 2;
  static constexpr int kVBytes = kFaTile * kD * 2;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kPBytes = kFaTile * kFaTile * 2;
  static constexpr int kStages = 2;
  static constexpr int kBarBytes = 128;
  static constexpr int kXchgBytes = 2 * 2 * kFaTile * 4;      // row-statistic exchange between the two threads of a row
  static constexpr int kAlignSlack = 768;                     // dynamic smem starts 1 KB-aligned when the kernel has no static smem; checked at run time
  static constexpr int kUsed = 2 * kQBytes + kStages * kStageBytes + kPBytes + kBarBytes + kXchgBytes;   // Q is double-buffered across work items
  static constexpr int kTotal = kAlignSlack + kUsed;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// Persistent kernel: one CTA per SM walks a list of (query tile, head, batch) work items, longest first, so that barrier / TMEM
// set-up happens once and the loads of the next item (Q into the other Q buffer, its first K/V tiles) are in flight while the
// current item finishes; measured per-CTA set-up + drain was ~6 us against ~1.8 us per KV tile in the one-CTA-per-tile version.
template <int kD, bool kCausal>
__global__ void __launch_bounds__(kFaThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int B, int Sq, int Sk, int H, float scale_log2) {
  using S = FaSmem<kD>;
  constexpr int kPanels = kD / 64;                 // 64-element (128-byte) column panels of a D-wide tile
  constexpr int kTmemCols = 256;                   // S: [0,128), O: [128, 128 + kD)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_kv = smem_q + 2 * S::kQBytes;
  const uint32_t smem_p = smem_kv + S::kStages * S::kStageBytes;
  const uint32_t smem_bar = smem_p + S::kPBytes;
  const uint32_t smem_xchg = smem_bar + S::kBarBytes;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  if (smem_base - smem_u32(smem_raw) > (uint32_t)S::kAlignSlack) {
    if (threadIdx.x == 0) printf("pfx attention_fwd: dynamic shared memory base is misaligned by %u bytes\n", smem_base - smem_u32(smem_raw));
    __trap();
  }
  auto q_full = [&](int i) { return smem_bar + 8u * i; };
  auto q_empty = [&](int i) { return smem_bar + 8u * (2 + i); };
  auto kv_full = [&](int s) { return smem_bar + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return smem_bar + 8u * (6 + s); };
  const uint32_t s_full = smem_bar + 8u * 8, s_free = smem_bar + 8u * 9, p_full = smem_bar + 8u * 10, pv_full = smem_bar + 8u * 11;
  const uint32_t tmem_slot = smem_bar + 8u * 12;

  const uint32_t warp = warp_id(), lane = lane_id();
  const int n_q_tiles = (Sq + kFaTile - 1) / kFaTile;
  const int n_kv_all = (Sk + kFaTile - 1) / kFaTile;
  const int hb_count = H * B;
  const int n_items = n_q_tiles * hb_count;
  // item w -> (q_tile, h, b): all (h, b) of the longest query tile first, then the next shorter one, ...
  auto item_q_tile = [&](int w) { return n_q_tiles - 1 - w / hb_count; };
  auto item_n_kv = [&](int q_tile) {
    return kCausal ? min(n_kv_all, (q_tile * kFaTile + kFaTile - 1 + (Sk - Sq)) / kFaTile + 1) : n_kv_all;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < 2; ++i) { mbar_init(q_full(i), 1); mbar_init(q_empty(i), 1); }
      for (int s2 = 0; s2 < S::kStages; ++s2) { mbar_init(kv_full(s2), 1); mbar_init(kv_empty(s2), 1); }
      mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(pv_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, kTmemCols);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 128;

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      uint32_t t = 0;                                   // KV tiles issued so far (all items): ring stage / phase
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int q_tile = item_q_tile(w), hb = w % hb_count, h = hb % H, b = hb / H;
        const int n_kv = item_n_kv(q_tile);
        const int qb = it & 1;
        mbar_wait(q_empty(qb), (((uint32_t)it >> 1) & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(qb), S::kQBytes);
        for (int p = 0; p < kPanels; ++p)
          tma_load_2d(&tmap_q, q_full(qb), smem_q + qb * S::kQBytes + p * (kFaTile * 128), h * kD + p * 64, b * Sq + q_tile * kFaTile);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int stage = t & 1;
          mbar_wait(kv_empty(stage), ((t >> 1) & 1u) ^ 1u);
          const uint32_t sk = smem_kv + stage * S::kStageBytes, sv = sk + S::kKBytes;
          mbar_arrive_expect_tx(kv_full(stage), S::kStageBytes);
          const int key0 = b * Sk + j * kFaTile;
          for (int p = 0; p < kPanels; ++p) tma_load_2d(&tmap_k, kv_full(stage), sk + p * (kFaTile * 128), h * kD + p * 64, key0);
          for (int kb = 0; kb < 2; ++kb)            // V as the MN-major B operand: [64 keys x 64 channels] boxes, channel chunks 8 KB apart
            for (int nc = 0; nc < kPanels; ++nc)
              tma_load_2d(&tmap_v, kv_full(stage), sv + (kb * kPanels + nc) * 8192, h * kD + nc * 64, key0 + kb * 64);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(1, 1, 1, false, false, kFaTile, kFaTile);
      const uint32_t idesc_pv = umma_idesc(1, 1, 1, false, true, kFaTile, kD);
      constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);        // K-major, 128-byte swizzle
      constexpr uint64_t kDescMN = umma_desc_hi_lo(8192, 1024);     // MN-major: 64-channel chunks 8 KB apart
      uint32_t t = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int n_kv = item_n_kv(item_q_tile(w));
        const int qb = it & 1;
        const uint32_t sq = smem_q + qb * S::kQBytes;
        mbar_wait(q_full(qb), ((uint32_t)it >> 1) & 1u);
        auto issue_s = [&](uint32_t tt) {              // S = Q K^T for global tile tt
          const int stage = tt & 1;
          mbar_wait(kv_full(stage), (tt >> 1) & 1u);
          mbar_wait(s_free, (tt & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t sk = smem_kv + stage * S::kStageBytes;
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t off = (k / 4) * (kFaTile * 128) + (k % 4) * 32;
            umma_f16<1>(tmem_s, umma_desc(sq + off, kDescK), umma_desc(sk + off, kDescK), idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit<1>(s_full);
        };
        issue_s(t);
        for (int j = 0; j < n_kv; ++j, ++t) {
          if (j + 1 < n_kv) issue_s(t + 1);          // next scores while the softmax of this tile runs
          else umma_commit<1>(q_empty(qb));          // every S MMA of this item has been issued: its Q buffer frees when they finish
          const int stage = t & 1;
          mbar_wait(p_full, t & 1u);                 // P is in shared memory AND O has been rescaled to this tile's reference maximum
          tcgen05_fence_after();
          const uint32_t sv = smem_kv + stage * S::kStageBytes + S::kKBytes;
#pragma unroll
          for (int k = 0; k < kFaTile / 16; ++k) {
            const uint32_t a_off = (k / 4) * (kFaTile * 128) + (k % 4) * 32;                 // P: two 64-key panels
            const uint32_t b_off = (k / 4) * (kPanels * 8192) + (k % 4) * 2048;              // V: 64-key blocks, 16 keys = 2 KB
            umma_f16<1>(tmem_pv, umma_desc(smem_p + a_off, kDescK), umma_desc(sv + b_off, kDescMN), idesc_pv, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit<1>(pv_full);
          umma_commit<1>(kv_empty(stage));
        }
      }
    }
  } else {
    // ======================================================================================= softmax / epilogue
    // Two threads per query row: warps 2-5 own score columns [0, 64) and output channels [0, D/2), warps 6-9 the other halves (both
    // groups map onto the same four TMEM lane quarters).  The row maximum is exchanged through shared memory once per KV tile.
    float (*s_xchg)[2][kFaTile] = reinterpret_cast<float (*)[2][kFaTile]>(smem_gen + (smem_xchg - smem_base));
    const uint32_t q = warp & 3u;
    const int half = (int)((warp - 2u) >> 2);
    const int row = (int)(q * 32u + lane);               // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (q * 32u) << 16;
    constexpr int kDH = kD / 2;
    // O is accumulated by the tensor core in TMEM across all KV tiles (FlashAttention-4 style).  Probabilities are taken relative
    // to a reference maximum m_ref that is only advanced — and O / l rescaled — when the true running maximum has moved by more
    // than 2^8, so the rescale (a TMEM load + store of this thread's D/2 channels) is rare after the first tiles and nothing in
    // the per-tile critical path waits for the PV MMA.
    constexpr float kRescaleThreshold = 8.f;
    uint8_t* p_row = smem_gen + (smem_p - smem_base) + half * (kFaTile * 128) + (row / 8) * 1024 + (row % 8) * 128;
    uint32_t t = 0, x = 0;                              // global tile counter, exchange-buffer counter
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int q_tile = item_q_tile(w), hb = w % hb_count, h = hb % H, b = hb / H;
      const int n_kv = item_n_kv(q_tile);
      const int row_g = q_tile * kFaTile + row;
      float m_ref = -INFINITY, l = 0.f;
      for (int j = 0; j < n_kv; ++j, ++t, ++x) {
        const int col0 = j * kFaTile + half * 64;
        mbar_wait(s_full, t & 1u);
        tcgen05_fence_after();
        const bool need_mask = (kCausal && col0 + 63 > row_g + (Sk - Sq)) || (col0 + 64 > Sk);
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_s + lane_addr + half * 64, r0);
        tmem_ld_32x32b_x32(tmem_s + lane_addr + half * 64 + 32, r1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);            // S drained into registers: the next QK^T may overwrite it
        float sc[64];
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float v = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]) * scale_log2;
          if (need_mask) {
            const int cg = col0 + i;
            if (cg >= Sk || (kCausal && cg > row_g + (Sk - Sq))) v = -INFINITY;
          }
          sc[i] = v;
          mx[i & 3] = fmaxf(mx[i & 3], v);
        }
        const float m_part = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        s_xchg[x & 1][half][row] = m_part;
        // only the two warps that share these 32 rows have to meet (named barrier 2 + lane quarter, 64 threads)
        asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
        const float m_new = fmaxf(m_part, s_xchg[x & 1][half ^ 1][row]);      // this tile's row maximum (identical in both halves)
        const bool want = (m_new > m_ref + kRescaleThreshold) || (m_ref == -INFINITY && m_new != -INFINITY);
        const bool rescale = __any_sync(0xffffffffu, want);
        if (j > 0) {                                   // PV of the previous tile must be complete before O is touched or P is overwritten
          mbar_wait(pv_full, (t - 1) & 1u);
          tcgen05_fence_after();
        }
        if (rescale) {
          const float m_next = fmaxf(m_ref, m_new);
          const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_next);
          if (j > 0) {
#pragma unroll
            for (int c = 0; c < kDH; c += 32) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(tmem_pv + lane_addr + half * kDH + c, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st_32x32b_x32(tmem_pv + lane_addr + half * kDH + c, r);
            }
            tmem_st_wait();
          }
          l *= f;
          m_ref = m_next;
        }
        const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int g = 0; g < 8; ++g) {                  // eight 16-byte chunks (8 keys each) of this half's 64-key panel
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { pv[i] = exp2f(sc[g * 8 + i] - m_use); ls[i & 3] += pv[i]; }
          const int chunk = g ^ (row % 8);
          uint4 v;
          v.x = pack_bf16x2(pv[0], pv[1]); v.y = pack_bf16x2(pv[2], pv[3]); v.z = pack_bf16x2(pv[4], pv[5]); v.w = pack_bf16x2(pv[6], pv[7]);
          *reinterpret_cast<uint4*>(p_row + chunk * 16) = v;
        }
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        tcgen05_fence_before();
        fence_proxy_async_smem();                    // generic-proxy stores of P -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      // item epilogue: wait for the last PV, read this thread's half of O once, normalise by the full row sum.  The next item's
      // loads and first QK^T are already running underneath.
      mbar_wait(pv_full, (t - 1) & 1u);
      tcgen05_fence_after();
      s_xchg[x & 1][half][row] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
      const float l_all = l + s_xchg[x & 1][half ^ 1][row];
      ++x;
      const float inv = l_all > 0.f ? 1.f / l_all : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)(b * Sq + row_g) * H + h) * kD + half * kDH);
#pragma unroll
      for (int c = 0; c < kDH; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_pv + lane_addr + half * kDH + c, r);
        tmem_ld_wait();
        if (row_g < Sq) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[i + 0]) * inv, __uint_as_float(r[i + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
            dst[(c + i) / 8] = v;
          }
        }
      }
      if (row_g < Sq && lse != nullptr && half == 0)
        lse[((size_t)b * H + h) * Sq + row_g] = (l_all > 0.f) ? (m_ref + log2f(l_all)) * 0.6931471805599453f : -INFINITY;
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, kTmemCols);
}

template <int kD, bool kCausal>
cudaError_t launch_fa(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Sk, int H, float scale, cudaStream_t st) {
  using S = FaSmem<kD>;
  CUtensorMap tq, tk, tv;
  const uint64_t row_bytes = (uint64_t)H * kD * 2;
  bool ok = make_tmap_2d(&tq, q, 2, 1, (uint64_t)H * kD, (uint64_t)B * Sq, row_bytes, 64, kFaTile);
  ok &= make_tmap_2d(&tk, k, 2, 1, (uint64_t)H * kD, (uint64_t)B * Sk, row_bytes, 64, kFaTile);
  ok &= make_tmap_2d(&tv, v, 2, 1, (uint64_t)H * kD, (uint64_t)B * Sk, row_bytes, 64, 64);
  if (!ok) return cudaErrorInvalidValue;
  auto kern = attention_fwd_kernel<kD, kCausal>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long items = (long)((Sq + kFaTile - 1) / kFaTile) * H * B;
  const int grid = (int)(items < sms ? items : sms);
  kern<<<grid, kFaThreads, S::kTotal, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, B, Sq, Sk, H, scale * 1.4426950408889634f);
  return cudaGetLastError();
}

}  // namespace

cudaError_t attention_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Sk, int H, int D, float scale,
                          bool causal, cudaStream_t st) {
  if ((D != 64 && D != 128) || B < 1 || Sq < 1 || Sk < 1 || (causal && Sk < Sq)) return cudaErrorInvalidValue;
  if (D == 128) return causal ? launch_fa<128, true>(q, k, v, out, lse, B, Sq, Sk, H, scale, st) : launch_fa<128, false>(q, k, v, out, lse, B, Sq, Sk, H, scale, st);
  return causal ? launch_fa<64, true>(q, k, v, out, lse, B, Sq, Sk, H, scale, st) : launch_fa<64, false>(q, k, v, out, lse, B, Sq, Sk, H, scale, st);
}

}  // namespace pfx

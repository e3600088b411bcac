// Flash-style attention forward on tcgen05 / TMEM / TMA (sm_100a), bf16, head_dim 64 or 128, causal or full, optional dropout.
//
//   O[b, s, h, :] = softmax(scale * Q[b, s, h, :] . K[b, :, h, :]^T (+ causal mask)) (dropout) @ V[b, :, h, :]
//
// Inputs / output are [B, S, H, D] VIEWS (3-D tensor maps over {columns, S, B}): contiguous tensors, the q / k / v slices of a packed
// projection output and sequence-major storage are all read in place.
//
// One CTA owns TWO adjacent 128-query tiles of one (batch, head) and streams the K / V tiles they share.  The two tiles ping-pong:
// while the eight softmax warps work on the scores of one tile, the tensor core runs P V of the other tile and its next Q K^T, so
// neither side waits for the other in steady state (the first version, one tile per CTA, alternated the two and sat at 15 % tensor
// activity / 335 TFLOP/s at S = 1024):
//   warp 0    TMA producer: both Q tiles once per work item, then K / V tiles of 128 keys through a 2-stage ring
//   warp 1    MMA issuer:   S_x = Q_x K^T (UMMA 128x128x16, operands from shared memory) -> TMEM; P_x V (UMMA 128xDx16 with the A operand
//                           read from TENSOR MEMORY: P_x is written by the softmax threads as packed bf16 over the first 64 columns of S_x
//                           — layout verified by probe_kernels.cu — so probabilities never touch shared memory); O_x accumulates in TMEM
//   warps 2-9 softmax:      two threads per query row (TMEM lane): scores -> running max (exchanged through shared memory) -> exp2 with the
//                           scale folded into one FFMA -> dropout (counter hash, pfx_attn.cuh) -> P to TMEM; O is rescaled lazily (only
//                           when the running maximum moved by more than 2^8), so no softmax thread waits for a P V product per tile
// TMEM: S_0 / P_0 [0, 128), S_1 / P_1 [128, 256), O_0 [256, 256 + D), O_1 [384, 384 + D).  Causal tiles above the diagonal are never loaded;
// masks are applied only on diagonal / ragged tiles.  Output O (bf16) and the row-wise log-sum-exp (fp32, [B, H, S]) are written from
// registers.  Reference call site: flash_attention in hybrid_model.py:284-301 (FlashAttention-2 library on Ampere mma.sync).
#include <cstdio>

#include "pfx_ptx.cuh"
#include "pfx_gemm.h"
#include "pfx_kernels.h"
#include "pfx_attn.h"
#include "pfx_attn.cuh"
#include <cudaTypedefs.h>

namespace pfx {

namespace {

constexpr int kFaThreads = 320;        // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two column halves x four TMEM lane quarters)
constexpr int kFaTile = 128;          // queries per tile, keys per KV tile

template <int kD>
struct FaSmem {
  static constexpr int kQBytes = kFaTile * kD * 2;            // one query tile
  static constexpr int kKBytes = kFaTile * kD * 2;
  static constexpr int kVBytes = kFaTile * kD * 2;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kStages = 2;
  static constexpr int kBarBytes = 128;
  static constexpr int kXchgBytes = 2 * 2 * kFaTile * 4;      // row-statistic exchange between the two threads of a row (double buffered)
  static constexpr int kAlignSlack = 1024;
  static constexpr int kUsed = 2 * kQBytes + kStages * kStageBytes + kBarBytes + kXchgBytes;
  static constexpr int kTotal = kAlignSlack + kUsed;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tma3(const CUtensorMap* m, bool swapped, uint32_t bar, uint32_t dst, int col, int s, int b) {
  if (swapped) tma_load_3d(m, bar, dst, col, b, s); else tma_load_3d(m, bar, dst, col, s, b);
}


template <int kD, bool kCausal>
__global__ void __launch_bounds__(kFaThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     __nv_bfloat16* __restrict__ out, int64_t o_sb, int64_t o_ss, int64_t o_sh, float* __restrict__ lse, int B, int Sq, int Sk, int H,
                     float scale_log2, int hs_q, int hs_k, int hs_v, uint32_t swapped, uint32_t drop_thresh16, float inv_keep, uint64_t seed) {
  using S = FaSmem<kD>;
  constexpr int kPanels = kD / 64;                 // 64-element (128-byte) column panels of a D-wide tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;                                  // two query tiles
  const uint32_t smem_kv = smem_q + 2 * S::kQBytes;
  const uint32_t smem_bar = smem_kv + S::kStages * S::kStageBytes;
  const uint32_t smem_xchg = smem_bar + S::kBarBytes;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_full = smem_bar, q_empty = smem_bar + 8;
  auto kv_full = [&](int s) { return smem_bar + 8u * (2 + s); };
  auto kv_empty = [&](int s) { return smem_bar + 8u * (4 + s); };
  auto s_full = [&](int x) { return smem_bar + 8u * (6 + x); };
  auto p_full = [&](int x) { return smem_bar + 8u * (8 + x); };
  auto pv_full = [&](int x) { return smem_bar + 8u * (10 + x); };
  const uint32_t tmem_slot = smem_bar + 8u * 12;

  const uint32_t warp = warp_id(), lane = lane_id();
  const int n_q_tiles = (Sq + kFaTile - 1) / kFaTile;
  const int n_pairs = (n_q_tiles + 1) / 2;
  const int n_kv_all = (Sk + kFaTile - 1) / kFaTile;
  const int hb_count = H * B;
  const int n_items = n_pairs * hb_count;
  // item w -> (query-tile pair, h, b): all (h, b) of the longest pair first, then the next shorter one, ...
  auto item_pair = [&](int w) { return n_pairs - 1 - w / hb_count; };
  auto tile_n_kv = [&](int q_tile) {                 // key tiles a query tile attends to (0 for the phantom second tile of an odd count)
    if (q_tile >= n_q_tiles) return 0;
    return kCausal ? min(n_kv_all, (q_tile * kFaTile + kFaTile - 1 + (Sk - Sq)) / kFaTile + 1) : n_kv_all;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1); mbar_init(q_empty, 1);
      for (int s2 = 0; s2 < S::kStages; ++s2) { mbar_init(kv_full(s2), 1); mbar_init(kv_empty(s2), 1); }
      for (int x = 0; x < 2; ++x) { mbar_init(s_full(x), 1); mbar_init(p_full(x), 8); mbar_init(pv_full(x), 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  auto tmem_s = [&](int x) { return tmem_base + 128u * x; };
  auto tmem_o = [&](int x) { return tmem_base + 256u + 128u * x; };

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      uint32_t t = 0;                                   // KV tiles issued so far (all items): ring stage / phase
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int pair = item_pair(w), hb = w % hb_count, h = hb % H, b = hb / H;
        const int n_kv = max(tile_n_kv(2 * pair), tile_n_kv(2 * pair + 1));
        mbar_wait(q_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, 2 * S::kQBytes);
        for (int x = 0; x < 2; ++x)
          for (int p = 0; p < kPanels; ++p)
            tma3(&tmap_q, swapped & 1u, q_full, smem_q + x * S::kQBytes + p * (kFaTile * 128), h * hs_q + p * 64, (2 * pair + x) * kFaTile, b);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int stage = t & 1;
          mbar_wait(kv_empty(stage), ((t >> 1) & 1u) ^ 1u);
          const uint32_t sk = smem_kv + stage * S::kStageBytes, sv = sk + S::kKBytes;
          mbar_arrive_expect_tx(kv_full(stage), S::kStageBytes);
          const int key0 = j * kFaTile;
          for (int p = 0; p < kPanels; ++p) tma3(&tmap_k, swapped & 2u, kv_full(stage), sk + p * (kFaTile * 128), h * hs_k + p * 64, key0, b);
          for (int kb = 0; kb < 2; ++kb)            // V as the MN-major B operand: [64 keys x 64 channels] boxes, channel chunks 8 KB apart
            for (int nc = 0; nc < kPanels; ++nc)
              tma3(&tmap_v, swapped & 4u, kv_full(stage), sv + (kb * kPanels + nc) * 8192, h * hs_v + nc * 64, key0 + kb * 64, b);
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
      uint32_t t = 0;                                   // global KV tile counter (ring position)
      uint32_t cnt[2] = {0, 0};                         // tiles processed per slot x: phases of s_full / p_full / pv_full
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int pair = item_pair(w);
        const int nk[2] = {tile_n_kv(2 * pair), tile_n_kv(2 * pair + 1)};
        const int n_kv = max(nk[0], nk[1]);
        mbar_wait(q_full, (uint32_t)it & 1u);
        auto issue_s = [&](int x, uint32_t tt) {          // S_x = Q_x K^T for ring position tt
          const int stage = tt & 1;
          const uint32_t sq = smem_q + x * S::kQBytes, sk = smem_kv + stage * S::kStageBytes;
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t off = (k / 4) * (kFaTile * 128) + (k % 4) * 32;
            umma_f16<1>(tmem_s(x), umma_desc(sq + off, kDescK), umma_desc(sk + off, kDescK), idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit<1>(s_full(x));
        };
        mbar_wait(kv_full(t & 1), (t >> 1) & 1u);
        tcgen05_fence_after();
        for (int x = 0; x < 2; ++x) if (nk[x] > 0) issue_s(x, t);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int stage = t & 1;
          const uint32_t sv = smem_kv + stage * S::kStageBytes + S::kKBytes;
          for (int x = 0; x < 2; ++x) {
            if (j >= nk[x]) continue;
            mbar_wait(p_full(x), cnt[x] & 1u);          // P_x is in tensor memory AND O_x has been rescaled to this tile's reference maximum
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < kFaTile / 16; ++k) {
              const uint32_t b_off = (k / 4) * (kPanels * 8192) + (k % 4) * 2048;              // V: 64-key blocks, 16 keys = 2 KB
              umma_f16_ts(tmem_o(x), tmem_s(x) + k * 8, umma_desc(sv + b_off, kDescMN), idesc_pv, (j | k) != 0 ? 1u : 0u);
            }
            umma_commit<1>(pv_full(x));
            ++cnt[x];
            if (j + 1 < nk[x]) {                        // next scores of this tile right behind its P V: the other tile's softmax is running now
              mbar_wait(kv_full((t + 1) & 1), ((t + 1) >> 1) & 1u);
              tcgen05_fence_after();
              issue_s(x, t + 1);
            }
          }
          umma_commit<1>(kv_empty(stage));              // every product that reads this stage has been issued
        }
        umma_commit<1>(q_empty);                        // ... and every product that reads the two Q tiles
      }
    }
  } else {
    // ======================================================================================= softmax / epilogue
    float (*s_xchg)[2][kFaTile] = reinterpret_cast<float (*)[2][kFaTile]>(smem_gen + (smem_xchg - smem_base));
    const uint32_t q = warp & 3u;
    const int half = (int)((warp - 2u) >> 2);
    const int row = (int)(q * 32u + lane);               // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (q * 32u) << 16;
    constexpr int kDH = kD / 2;
    constexpr float kRescaleThreshold = 8.f;
    uint32_t xc = 0;                                     // exchange-buffer counter
    uint32_t cnt[2] = {0, 0};
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int pair = item_pair(w), hb = w % hb_count, h = hb % H, b = hb / H;
      const int nk[2] = {tile_n_kv(2 * pair), tile_n_kv(2 * pair + 1)};
      const int n_kv = max(nk[0], nk[1]);
      float m_ref[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
      const uint32_t rng_key = attn_rng_key(seed, (uint32_t)(b * H + h));
      const uint32_t rng_pairs = ((uint32_t)Sk + 1u) >> 1;
      for (int j = 0; j < n_kv; ++j) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (j >= nk[x]) continue;
          const int q_tile = 2 * pair + x;
          const int row_g = q_tile * kFaTile + row;
          const int col0 = j * kFaTile + half * 64;
          mbar_wait(s_full(x), cnt[x] & 1u);
          tcgen05_fence_after();
          const bool need_mask = (kCausal && col0 + 63 > q_tile * kFaTile + (Sk - Sq)) || (col0 + 64 > Sk);
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(tmem_s(x) + lane_addr + half * 64, r0);
          tmem_ld_32x32b_x32(tmem_s(x) + lane_addr + half * 64 + 32, r1);
          tmem_ld_wait();
          float sc[64];
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          if (need_mask) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              float v = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]);
              const int cg = col0 + i;
              if (cg >= Sk || (kCausal && cg > row_g + (Sk - Sq))) v = -INFINITY;
              sc[i] = v;
              mx[i & 3] = fmaxf(mx[i & 3], v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              const float v = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]);
              sc[i] = v;
              mx[i & 3] = fmaxf(mx[i & 3], v);
            }
          }
          const float m_part = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale_log2;     // log2 units from here on
          s_xchg[xc & 1][half][row] = m_part;
          // only the two warps that share these 32 rows have to meet (named barrier 2 + lane quarter, 64 threads); it also orders the two
          // halves' TMEM reads of S_x before either half overwrites its part of the aliased P_x columns
          asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
          const float m_new = fmaxf(m_part, s_xchg[xc & 1][half ^ 1][row]);      // this tile's row maximum (identical in both halves)
          ++xc;
          const bool want = (m_new > m_ref[x] + kRescaleThreshold) || (m_ref[x] == -INFINITY && m_new != -INFINITY);
          const bool rescale = __any_sync(0xffffffffu, want);
          if (rescale) {
            const float m_next = fmaxf(m_ref[x], m_new);
            const float f = (m_ref[x] == -INFINITY) ? 0.f : exp2f(m_ref[x] - m_next);
            if (j > 0) {
              mbar_wait(pv_full(x), (cnt[x] - 1u) & 1u);           // the previous P V of this tile has landed in O_x
              tcgen05_fence_after();
#pragma unroll
              for (int c = 0; c < kDH; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_o(x) + lane_addr + half * kDH + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                tmem_st_32x32b_x32(tmem_o(x) + lane_addr + half * kDH + c, r);
              }
            }
            l[x] *= f;
            m_ref[x] = m_next;
          }
          const float m_use = (m_ref[x] == -INFINITY) ? 0.f : m_ref[x];
          float ls[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t pw[32];                                // this half's 64 probabilities as packed bf16
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { pv[i] = exp2f(fmaf(sc[g * 8 + i], scale_log2, -m_use)); ls[i & 3] += pv[i]; }
            if (drop_thresh16 != 0u) {               // the row sum keeps every probability; only the P V product sees the mask
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const uint32_t kg = (uint32_t)(col0 + g * 8 + i);
                const uint32_t bits = attn_rng_pair(rng_key, (uint32_t)row_g, kg >> 1, rng_pairs);
                if ((bits & 0xFFFFu) < drop_thresh16) pv[i] = 0.f;
                if ((bits >> 16) < drop_thresh16) pv[i + 1] = 0.f;
              }
            }
            pw[g * 4 + 0] = pack_bf16x2(pv[0], pv[1]); pw[g * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
            pw[g * 4 + 2] = pack_bf16x2(pv[4], pv[5]); pw[g * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
          }
          l[x] += (ls[0] + ls[1]) + (ls[2] + ls[3]);
          tmem_st_32x32b_x32(tmem_s(x) + lane_addr + half * 32, pw);       // P_x over the first 64 columns of S_x: keys [64 half, 64 half + 64)
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(x));
          ++cnt[x];
        }
      }
      // item epilogue: wait for the last P V of each tile, read this thread's half of O once, normalise by the full row sum
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (nk[x] == 0) continue;
        const int row_g = (2 * pair + x) * kFaTile + row;
        mbar_wait(pv_full(x), (cnt[x] - 1u) & 1u);
        tcgen05_fence_after();
        s_xchg[xc & 1][half][row] = l[x];
        asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
        const float l_all = l[x] + s_xchg[xc & 1][half ^ 1][row];
        ++xc;
        const float inv = l_all > 0.f ? inv_keep / l_all : 0.f;
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)b * o_sb + (size_t)row_g * o_ss + (size_t)h * o_sh + half * kDH);
#pragma unroll
        for (int c = 0; c < kDH; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_o(x) + lane_addr + half * kDH + c, r);
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
          lse[((size_t)b * H + h) * Sq + row_g] = (l_all > 0.f) ? (m_ref[x] + log2f(l_all)) * 0.6931471805599453f : -INFINITY;
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

template <int kD, bool kCausal>
cudaError_t launch_fa(const AttnView& q, const AttnView& k, const AttnView& v, const AttnView& out, float* lse, int B, int Sq, int Sk, int H, float scale,
                      AttnDropout drop, cudaStream_t st) {
  using S = FaSmem<kD>;
  CUtensorMap tq, tk, tv;
  bool sq = false, sk = false, sv = false;
  auto inner = [&](const AttnView& t) { return (uint64_t)(H - 1) * t.sh + kD; };
  bool ok = make_tmap_bshd(&tq, q.ptr, 1, inner(q), Sq, B, (uint64_t)q.ss * 2, (uint64_t)q.sb * 2, 64, kFaTile, &sq);
  ok &= make_tmap_bshd(&tk, k.ptr, 1, inner(k), Sk, B, (uint64_t)k.ss * 2, (uint64_t)k.sb * 2, 64, kFaTile, &sk);
  ok &= make_tmap_bshd(&tv, v.ptr, 1, inner(v), Sk, B, (uint64_t)v.ss * 2, (uint64_t)v.sb * 2, 64, 64, &sv);
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
  const long items = (long)(((Sq + kFaTile - 1) / kFaTile + 1) / 2) * H * B;        // one work item = a pair of query tiles
  const int grid = (int)(items < sms ? items : sms);
  const uint32_t thresh = drop.p > 0.f ? (uint32_t)(drop.p * 65536.f + 0.5f) : 0u;
  const float inv_keep = drop.p > 0.f ? 1.f / (1.f - drop.p) : 1.f;
  const uint32_t swapped = (sq ? 1u : 0u) | (sk ? 2u : 0u) | (sv ? 4u : 0u);
  kern<<<grid, kFaThreads, S::kTotal, st>>>(tq, tk, tv, (__nv_bfloat16*)const_cast<void*>(out.ptr), out.sb, out.ss, out.sh, lse, B, Sq, Sk, H,
                                            scale * 1.4426950408889634f, (int)q.sh, (int)k.sh, (int)v.sh, swapped, thresh, inv_keep, drop.seed);
  return cudaGetLastError();
}

}  // namespace

cudaError_t attention_fwd_v2(const AttnView& q, const AttnView& k, const AttnView& v, const AttnView& out, float* lse, int B, int Sq, int Sk, int H,
                             int D, float scale, bool causal, AttnDropout drop, cudaStream_t st) {
  if ((D != 64 && D != 128) || B < 1 || Sq < 1 || Sk < 1 || (causal && Sk < Sq)) return cudaErrorInvalidValue;
  for (const AttnView* t : {&q, &k, &v, &out})
    if (t->ptr == nullptr || (reinterpret_cast<uintptr_t>(t->ptr) % 16) || t->sb % 8 || t->ss % 8 || t->sh % 8) return cudaErrorInvalidValue;
  if (D == 128) return causal ? launch_fa<128, true>(q, k, v, out, lse, B, Sq, Sk, H, scale, drop, st) : launch_fa<128, false>(q, k, v, out, lse, B, Sq, Sk, H, scale, drop, st);
  return causal ? launch_fa<64, true>(q, k, v, out, lse, B, Sq, Sk, H, scale, drop, st) : launch_fa<64, false>(q, k, v, out, lse, B, Sq, Sk, H, scale, drop, st);
}

cudaError_t attention_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Sq, int Sk, int H, int D, float scale,
                          bool causal, cudaStream_t st) {
  const int64_t hd = (int64_t)H * D;
  const AttnView vq{q, (int64_t)Sq * hd, hd, D}, vk{k, (int64_t)Sk * hd, hd, D}, vv{v, (int64_t)Sk * hd, hd, D}, vo{out, (int64_t)Sq * hd, hd, D};
  return attention_fwd_v2(vq, vk, vv, vo, lse, B, Sq, Sk, H, D, scale, causal, AttnDropout{0.f, 0}, st);
}

}  // namespace pfx

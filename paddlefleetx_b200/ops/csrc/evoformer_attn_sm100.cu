// Evoformer gated attention forward on tcgen05 / TMEM / TMA (sm_100a): the attention primitive of AlphaFold2 / HelixFold
// (reference ppfleetx/models/protein_folding/attentions.py:35-180, its ``fused_gate_attention`` path :126-142).
//
//   logits[g, h, q, k] = scale * Q[g, q, h, :] . K[g, k, h, :] + mask_bias[g, k] + pair_bias[g / groups_per_pair, h, q, k]
//   O[g, q, h, :]      = softmax_k(logits) @ V[g, k, h, :]  *  sigmoid(gate[g, q, h, :])
//
// g runs over batch x {MSA sequences | residues} (thousands of independent small attentions), the head width is 32, H is 4 or 8.  The
// reference materialises the [g, h, q, k] logits in HBM three times (bias add, softmax, gating are separate ops); here they live in
// tensor memory only.
//
// A 32-wide head is half of a 128-byte swizzle row, so one CTA owns a PAIR of heads of one 128-query tile: the Q / K / V tiles are
// loaded as [128 x 64] (both heads, one TMA box each), S_x = Q_x K_x^T runs over the two UMMA K-steps that belong to head x, and P_x V is
// issued at N = 64 over both heads' channels — only the 32 columns of head x are read back.  That wastes half of a product the tensor core
// is not short of and keeps every operand in the layouts the flash-attention kernel already runs on (K-major / MN-major, 128-byte
// swizzle, P from tensor memory).  The two heads ping-pong exactly like the two query tiles of attention_fwd_sm100.cu:
//   warp 0    TMA producer: Q tile once per work item, K / V tiles of 128 keys through a 2-stage ring
//   warp 1    MMA issuer
//   warps 2-9 softmax: two threads per query row; biases are added in registers (mask bias: one value per key, pair bias: this row's 64
//             contiguous values), running max / sum, P to tensor memory, lazy O rescale; the epilogue multiplies by sigmoid(gate)
// Output O (bf16, [g, q, h, 32]) and the row-wise log-sum-exp ([g, h, q] fp32, for the backward).
#include <cstdio>

#include "pfx_ptx.cuh"
#include "pfx_common.cuh"
#include "pfx_gemm.h"
#include "pfx_kernels.h"
#include "pfx_attn.h"
#include <cudaTypedefs.h>

namespace pfx {

namespace {

constexpr int kEvThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 softmax (two column halves x four TMEM lane quarters)
constexpr int kEvTile = 128;          // queries per tile, keys per KV tile
constexpr int kEvD = 64;              // two heads of 32 channels = one 128-byte row
constexpr int kEvHead = 32;

struct EvSmem {
  static constexpr int kQBytes = kEvTile * kEvD * 2;
  static constexpr int kKBytes = kEvTile * kEvD * 2;
  static constexpr int kVBytes = kEvTile * kEvD * 2;
  static constexpr int kStageBytes = kKBytes + kVBytes;
  static constexpr int kStages = 2;
  static constexpr int kBarBytes = 128;
  static constexpr int kXchgBytes = 2 * 2 * kEvTile * 4;
  static constexpr int kAlignSlack = 1024;
  static constexpr int kUsed = kQBytes + kStages * kStageBytes + kBarBytes + kXchgBytes;
  static constexpr int kTotal = kAlignSlack + kUsed;
};

struct EvParams {
  __nv_bfloat16* out;          // [G, Sq, H, 32]
  float* lse;                  // [G, H, Sq] natural log, or nullptr
  const float* mask_bias;      // [G, Sk] fp32 or nullptr
  const __nv_bfloat16* pair_bias;   // [G / groups_per_pair, H, Sq, Sk] bf16 or nullptr
  const __nv_bfloat16* gate;   // [G, Sq, H, 32] pre-sigmoid logits or nullptr
  int G, Sq, Sk, H, groups_per_pair;
  float scale_log2;
};

__global__ void __launch_bounds__(kEvThreads, 1)
evoformer_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                          const __grid_constant__ EvParams prm) {
  using S = EvSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_kv = smem_q + S::kQBytes;
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
  const int Sq = prm.Sq, Sk = prm.Sk, H = prm.H;
  const int n_q_tiles = (Sq + kEvTile - 1) / kEvTile;
  const int n_kv = (Sk + kEvTile - 1) / kEvTile;
  const int n_hp = H / 2;                              // head pairs
  const int n_items = prm.G * n_hp * n_q_tiles;        // item -> (g, head pair, query tile): the tiles of one group are neighbours (K / V stay in L2)
  auto item_qt = [&](int w) { return w % n_q_tiles; };
  auto item_hp = [&](int w) { return (w / n_q_tiles) % n_hp; };
  auto item_g = [&](int w) { return w / (n_q_tiles * n_hp); };

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
  auto tmem_s = [&](int x) { return tmem_base + 128u * x; };            // S_x / P_x of head x
  auto tmem_o = [&](int x) { return tmem_base + 256u + 128u * x; };     // O_x: 64 columns, [32 x, 32 x + 32) are head x's channels

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      uint32_t t = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int g = item_g(w), col = item_hp(w) * kEvD, q0 = item_qt(w) * kEvTile;
        mbar_wait(q_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full, S::kQBytes);
        tma_load_3d(&tmap_q, q_full, smem_q, col, q0, g);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int stage = t & 1;
          mbar_wait(kv_empty(stage), ((t >> 1) & 1u) ^ 1u);
          const uint32_t sk = smem_kv + stage * S::kStageBytes, sv = sk + S::kKBytes;
          mbar_arrive_expect_tx(kv_full(stage), S::kStageBytes);
          const int key0 = j * kEvTile;
          tma_load_3d(&tmap_k, kv_full(stage), sk, col, key0, g);
          for (int kb = 0; kb < 2; ++kb)            // V as the MN-major B operand: [64 keys x 64 channels] boxes
            tma_load_3d(&tmap_v, kv_full(stage), sv + kb * 8192, col, key0 + kb * 64, g);
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(1, 1, 1, false, false, kEvTile, kEvTile);
      const uint32_t idesc_pv = umma_idesc(1, 1, 1, false, true, kEvTile, kEvD);
      constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);        // K-major, 128-byte swizzle
      constexpr uint64_t kDescMN = umma_desc_hi_lo(8192, 1024);     // MN-major (a single 64-channel chunk)
      uint32_t t = 0;
      uint32_t cnt[2] = {0, 0};
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        mbar_wait(q_full, (uint32_t)it & 1u);
        auto issue_s = [&](int x, uint32_t tt) {          // S_x = Q[:, 32 x : 32 x + 32] K[:, 32 x : 32 x + 32]^T: the two K-steps of head x
          const uint32_t sk = smem_kv + (tt & 1) * S::kStageBytes;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t off = (uint32_t)(2 * x + k) * 32u;
            umma_f16<1>(tmem_s(x), umma_desc(smem_q + off, kDescK), umma_desc(sk + off, kDescK), idesc_s, k != 0 ? 1u : 0u);
          }
          umma_commit<1>(s_full(x));
        };
        mbar_wait(kv_full(t & 1), (t >> 1) & 1u);
        tcgen05_fence_after();
        issue_s(0, t); issue_s(1, t);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int stage = t & 1;
          const uint32_t sv = smem_kv + stage * S::kStageBytes + S::kKBytes;
          for (int x = 0; x < 2; ++x) {
            mbar_wait(p_full(x), cnt[x] & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < kEvTile / 16; ++k) {
              const uint32_t b_off = (k / 4) * 8192 + (k % 4) * 2048;
              umma_f16_ts(tmem_o(x), tmem_s(x) + k * 8, umma_desc(sv + b_off, kDescMN), idesc_pv, (j | k) != 0 ? 1u : 0u);
            }
            umma_commit<1>(pv_full(x));
            ++cnt[x];
            if (j + 1 < n_kv) {
              mbar_wait(kv_full((t + 1) & 1), ((t + 1) >> 1) & 1u);
              tcgen05_fence_after();
              issue_s(x, t + 1);
            }
          }
          umma_commit<1>(kv_empty(stage));
        }
        umma_commit<1>(q_empty);
      }
    }
  } else {
    // ======================================================================================= softmax / epilogue
    float (*s_xchg)[2][kEvTile] = reinterpret_cast<float (*)[2][kEvTile]>(smem_gen + (smem_xchg - smem_base));
    const uint32_t q = warp & 3u;
    const int half = (int)((warp - 2u) >> 2);
    const int row = (int)(q * 32u + lane);
    const uint32_t lane_addr = (q * 32u) << 16;
    constexpr float kRescaleThreshold = 8.f;
    constexpr float kLog2e = 1.4426950408889634f;
    uint32_t xc = 0;
    uint32_t cnt[2] = {0, 0};
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const int g = item_g(w), hp = item_hp(w), q_tile = item_qt(w);
      const int row_g = q_tile * kEvTile + row;
      const int row_c = min(row_g, Sq - 1);                 // rows past Sq of a ragged tile: read a valid bias row, never stored
      float m_ref[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
      const float* mb_row = prm.mask_bias ? prm.mask_bias + (size_t)g * Sk : nullptr;
      for (int j = 0; j < n_kv; ++j) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          const int head = 2 * hp + x;
          const int col0 = j * kEvTile + half * 64;
          mbar_wait(s_full(x), cnt[x] & 1u);
          tcgen05_fence_after();
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(tmem_s(x) + lane_addr + half * 64, r0);
          tmem_ld_32x32b_x32(tmem_s(x) + lane_addr + half * 64 + 32, r1);
          tmem_ld_wait();
          float sc[64];                                       // logits in log2 units
#pragma unroll
          for (int i = 0; i < 64; ++i) sc[i] = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]) * prm.scale_log2;
          const bool full_tile = col0 + 64 <= Sk;
          if (prm.pair_bias) {
            const __nv_bfloat16* pb = prm.pair_bias + (((size_t)(g / prm.groups_per_pair) * H + head) * Sq + row_c) * Sk + col0;
            if (full_tile && (Sk % 8 == 0)) {
#pragma unroll
              for (int v8 = 0; v8 < 8; ++v8) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(pb) + v8);
                const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(b2[e]);
                  sc[v8 * 8 + 2 * e] = fmaf(f.x, kLog2e, sc[v8 * 8 + 2 * e]);
                  sc[v8 * 8 + 2 * e + 1] = fmaf(f.y, kLog2e, sc[v8 * 8 + 2 * e + 1]);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) if (col0 + i < Sk) sc[i] = fmaf(__bfloat162float(pb[i]), kLog2e, sc[i]);
            }
          }
          if (mb_row) {
            if (full_tile && (Sk % 4 == 0)) {
#pragma unroll
              for (int v4 = 0; v4 < 16; ++v4) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(mb_row + col0) + v4);
                sc[v4 * 4 + 0] = fmaf(f.x, kLog2e, sc[v4 * 4 + 0]); sc[v4 * 4 + 1] = fmaf(f.y, kLog2e, sc[v4 * 4 + 1]);
                sc[v4 * 4 + 2] = fmaf(f.z, kLog2e, sc[v4 * 4 + 2]); sc[v4 * 4 + 3] = fmaf(f.w, kLog2e, sc[v4 * 4 + 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 64; ++i) if (col0 + i < Sk) sc[i] = fmaf(__ldg(mb_row + col0 + i), kLog2e, sc[i]);
            }
          }
          if (!full_tile) {
#pragma unroll
            for (int i = 0; i < 64; ++i) if (col0 + i >= Sk) sc[i] = -INFINITY;
          }
          float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 64; ++i) mx[i & 3] = fmaxf(mx[i & 3], sc[i]);
          const float m_part = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
          s_xchg[xc & 1][half][row] = m_part;
          asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
          const float m_new = fmaxf(m_part, s_xchg[xc & 1][half ^ 1][row]);
          ++xc;
          const bool want = (m_new > m_ref[x] + kRescaleThreshold) || (m_ref[x] == -INFINITY && m_new != -INFINITY);
          const bool rescale = __any_sync(0xffffffffu, want);
          if (rescale) {
            const float m_next = fmaxf(m_ref[x], m_new);
            const float f = (m_ref[x] == -INFINITY) ? 0.f : exp2f(m_ref[x] - m_next);
            if (j > 0) {
              mbar_wait(pv_full(x), (cnt[x] - 1u) & 1u);
              tcgen05_fence_after();
              uint32_t r[32];                                 // this half rescales 32 of O_x's 64 columns
              tmem_ld_32x32b_x32(tmem_o(x) + lane_addr + half * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
              tmem_st_32x32b_x32(tmem_o(x) + lane_addr + half * 32, r);
            }
            l[x] *= f;
            m_ref[x] = m_next;
          }
          const float m_use = (m_ref[x] == -INFINITY) ? 0.f : m_ref[x];
          float ls[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t pw[32];
#pragma unroll
          for (int gi = 0; gi < 8; ++gi) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { pv[i] = exp2f(sc[gi * 8 + i] - m_use); ls[i & 3] += pv[i]; }
            pw[gi * 4 + 0] = pack_bf16x2(pv[0], pv[1]); pw[gi * 4 + 1] = pack_bf16x2(pv[2], pv[3]);
            pw[gi * 4 + 2] = pack_bf16x2(pv[4], pv[5]); pw[gi * 4 + 3] = pack_bf16x2(pv[6], pv[7]);
          }
          l[x] += (ls[0] + ls[1]) + (ls[2] + ls[3]);
          tmem_st_32x32b_x32(tmem_s(x) + lane_addr + half * 32, pw);
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(x));
          ++cnt[x];
        }
      }
      // item epilogue: O_x / l, gate, store this thread's 16 of head x's 32 channels
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int head = 2 * hp + x;
        mbar_wait(pv_full(x), (cnt[x] - 1u) & 1u);
        tcgen05_fence_after();
        s_xchg[xc & 1][half][row] = l[x];
        asm volatile("bar.sync %0, 64;" ::"r"(2u + q) : "memory");
        const float l_all = l[x] + s_xchg[xc & 1][half ^ 1][row];
        ++xc;
        const float inv = l_all > 0.f ? 1.f / l_all : 0.f;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_o(x) + lane_addr + kEvHead * x, r);       // head x's 32 channels
        tmem_ld_wait();
        if (row_g < Sq) {
          const size_t base = (((size_t)g * Sq + row_g) * H + head) * kEvHead + 16 * half;
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(half ? r[16 + i] : r[i]) * inv;          // static register indices
          if (prm.gate) {
            const uint4 g0 = __ldg(reinterpret_cast<const uint4*>(prm.gate + base)), g1 = __ldg(reinterpret_cast<const uint4*>(prm.gate + base) + 1);
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&g0);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&g1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fa = __bfloat1622float2(a2[e]), fb = __bfloat1622float2(b2[e]);
              o[2 * e] *= 1.f / (1.f + __expf(-fa.x)); o[2 * e + 1] *= 1.f / (1.f + __expf(-fa.y));
              o[8 + 2 * e] *= 1.f / (1.f + __expf(-fb.x)); o[8 + 2 * e + 1] *= 1.f / (1.f + __expf(-fb.y));
            }
          }
          uint4 v0, v1;
          v0.x = pack_bf16x2(o[0], o[1]); v0.y = pack_bf16x2(o[2], o[3]); v0.z = pack_bf16x2(o[4], o[5]); v0.w = pack_bf16x2(o[6], o[7]);
          v1.x = pack_bf16x2(o[8], o[9]); v1.y = pack_bf16x2(o[10], o[11]); v1.z = pack_bf16x2(o[12], o[13]); v1.w = pack_bf16x2(o[14], o[15]);
          uint4* dst = reinterpret_cast<uint4*>(prm.out + base);
          dst[0] = v0; dst[1] = v1;
          if (prm.lse != nullptr && half == 0)
            prm.lse[((size_t)g * H + head) * Sq + row_g] = (l_all > 0.f) ? (m_ref[x] + log2f(l_all)) * 0.6931471805599453f : -INFINITY;
        }
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}


// =====================================================================================================================================
// Backward.  Same transposed five-product formulation and two-deep software pipeline as attention_bwd_sm100.cu (keys on the TMEM lanes,
// 64-query inner tiles, P^T through tensor memory, dQ by TMA reduce-add), specialised for a 32-wide head: one CTA = (group, head, key
// tile); the [.. x 64] TMA tiles carry the head PAIR, the two UMMA K-steps of this head select its half for S^T / dP^T, the N = 64
// products (dV, dK) and the M = 128 product (dQ^T) compute the other head's columns / lanes as a by-product that is never read.
//   logits^T = scale K Q^T + mask_bias[k] + pair_bias[q, k]      P^T = exp2(logits^T log2e - lse2[q])
//   dV  += P^T dO'                 dO' = dOut * sigmoid(gate)   (written once by the prep kernel, with dGate and delta = rowsum(dOut o out))
//   dS^T = P^T o (dP^T - delta[q])      dP^T = V dO'^T
//   dK  += scale dS^T Q      dQ^T = scale K^T dS^T      dPairBias[q, k] += dS (fp32 red.add: the groups that share a pair bias collide here)
constexpr int kEbThreads = 448;       // warp 0 TMA, warp 1 MMA, warps 2-9 math (2 column halves x 4 lane quarters), warps 10-13 dQ^T drain
constexpr int kEbKv = 128, kEbQt = 64;

struct EbSmem {
  static constexpr int kKBytes = kEbKv * kEvD * 2;          // 16 KB: K / V tile of the head pair
  static constexpr int kQBytes = kEbQt * kEvD * 2;          //  8 KB: Q / dO' tile
  static constexpr int kDsBytes = kEbKv * kEbQt * 2;        // 16 KB
  static constexpr int kStatBytes = 2 * kEbQt * 4;
  static constexpr int kStages = 3;
  static constexpr int kDqBytes = kEbQt * kEvHead * 4;      //  8 KB: dQ tile [64 q][32 d] fp32
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kKBytes;             // V right behind K: the dQ^T product's second (unused) 64-channel chunk reads it
  static constexpr int kOffQ = kOffV + kKBytes;
  static constexpr int kOffDo = kOffQ + kStages * kQBytes;
  static constexpr int kOffDs = kOffDo + kStages * kQBytes;
  static constexpr int kOffDq = kOffDs + 2 * kDsBytes;
  static constexpr int kOffStat = kOffDq + kDqBytes;
  static constexpr int kOffBar = kOffStat + kStages * kStatBytes;
  static constexpr int kTotal = kOffBar + 256 + 1024;
};

struct EbParams {
  const float* lse2;                 // [G*H, Sq_pad] log2 units, +inf in the padding
  const float* delta;                // [G*H, Sq_pad]
  const float* mask_bias;            // [G, Sk] or null
  const __nv_bfloat16* pair_bias;    // [G / gpp, H, Sq, Sk] or null
  float* dpair;                      // fp32 [G / gpp, H, Sq, Sk] or null
  __nv_bfloat16* dk;                 // [G, Sk, H, 32]
  __nv_bfloat16* dv;
  int G, Sq, Sk, H, Sq_pad, gpp;
  float scale, scale_log2;
};

__global__ void __launch_bounds__(kEbThreads, 1)
evoformer_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                          const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ EbParams prm) {
  using S = EbSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t s_k = smem_base + S::kOffK, s_v = smem_base + S::kOffV, s_q = smem_base + S::kOffQ, s_do = smem_base + S::kOffDo;
  const uint32_t s_ds = smem_base + S::kOffDs, s_dq = smem_base + S::kOffDq, s_stat = smem_base + S::kOffStat;
  const uint32_t s_bar = smem_base + S::kOffBar;
  auto bar = [&](int i) { return s_bar + 8u * i; };
  const uint32_t kv_full = bar(0), dkv_full = bar(1), tmem_slot = bar(2);
  auto q_full = [&](int s) { return bar(3 + s); };
  auto q_empty = [&](int s) { return bar(6 + s); };
  auto s_full = [&](int b) { return bar(9 + b); };
  auto dp_full = [&](int b) { return bar(11 + b); };
  auto dp_free = [&](int b) { return bar(13 + b); };
  auto p_ready = [&](int b) { return bar(15 + b); };
  auto ds_ready = [&](int b) { return bar(17 + b); };
  auto ds_free = [&](int b) { return bar(19 + b); };
  auto dq_full = [&](int b) { return bar(21 + b); };
  auto dq_free = [&](int b) { return bar(23 + b); };

  const uint32_t warp = warp_id(), lane = lane_id();
  const int n_kv_tiles = (prm.Sk + kEbKv - 1) / kEbKv;
  const int gh = blockIdx.x / n_kv_tiles;            // (group, head): its key tiles are neighbours in launch order (shared Q / dO' in L2)
  const int j = blockIdx.x - gh * n_kv_tiles;
  const int h = gh % prm.H, g = gh / prm.H;
  const int hx = h & 1, col = (h >> 1) * kEvD;        // head inside its pair; first column of the pair
  const int k0 = j * kEbKv;
  const int n_iter = (prm.Sq + kEbQt - 1) / kEbQt;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_dq);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1); mbar_init(dkv_full, 1);
      for (int s2 = 0; s2 < S::kStages; ++s2) { mbar_init(q_full(s2), 1); mbar_init(q_empty(s2), 1); }
      for (int s2 = 0; s2 < 2; ++s2) {
        mbar_init(s_full(s2), 1); mbar_init(dp_full(s2), 1); mbar_init(dp_free(s2), 8); mbar_init(p_ready(s2), 8);
        mbar_init(ds_ready(s2), 8); mbar_init(ds_free(s2), 1); mbar_init(dq_full(s2), 1); mbar_init(dq_free(s2), 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  auto t_s = [&](int x) { return tmem + 64u * x; };            // S^T[2]; also P^T (packed, first 32 columns) and later dQ^T of the same query tile
  auto t_dp = [&](int x) { return tmem + 128u + 64u * x; };
  const uint32_t t_dv = tmem + 256, t_dk = tmem + 384;          // 64 columns each; [32 hx, 32 hx + 32) are this head's channels

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * S::kKBytes);
      tma_load_3d(&tmap_k, kv_full, s_k, col, k0, g);
      tma_load_3d(&tmap_v, kv_full, s_v, col, k0, g);
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % S::kStages, q0 = it * kEbQt;
        mbar_wait(q_empty(st), (((uint32_t)(it / S::kStages)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(st), 2 * S::kQBytes + S::kStatBytes);
        tma_load_3d(&tmap_q, q_full(st), s_q + st * S::kQBytes, col, q0, g);
        tma_load_3d(&tmap_do, q_full(st), s_do + st * S::kQBytes, col, q0, g);
        const size_t so = (size_t)gh * prm.Sq_pad + q0;
        bulk_load_1d(s_stat + st * S::kStatBytes, prm.lse2 + so, kEbQt * 4, q_full(st));
        bulk_load_1d(s_stat + st * S::kStatBytes + kEbQt * 4, prm.delta + so, kEbQt * 4, q_full(st));
      }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(1, 1, 1, false, false, kEbKv, kEbQt);
      const uint32_t idesc_acc = umma_idesc(1, 1, 1, false, true, kEbKv, kEvD);
      const uint32_t idesc_dq = umma_idesc(1, 1, 1, true, true, 128, kEbQt);
      constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);
      constexpr uint64_t kDescMN8 = umma_desc_hi_lo(8192, 1024);
      constexpr uint64_t kDescMN16 = umma_desc_hi_lo(16384, 1024);
      auto issue_s_dp = [&](int it) {
        const int x = it & 1, st = it % S::kStages;
        const uint32_t sq = s_q + st * S::kQBytes, sdo = s_do + st * S::kQBytes;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {                 // the two K-steps (32 channels) of this head inside the pair's 128-byte rows
          const uint32_t off = (uint32_t)(2 * hx + kk) * 32u;
          umma_f16<1>(t_s(x), umma_desc(s_k + off, kDescK), umma_desc(sq + off, kDescK), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit<1>(s_full(x));
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint32_t off = (uint32_t)(2 * hx + kk) * 32u;
          umma_f16<1>(t_dp(x), umma_desc(s_v + off, kDescK), umma_desc(sdo + off, kDescK), idesc_s, kk != 0 ? 1u : 0u);
        }
        umma_commit<1>(dp_full(x));
      };
      mbar_wait(kv_full, 0);
      for (int it = 0; it < 2 && it < n_iter; ++it) {
        mbar_wait(q_full(it), 0);
        tcgen05_fence_after();
        issue_s_dp(it);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int x = it & 1, st = it % S::kStages;
        const uint32_t ph = ((uint32_t)it >> 1) & 1u;
        const uint32_t sdo = s_do + st * S::kQBytes, sq = s_q + st * S::kQBytes, sds = s_ds + x * S::kDsBytes;
        mbar_wait(p_ready(x), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < kEbQt / 16; ++kk)
          umma_f16_ts(t_dv, t_s(x) + kk * 8, umma_desc(sdo + kk * 2048, kDescMN8), idesc_acc, (it | kk) != 0 ? 1u : 0u);
        mbar_wait(ds_ready(x), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < kEbQt / 16; ++kk)
          umma_f16<1>(t_dk, umma_desc(sds + kk * 32, kDescK), umma_desc(sq + kk * 2048, kDescMN8), idesc_acc, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < kEbKv / 16; ++kk)
          umma_f16<1>(t_s(x), umma_desc(s_k + kk * 2048, kDescMN16), umma_desc(sds + kk * 2048, kDescMN16), idesc_dq, kk != 0 ? 1u : 0u);
        umma_commit<1>(dq_full(x));
        umma_commit<1>(ds_free(x));
        umma_commit<1>(q_empty(st));
        if (it + 2 < n_iter) {
          mbar_wait(q_full((it + 2) % S::kStages), ((uint32_t)((it + 2) / S::kStages)) & 1u);
          mbar_wait(dq_free(x), ph);
          mbar_wait(dp_free(x), ph);
          tcgen05_fence_after();
          issue_s_dp(it + 2);
        }
      }
      umma_commit<1>(dkv_full);
    }
  } else if (warp < 10) {
    // ======================================================================================= softmax-gradient math
    constexpr float kLog2e = 1.4426950408889634f;
    const uint32_t quarter = warp & 3u;
    const int half = (int)((warp - 2u) >> 2);
    const int k_row = (int)(quarter * 32u + lane);
    const int kg = k0 + k_row;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const bool k_valid = kg < prm.Sk;
    const int kc = k_valid ? kg : prm.Sk - 1;               // clamped key index for bias loads
    const uint32_t row_off = (uint32_t)(k_row / 8) * 1024u + (uint32_t)(k_row % 8) * 128u;
    const float mb2 = prm.mask_bias ? __ldg(prm.mask_bias + (size_t)g * prm.Sk + kc) * kLog2e : 0.f;
    const size_t pair_base = ((size_t)(g / prm.gpp) * prm.H + h) * prm.Sq;     // row index base of [.., Sq, Sk]
    const float inv_scale = 1.f / prm.scale;
    for (int it = 0; it < n_iter; ++it) {
      const int x = it & 1, st = it % S::kStages;
      const uint32_t ph = ((uint32_t)it >> 1) & 1u;
      const int q0 = it * kEbQt + 32 * half;
      mbar_wait(s_full(x), ph);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_s(x) + lane_addr + 32 * half, r);
      tmem_ld_wait();
      mbar_wait(q_full(st), ((uint32_t)(it / S::kStages)) & 1u);
      const float4* lse4 = reinterpret_cast<const float4*>(smem_gen + (s_stat - smem_base) + st * S::kStatBytes) + 8 * half;
      const float4* dl4 = lse4 + kEbQt / 4;
      float p[32];                                            // P (unscaled probabilities)
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 l = lse4[c4];
        const float ls[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c4 * 4 + e;
          float bias2 = mb2;
          if (prm.pair_bias) {
            const int qc = min(q0 + c, prm.Sq - 1);
            bias2 = fmaf(__bfloat162float(prm.pair_bias[(pair_base + qc) * prm.Sk + kc]), kLog2e, bias2);
          }
          const float v = exp2f(fmaf(__uint_as_float(r[c]), prm.scale_log2, bias2) - ls[e]);     // lse2 = +inf past Sq -> 0
          p[c] = k_valid ? v : 0.f;
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(2u + quarter) : "memory");
      {
        uint32_t pw[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) pw[c >> 1] = pack_bf16x2(p[c], p[c + 1]);
        tmem_st_32x32b_x16(t_s(x) + lane_addr + 16 * half, pw);
        tmem_st_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(x));
      mbar_wait(dp_full(x), ph);
      tcgen05_fence_after();
      tmem_ld_32x32b_x32(t_dp(x) + lane_addr + 32 * half, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dp_free(x));
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 d = dl4[c4];
        const float dl[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int c = c4 * 4 + e; p[c] *= (__uint_as_float(r[c]) - dl[e]) * prm.scale; }     // dS * scale
      }
      if (prm.dpair != nullptr && k_valid) {                 // dPairBias[q, k] += dS: consecutive lanes = consecutive keys (coalesced reds)
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (q0 + c < prm.Sq) red_add_f32(prm.dpair + (pair_base + q0 + c) * prm.Sk + kg, p[c] * inv_scale);
      }
      mbar_wait(ds_free(x), ph ^ 1u);
      {
        const uint32_t dst_row = s_ds + x * S::kDsBytes + row_off;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const uint32_t chunk = (uint32_t)(4 * half + gq) ^ (uint32_t)(k_row % 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + chunk * 16u), "r"(pack_bf16x2(p[gq * 8 + 0], p[gq * 8 + 1])),
                       "r"(pack_bf16x2(p[gq * 8 + 2], p[gq * 8 + 3])), "r"(pack_bf16x2(p[gq * 8 + 4], p[gq * 8 + 5])),
                       "r"(pack_bf16x2(p[gq * 8 + 6], p[gq * 8 + 7])) : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready(x));
    }
    // ---- epilogue: dV and dK rows of this key tile, this head's 32 channels (16 per thread)
    mbar_wait(dkv_full, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      __nv_bfloat16* base = which == 0 ? prm.dv : prm.dk;
      uint32_t r[32];
      tmem_ld_32x32b_x32((which == 0 ? t_dv : t_dk) + lane_addr + kEvHead * hx, r);
      tmem_ld_wait();
      if (k_valid) {
        __nv_bfloat16* dst = base + (((size_t)g * prm.Sk + kg) * prm.H + h) * kEvHead + 16 * half;
        uint4 v0, v1;
        uint32_t rr[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) rr[i] = half ? r[16 + i] : r[i];                              // static register indices
        v0.x = pack_bf16x2(__uint_as_float(rr[0]), __uint_as_float(rr[1])); v0.y = pack_bf16x2(__uint_as_float(rr[2]), __uint_as_float(rr[3]));
        v0.z = pack_bf16x2(__uint_as_float(rr[4]), __uint_as_float(rr[5])); v0.w = pack_bf16x2(__uint_as_float(rr[6]), __uint_as_float(rr[7]));
        v1.x = pack_bf16x2(__uint_as_float(rr[8]), __uint_as_float(rr[9])); v1.y = pack_bf16x2(__uint_as_float(rr[10]), __uint_as_float(rr[11]));
        v1.z = pack_bf16x2(__uint_as_float(rr[12]), __uint_as_float(rr[13])); v1.w = pack_bf16x2(__uint_as_float(rr[14]), __uint_as_float(rr[15]));
        reinterpret_cast<uint4*>(dst)[0] = v0;
        reinterpret_cast<uint4*>(dst)[1] = v1;
      }
    }
    tcgen05_fence_before();
  } else {
    // ======================================================================================= dQ^T drain: lanes [32 hx, 32 hx + 32) are this head's channels
    const uint32_t quarter = warp & 3u;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const bool mine = (int)quarter == hx;
    const bool leader = warp == 10 && lane == 0;
    for (int it = 0; it < n_iter; ++it) {
      const int x = it & 1;
      const int q0 = it * kEbQt;
      mbar_wait(dq_full(x), ((uint32_t)it >> 1) & 1u);
      tcgen05_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(t_s(x) + lane_addr, r0);
      tmem_ld_32x32b_x32(t_s(x) + lane_addr + 32, r1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free(x));
      if (leader) tma_store_wait_read<0>();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (mine) {                                              // staging tile [64 q][32 d] fp32: this warp writes one 128-byte row per query
        const uint32_t colb = s_dq + lane * 4u;
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("st.shared.b32 [%0], %1;" ::"r"(colb + (uint32_t)c * 128u), "r"(r0[c]) : "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("st.shared.b32 [%0], %1;" ::"r"(colb + (uint32_t)(32 + c) * 128u), "r"(r1[c]) : "memory");
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        tma_reduce_add_2d(&tmap_dq, s_dq, h * kEvHead, g * prm.Sq + q0);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait<0>();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// prep: dO' = dOut * sigmoid(gate) (bf16), dGate = dOut * out * (1 - sigmoid(gate)), delta = rowsum(dOut o out), lse -> log2 units (padded)
__global__ void evoformer_bwd_prep_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ gate,
                                          const float* __restrict__ lse, __nv_bfloat16* __restrict__ do_pre, __nv_bfloat16* __restrict__ dgate,
                                          float* __restrict__ lse2, float* __restrict__ delta, int G, int Sq, int H, int Sq_pad) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (g, s_pad, h)
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)G * Sq_pad * H) return;
  const int h = (int)(w % H);
  const int s = (int)((w / H) % Sq_pad);
  const int g = (int)(w / ((int64_t)H * Sq_pad));
  const size_t oi = ((size_t)g * H + h) * Sq_pad + s;
  if (s >= Sq) {
    if (lane == 0) { lse2[oi] = INFINITY; delta[oi] = 0.f; }
    return;
  }
  const size_t e = (((size_t)g * Sq + s) * H + h) * kEvHead + lane;
  const float o = __bfloat162float(out[e]), d = __bfloat162float(dout[e]);
  float sig = 1.f;
  if (gate != nullptr) {
    sig = 1.f / (1.f + __expf(-__bfloat162float(gate[e])));
    dgate[e] = __float2bfloat16(d * o * (1.f - sig));
  }
  do_pre[e] = __float2bfloat16(d * sig);
  const float acc = warp_sum(d * o);
  if (lane == 0) {
    delta[oi] = acc;
    lse2[oi] = lse[((size_t)g * H + h) * Sq + s] * 1.4426950408889634f;
  }
}

__global__ void evoformer_f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(src + i * 8), b = *reinterpret_cast<const float4*>(src + i * 8 + 4);
    uint4 v;
    v.x = pack_bf16x2(a.x, a.y); v.y = pack_bf16x2(a.z, a.w); v.z = pack_bf16x2(b.x, b.y); v.w = pack_bf16x2(b.z, b.w);
    *reinterpret_cast<uint4*>(dst + i * 8) = v;
  }
}

}  // namespace

// q [G, Sq, H, 32], k / v [G, Sk, H, 32], out [G, Sq, H, 32] bf16 contiguous; mask_bias [G, Sk] fp32 (or null); pair_bias [G / groups_per_pair,
// H, Sq, Sk] bf16 (or null); gate [G, Sq, H, 32] bf16 pre-sigmoid (or null); lse [G, H, Sq] fp32 (or null).  H even.
cudaError_t evoformer_attention_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const float* mask_bias, const void* pair_bias,
                                    const void* gate, int G, int Sq, int Sk, int H, int groups_per_pair, float scale, cudaStream_t st) {
  if (G < 1 || Sq < 1 || Sk < 1 || H < 2 || (H & 1) || groups_per_pair < 1 || G % groups_per_pair) return cudaErrorInvalidValue;
  for (const void* p : {q, k, v, (const void*)out})
    if (p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16)) return cudaErrorInvalidValue;
  using S = EvSmem;
  const uint64_t row = (uint64_t)H * kEvHead;           // elements per (g, position)
  CUtensorMap tq, tk, tv;
  bool sw = false;
  bool ok = make_tmap_bshd(&tq, q, 1, row, Sq, G, row * 2, (uint64_t)Sq * row * 2, 64, kEvTile, &sw);
  ok &= !sw;
  ok &= make_tmap_bshd(&tk, k, 1, row, Sk, G, row * 2, (uint64_t)Sk * row * 2, 64, kEvTile, &sw);
  ok &= !sw;
  ok &= make_tmap_bshd(&tv, v, 1, row, Sk, G, row * 2, (uint64_t)Sk * row * 2, 64, 64, &sw);
  ok &= !sw;
  if (!ok) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(evoformer_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long items = (long)G * (H / 2) * ((Sq + kEvTile - 1) / kEvTile);
  const int grid = (int)(items < sms ? items : sms);
  EvParams prm{};
  prm.out = (__nv_bfloat16*)out; prm.lse = lse; prm.mask_bias = mask_bias; prm.pair_bias = (const __nv_bfloat16*)pair_bias;
  prm.gate = (const __nv_bfloat16*)gate; prm.G = G; prm.Sq = Sq; prm.Sk = Sk; prm.H = H; prm.groups_per_pair = groups_per_pair;
  prm.scale_log2 = scale * 1.4426950408889634f;
  evoformer_attn_fwd_kernel<<<grid, kEvThreads, S::kTotal, st>>>(tq, tk, tv, prm);
  return cudaGetLastError();
}

// Backward.  Workspaces: do_pre bf16 like q, dq_acc fp32 [G, Sq, H, 32], lse2 / delta fp32 [G*H, round_up(Sq, 64)].  dpair (fp32, zeroed by
// the caller) accumulates the pair-bias gradient over the groups; dgate is written only when gate != null.
cudaError_t evoformer_attention_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, const void* gate, const float* lse,
                                    const float* mask_bias, const void* pair_bias, void* dq, void* dk, void* dv, void* dgate, float* dpair,
                                    void* do_pre, float* dq_acc, float* lse2, float* delta, int G, int Sq, int Sk, int H, int groups_per_pair,
                                    float scale, cudaStream_t st) {
  if (G < 1 || Sq < 1 || Sk < 1 || H < 2 || (H & 1) || groups_per_pair < 1 || G % groups_per_pair) return cudaErrorInvalidValue;
  using S = EbSmem;
  const int Sq_pad = (Sq + kEbQt - 1) / kEbQt * kEbQt;
  const uint64_t row = (uint64_t)H * kEvHead;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)G * Sq * row * sizeof(float), st);
  if (e != cudaSuccess) return e;
  {
    const int64_t warps = (int64_t)G * Sq_pad * H;
    evoformer_bwd_prep_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, (const __nv_bfloat16*)gate, lse,
                                                                           (__nv_bfloat16*)do_pre, (__nv_bfloat16*)dgate, lse2, delta, G, Sq, H, Sq_pad);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  CUtensorMap tq, tk, tv, tdo, tdq;
  bool sw = false, ok = true;
  ok &= make_tmap_bshd(&tq, q, 1, row, Sq, G, row * 2, (uint64_t)Sq * row * 2, 64, kEbQt, &sw); ok &= !sw;
  ok &= make_tmap_bshd(&tk, k, 1, row, Sk, G, row * 2, (uint64_t)Sk * row * 2, 64, kEbKv, &sw); ok &= !sw;
  ok &= make_tmap_bshd(&tv, v, 1, row, Sk, G, row * 2, (uint64_t)Sk * row * 2, 64, kEbKv, &sw); ok &= !sw;
  ok &= make_tmap_bshd(&tdo, do_pre, 1, row, Sq, G, row * 2, (uint64_t)Sq * row * 2, 64, kEbQt, &sw); ok &= !sw;
  ok &= make_tmap_2d_plain(&tdq, dq_acc, 3, row, (uint64_t)G * Sq, row * 4, kEvHead, kEbQt);
  if (!ok) return cudaErrorInvalidValue;
  EbParams prm{};
  prm.lse2 = lse2; prm.delta = delta; prm.mask_bias = mask_bias; prm.pair_bias = (const __nv_bfloat16*)pair_bias; prm.dpair = dpair;
  prm.dk = (__nv_bfloat16*)dk; prm.dv = (__nv_bfloat16*)dv;
  prm.G = G; prm.Sq = Sq; prm.Sk = Sk; prm.H = H; prm.Sq_pad = Sq_pad; prm.gpp = groups_per_pair;
  prm.scale = scale; prm.scale_log2 = scale * 1.4426950408889634f;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(evoformer_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal)) != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_kv = (Sk + kEbKv - 1) / kEbKv;
  evoformer_attn_bwd_kernel<<<(unsigned)((int64_t)n_kv * G * H), kEbThreads, S::kTotal, st>>>(tq, tk, tv, tdo, tdq, prm);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  evoformer_f32_to_bf16_kernel<<<1184, 256, 0, st>>>(dq_acc, (__nv_bfloat16*)dq, (int64_t)G * Sq * (int64_t)row / 8);
  return cudaGetLastError();
}

}  // namespace pfx

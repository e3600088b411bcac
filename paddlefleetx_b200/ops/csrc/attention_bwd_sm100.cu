// Flash-attention backward on tcgen05 / TMEM / TMA (sm_100a), bf16, head_dim 128, causal or full, optional dropout.
//
// One CTA owns a tile of 128 keys of one (batch, head) and walks the 64-query tiles that attend to it.  Everything is
// computed TRANSPOSED (keys on the 128 TMEM lanes, queries / channels on the columns) so that every UMMA runs at M = 128
// while the query tile — the N or K extent of the five products — is only 64 wide, which is what lets S^T (double
// buffered), dP^T, dQ^T and the two persistent accumulators dV, dK share the 512 TMEM columns:
//
//     S^T  = K_j Q_i^T                      [128 k x 64 q]    A = K_j  (K-major),   B = Q_i  (K-major)       cols   0..127 (2 buffers)
//     dP^T = V_j dO_i^T                     [128 k x 64 q]    A = V_j  (K-major),   B = dO_i (K-major)       cols 128..255 (2 buffers)
//     P^T  = exp2(S^T * c - lse_i)   masked, dropout applied; written as packed bf16 over the first 32 columns of its S^T buffer
//     dS^T = P^T o (drop(dP^T) - delta_i) * scale     -> shared memory (bf16, swizzled, 2 buffers)
//     dV  += P_drop^T dO_i                  [128 k x 128 d]   A = P^T  (TENSOR MEMORY),  B = dO_i (MN-major)  cols 256..383
//     dK  += dS^T Q_i                       [128 k x 128 d]   A = dS^T (K-major),   B = Q_i  (MN-major)      cols 384..511
//     dQ^T = K_j^T dS^T                     [128 d x 64 q]    A = K_j  (MN-major),  B = dS^T (MN-major)      into the S^T buffer of tile i
// The MMA warp runs TWO query tiles ahead (S^T / dP^T of tile i+2 are issued right behind the three products of tile i; Q_i / dO_i ride a
// 3-stage ring), so the tensor core and the math warps only meet on data, not on each other's schedule — the first version alternated
// them and spent ~7000 cycles per tile against 1280 cycles of tensor work.
//
// The same shared-memory tiles serve as K-major and MN-major operands (a 128-byte-swizzled [rows x 64] panel is both), so Q_i,
// dO_i, K_j, V_j are loaded once by TMA straight from the framework layout ([B, S, H, D] views, packed QKV included) and P^T / dS^T
// are written once by the threads that produce them (P^T never leaves the tensor-memory / register domain).  dQ is accumulated across key tiles in an fp32 workspace: the drain warps move
// each dQ^T tile TMEM -> registers -> (transposed) shared memory and ONE TMA reduce-add (cp.reduce.async.bulk.tensor) folds it into
// global memory — per-lane red.global instructions cost ~1.3 cycles per lane on the SM and were 5x the tensor-core time of a tile
// (first version: 0.55 ms at B8 S1024 H32); the workspace is converted to bf16 afterwards; lse (log2 units) and delta = rowsum(dO o O) come from a small preprocessing kernel.
//
// Warp roles (448 threads): warp 0 TMA producer (K/V once, then a 2-stage ring of Q_i / dO_i / lse_i / delta_i), warp 1 MMA
// issuer, warps 2-9 softmax-gradient math (two threads per key row, 32 query columns each), warps 10-13 drain dQ^T.
// Reference call site: flash_attention in hybrid_model.py:284-301 (FlashAttention-2 library, mma.sync on Ampere).
#include <cstdio>

#include "pfx_ptx.cuh"
#include "pfx_common.cuh"
#include "pfx_gemm.h"
#include "pfx_attn.h"
#include "pfx_attn.cuh"

namespace pfx {

namespace {

constexpr int kBwThreads = 448;
constexpr int kKv = 128;            // keys per CTA
constexpr int kQt = 64;             // queries per inner iteration

template <int kHd>                  // head dim: 128 (two 64-channel panels per tile) or 64 (one)
struct BwSmem {
  static constexpr int kPanels = kHd / 64;
  static constexpr int kKBytes = kKv * kHd * 2;          // [128 x 64] panels of 16 KB
  static constexpr int kQBytes = kQt * kHd * 2;          // [64 x 64] panels of 8 KB
  static constexpr int kDsBytes = kKv * kQt * 2;         // 16 KB: one [128 x 64] panel
  static constexpr int kStatBytes = 2 * kQt * 4;         // lse2 + delta of one query tile
  static constexpr int kStages = 3;                      // Q_i / dO_i / stats ring: the MMA warp runs two query tiles ahead of the math warps
  static constexpr int kDqBytes = kQt * kHd * 4;         // dQ tile [64 q][kHd d] fp32, row-major (TMA reduce source)
  static constexpr int kOffK = 0;
  static constexpr int kOffV = kOffK + kKBytes;
  static constexpr int kOffQ = kOffV + kKBytes;
  static constexpr int kOffDo = kOffQ + kStages * kQBytes;
  static constexpr int kOffDs = kOffDo + kStages * kQBytes;      // 2 buffers
  static constexpr int kOffDq = kOffDs + 2 * kDsBytes;
  static constexpr int kOffStat = kOffDq + kDqBytes;
  static constexpr int kOffBar = kOffStat + kStages * kStatBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kUsed = kOffBar + kBarBytes;
  static constexpr int kTotal = kUsed + 1024;            // alignment slack
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

struct BwOut { void* ptr; int64_t sb, ss, sh; };

struct BwParams {
  const float* lse2;       // [B*H, Sq_pad], log2 units, +inf in the padding
  const float* delta;      // [B*H, Sq_pad]
  float* dq_acc;           // [B, Sq, H, 128] fp32
  BwOut dk, dv;
  int B, Sq, Sk, H, Sq_pad;
  int hs_q, hs_k, hs_v, hs_do;        // head strides (elements) of the four TMA views
  uint32_t swapped;                    // bit t: tensor map t (q, k, v, do) has {cols, B, S} coordinate order
  float scale, scale_log2;
  int causal;
  uint32_t drop_thresh16;              // 0 = no dropout
  float inv_keep;
  uint64_t seed;
};

__device__ __forceinline__ void tma_tile(const CUtensorMap* m, bool swapped, uint32_t bar, uint32_t dst, int col, int s, int b) {
  if (swapped) tma_load_3d(m, bar, dst, col, b, s); else tma_load_3d(m, bar, dst, col, s, b);
}

template <int kHd>
__global__ void __launch_bounds__(kBwThreads, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     const __grid_constant__ CUtensorMap tmap_do, const __grid_constant__ CUtensorMap tmap_dq, const __grid_constant__ BwParams prm) {
  using S = BwSmem<kHd>;
  constexpr int kPanels = S::kPanels;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t s_k = smem_base + S::kOffK, s_v = smem_base + S::kOffV, s_q = smem_base + S::kOffQ, s_do = smem_base + S::kOffDo;
  const uint32_t s_ds = smem_base + S::kOffDs, s_dq = smem_base + S::kOffDq, s_stat = smem_base + S::kOffStat;
  const uint32_t s_bar = smem_base + S::kOffBar;
  auto bar = [&](int i) { return s_bar + 8u * i; };
  const uint32_t kv_full = bar(0), dkv_full = bar(1), tmem_slot = bar(2);
  auto q_full = [&](int s) { return bar(3 + s); };           // 3 stages
  auto q_empty = [&](int s) { return bar(6 + s); };
  auto s_full = [&](int b) { return bar(9 + b); };           // everything below: 2 buffers, indexed by (iteration & 1)
  auto dp_full = [&](int b) { return bar(11 + b); };
  auto dp_free = [&](int b) { return bar(13 + b); };
  auto p_ready = [&](int b) { return bar(15 + b); };
  auto ds_ready = [&](int b) { return bar(17 + b); };
  auto ds_free = [&](int b) { return bar(19 + b); };
  auto dq_full = [&](int b) { return bar(21 + b); };
  auto dq_free = [&](int b) { return bar(23 + b); };

  const uint32_t warp = warp_id(), lane = lane_id();
  // CTA order: the key tiles of ONE (batch, head) are neighbours in launch order, so the CTAs that stream the same Q / dO tiles and reduce
  // into the same dQ rows run at the same time and meet in L2 (ncu of the key-tile-major order: 1.2 GB of DRAM reads for 0.34 GB of
  // operands, 25 % L2 hit rate).  Within a head the tiles ascend: under a causal mask the first ones carry the most work.
  const int n_kv_tiles = (prm.Sk + kKv - 1) / kKv;
  const int bh = blockIdx.x / n_kv_tiles;
  const int j = blockIdx.x - bh * n_kv_tiles;
  const int h = bh % prm.H, b = bh / prm.H;
  const int k0 = j * kKv;
  const int off = prm.Sk - prm.Sq;                // query q attends keys <= q + off
  const int n_q_tiles = (prm.Sq + kQt - 1) / kQt;
  const int i_begin = prm.causal ? max(0, k0 - off) / kQt : 0;
  const int n_iter = n_q_tiles - i_begin;         // >= 1 for every key tile that holds a valid key

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_dq);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(kv_full, 1); mbar_init(dkv_full, 1);
      for (int s = 0; s < S::kStages; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
      for (int s = 0; s < 2; ++s) {
        mbar_init(s_full(s), 1); mbar_init(dp_full(s), 1); mbar_init(dp_free(s), 8); mbar_init(p_ready(s), 8);
        mbar_init(ds_ready(s), 8); mbar_init(ds_free(s), 1); mbar_init(dq_full(s), 1); mbar_init(dq_free(s), 4);
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
  // TMEM columns: S^T[2] (each also hosts P^T as packed bf16 in its first 32 columns, and later the dQ^T of the same query tile), dP^T[2], dV, dK
  auto t_s = [&](int x) { return tmem + 64u * x; };
  auto t_dp = [&](int x) { return tmem + 128u + 64u * x; };
  const uint32_t t_dv = tmem + 256, t_dk = tmem + 384;

  if (warp == 0) {
    // ======================================================================================= TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(kv_full, 2 * S::kKBytes);
      for (int p = 0; p < kPanels; ++p) {
        tma_tile(&tmap_k, prm.swapped & 2u, kv_full, s_k + p * 16384, h * prm.hs_k + p * 64, k0, b);
        tma_tile(&tmap_v, prm.swapped & 4u, kv_full, s_v + p * 16384, h * prm.hs_v + p * 64, k0, b);
      }
      for (int it = 0; it < n_iter; ++it) {
        const int st = it % S::kStages, q0 = (i_begin + it) * kQt;
        mbar_wait(q_empty(st), (((uint32_t)(it / S::kStages)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(q_full(st), 2 * S::kQBytes + S::kStatBytes);
        for (int p = 0; p < kPanels; ++p) {
          tma_tile(&tmap_q, prm.swapped & 1u, q_full(st), s_q + st * S::kQBytes + p * 8192, h * prm.hs_q + p * 64, q0, b);
          tma_tile(&tmap_do, prm.swapped & 8u, q_full(st), s_do + st * S::kQBytes + p * 8192, h * prm.hs_do + p * 64, q0, b);
        }
        const size_t so = (size_t)bh * prm.Sq_pad + q0;
        bulk_load_1d(s_stat + st * S::kStatBytes, prm.lse2 + so, kQt * 4, q_full(st));
        bulk_load_1d(s_stat + st * S::kStatBytes + kQt * 4, prm.delta + so, kQt * 4, q_full(st));
      }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    // Software pipeline, two query tiles deep: while the math warps turn S^T_i / dP^T_i into P^T_i / dS^T_i, the tensor core already holds
    // S^T_{i+1} / dP^T_{i+1}, and it computes S^T_{i+2} / dP^T_{i+2} right behind the three products of tile i.
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(1, 1, 1, false, false, kKv, kQt);      // S^T, dP^T: both operands K-major
      const uint32_t idesc_acc = umma_idesc(1, 1, 1, false, true, kKv, kHd);     // dV, dK: A K-major (tensor / shared memory), B MN-major
      // dQ^T: both MN-major.  M is 128 even for a 64-wide head: the A descriptor's second 64-channel chunk (16 KB further) then lands on the
      // V tile — valid shared memory whose product fills accumulator lanes 64..127, which nobody reads.
      const uint32_t idesc_dq = umma_idesc(1, 1, 1, true, true, 128, kQt);
      constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);
      constexpr uint64_t kDescMN8 = umma_desc_hi_lo(8192, 1024);                 // 64-wide MN chunks 8 KB apart (Q_i / dO_i panels)
      constexpr uint64_t kDescMN16 = umma_desc_hi_lo(16384, 1024);               // ... 16 KB apart (K_j panels; dS^T has a single chunk)
      auto issue_s_dp = [&](int it) {                    // S^T_it and dP^T_it into buffer it & 1 (stage it % 3 has landed)
        const int x = it & 1, st = it % S::kStages;
        const uint32_t sq = s_q + st * S::kQBytes, sdo = s_do + st * S::kQBytes;
#pragma unroll
        for (int kk = 0; kk < kHd / 16; ++kk)
          umma_f16<1>(t_s(x), umma_desc(s_k + (kk / 4) * 16384 + (kk % 4) * 32, kDescK), umma_desc(sq + (kk / 4) * 8192 + (kk % 4) * 32, kDescK),
                      idesc_s, kk != 0 ? 1u : 0u);
        umma_commit<1>(s_full(x));
#pragma unroll
        for (int kk = 0; kk < kHd / 16; ++kk)
          umma_f16<1>(t_dp(x), umma_desc(s_v + (kk / 4) * 16384 + (kk % 4) * 32, kDescK), umma_desc(sdo + (kk / 4) * 8192 + (kk % 4) * 32, kDescK),
                      idesc_s, kk != 0 ? 1u : 0u);
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
        // dV += P_drop^T dO   (A = P^T read from tensor memory: packed bf16 in the first 32 columns of S^T[x])
        mbar_wait(p_ready(x), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < kQt / 16; ++kk)
          umma_f16_ts(t_dv, t_s(x) + kk * 8, umma_desc(sdo + kk * 2048, kDescMN8), idesc_acc, (it | kk) != 0 ? 1u : 0u);
        // dK += dS^T Q ; dQ^T = K^T dS^T  (into S^T[x]: its scores and P^T have been consumed by the products issued above)
        mbar_wait(ds_ready(x), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < kQt / 16; ++kk)
          umma_f16<1>(t_dk, umma_desc(sds + kk * 32, kDescK), umma_desc(sq + kk * 2048, kDescMN8), idesc_acc, (it | kk) != 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < kKv / 16; ++kk)
          umma_f16<1>(t_s(x), umma_desc(s_k + kk * 2048, kDescMN16), umma_desc(sds + kk * 2048, kDescMN16), idesc_dq, kk != 0 ? 1u : 0u);
        umma_commit<1>(dq_full(x));
        umma_commit<1>(ds_free(x));
        umma_commit<1>(q_empty(st));
        if (it + 2 < n_iter) {
          mbar_wait(q_full((it + 2) % S::kStages), ((uint32_t)((it + 2) / S::kStages)) & 1u);
          mbar_wait(dq_free(x), ph);                  // the drain warps have read dQ^T_it out of S^T[x]
          mbar_wait(dp_free(x), ph);                  // the math warps have read dP^T_it
          tcgen05_fence_after();
          issue_s_dp(it + 2);
        }
      }
      umma_commit<1>(dkv_full);
    }
  } else if (warp < 10) {
    // ======================================================================================= softmax-gradient math
    const uint32_t quarter = warp & 3u;                        // TMEM lane quarter this warp may touch
    const int half = (int)((warp - 2u) >> 2);                  // which 32 of the 64 query columns
    const int k_row = (int)(quarter * 32u + lane);             // key row inside the tile == TMEM lane
    const int kg = k0 + k_row;
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const bool k_valid = kg < prm.Sk;
    const uint32_t row_off = (uint32_t)(k_row / 8) * 1024u + (uint32_t)(k_row % 8) * 128u;
    const uint32_t pairs = ((uint32_t)prm.Sk + 1u) >> 1;
    const uint32_t key = attn_rng_key(prm.seed, (uint32_t)bh);
    const bool drop = prm.drop_thresh16 != 0;
    const float p_to_pdrop = prm.inv_keep / prm.scale;
    for (int it = 0; it < n_iter; ++it) {
      const int x = it & 1, st = it % S::kStages;
      const uint32_t ph = ((uint32_t)it >> 1) & 1u;
      const int q0 = (i_begin + it) * kQt + 32 * half;
      mbar_wait(s_full(x), ph);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_s(x) + lane_addr + 32 * half, r);
      tmem_ld_wait();
      mbar_wait(q_full(st), ((uint32_t)(it / S::kStages)) & 1u);   // lse / delta of this stage have landed (same barrier as the TMA tiles)
      const float4* lse4 = reinterpret_cast<const float4*>(smem_gen + (s_stat - smem_base) + st * S::kStatBytes) + 8 * half;
      const float4* dl4 = lse4 + kQt / 4;
      float p[32];                                                 // P * scale (the softmax scale is folded in here once)
      uint32_t keep = 0xFFFFFFFFu;
      // masks are decided per tile (warp-uniform branches): only diagonal / ragged tiles pay per-element compares
      const bool tile_mask = (prm.causal && (k0 + kKv - 1 > q0 + off)) || (k0 + kKv > prm.Sk);
      if (tile_mask) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 l = lse4[c4];
          const float ls[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            float v = exp2f(fmaf(__uint_as_float(r[c]), prm.scale_log2, -ls[e])) * prm.scale;
            if (!k_valid || (prm.causal && kg > q0 + c + off)) v = 0.f;
            p[c] = v;
          }
        }
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 l = lse4[c4];
          const float ls[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) p[c4 * 4 + e] = exp2f(fmaf(__uint_as_float(r[c4 * 4 + e]), prm.scale_log2, -ls[e])) * prm.scale;
        }
      }
      if (drop) {
        keep = 0u;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint32_t bits = attn_rng_pair(key, (uint32_t)(q0 + c), (uint32_t)kg >> 1, pairs);
          keep |= (attn_keep(bits, (uint32_t)kg, prm.drop_thresh16) ? 1u : 0u) << c;
        }
      }
      // P_drop^T -> tensor memory, over the first 32 columns of S^T[x] (this thread: 32 queries = 16 packed words).  Both threads of a key row
      // must have read their scores before either overwrites them: the two warps of a lane quarter meet on a named barrier.
      asm volatile("bar.sync %0, 64;" ::"r"(2u + quarter) : "memory");
      {
        uint32_t pw[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2)
          pw[c >> 1] = pack_bf16x2(((keep >> c) & 1u) ? p[c] * p_to_pdrop : 0.f, ((keep >> (c + 1)) & 1u) ? p[c + 1] * p_to_pdrop : 0.f);
        tmem_st_32x32b_x16(t_s(x) + lane_addr + 16 * half, pw);
        tmem_st_wait();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(x));
      // dS^T * scale = (P scale) o (drop(dP) - delta)
      mbar_wait(dp_full(x), ph);
      tcgen05_fence_after();
      tmem_ld_32x32b_x32(t_dp(x) + lane_addr + 32 * half, r);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dp_free(x));
      if (drop) {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 d = dl4[c4];
          const float dl[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const float dpe = ((keep >> c) & 1u) ? __uint_as_float(r[c]) * prm.inv_keep : 0.f;
            p[c] *= dpe - dl[e];
          }
        }
      } else {
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 d = dl4[c4];
          const float dl[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) { const int c = c4 * 4 + e; p[c] *= __uint_as_float(r[c]) - dl[e]; }
        }
      }
      mbar_wait(ds_free(x), ph ^ 1u);                 // the products of tile it - 2 have finished reading this dS^T buffer
      {
        const uint32_t dst_row = s_ds + x * S::kDsBytes + row_off;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t chunk = (uint32_t)(4 * half + g) ^ (uint32_t)(k_row % 8);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_row + chunk * 16u), "r"(pack_bf16x2(p[g * 8 + 0], p[g * 8 + 1])),
                       "r"(pack_bf16x2(p[g * 8 + 2], p[g * 8 + 3])), "r"(pack_bf16x2(p[g * 8 + 4], p[g * 8 + 5])),
                       "r"(pack_bf16x2(p[g * 8 + 6], p[g * 8 + 7])) : "memory");
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_ready(x));
    }
    // ---- epilogue: dV and dK of this key tile (this thread: its key row, one half of the channels)
    constexpr int kCh = kHd / 2;
    mbar_wait(dkv_full, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const BwOut& o = which == 0 ? prm.dv : prm.dk;
      const uint32_t t_acc = which == 0 ? t_dv : t_dk;
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(o.ptr) + (size_t)b * o.sb + (size_t)kg * o.ss + (size_t)h * o.sh + kCh * half;
#pragma unroll 1
      for (int c0 = 0; c0 < kCh; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_acc + lane_addr + kCh * half + c0, r);
        tmem_ld_wait();
        if (k_valid) {
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(r[e + 0]), __uint_as_float(r[e + 1]));
            v.y = pack_bf16x2(__uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
            v.z = pack_bf16x2(__uint_as_float(r[e + 4]), __uint_as_float(r[e + 5]));
            v.w = pack_bf16x2(__uint_as_float(r[e + 6]), __uint_as_float(r[e + 7]));
            *reinterpret_cast<uint4*>(dst + c0 + e) = v;
          }
        }
      }
    }
    tcgen05_fence_before();
  } else {
    // ======================================================================================= dQ^T drain: TMEM -> smem (transposed) -> TMA reduce-add
    const uint32_t quarter = warp & 3u;
    const int d = (int)(quarter * 32u + lane);                  // channel == TMEM lane
    const uint32_t lane_addr = (quarter * 32u) << 16;
    const bool leader = warp == 10 && lane == 0;
    for (int it = 0; it < n_iter; ++it) {
      const int x = it & 1;
      const int q0 = (i_begin + it) * kQt;
      mbar_wait(dq_full(x), ((uint32_t)it >> 1) & 1u);
      tcgen05_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(t_s(x) + lane_addr, r0);
      tmem_ld_32x32b_x32(t_s(x) + lane_addr + 32, r1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free(x));
      if (leader) tma_store_wait_read<0>();                     // the previous tile's reduce has finished reading the staging buffer
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t col = s_dq + (uint32_t)d * 4u;             // staging tile [64 q][kHd d] fp32: a warp writes 32 consecutive floats of one row
      if (d < kHd) {                                            // (64-wide heads: accumulator lanes 64..127 hold the unused second chunk)
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col + (uint32_t)c * (kHd * 4u)), "r"(r0[c]) : "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) asm volatile("st.shared.b32 [%0], %1;" ::"r"(col + (uint32_t)(32 + c) * (kHd * 4u)), "r"(r1[c]) : "memory");
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (leader) {
        // rows of the 2-D [B*Sq, H*128] view; query rows past Sq of a ragged tile carry exact zeros (P = 0 there), rows past the tensor are dropped
        tma_reduce_add_2d(&tmap_dq, s_dq, h * kHd, b * prm.Sq + q0);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait<0>();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem, 512);
}

// ---- preprocessing: delta = rowsum(dO o O), lse in log2 units; both padded to a multiple of 64 queries per (b, h)
template <int kHd>
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, int64_t o_sb, int64_t o_ss, int64_t o_sh, const __nv_bfloat16* __restrict__ dout,
                                     int64_t d_sb, int64_t d_ss, int64_t d_sh, const float* __restrict__ lse, float* __restrict__ lse2,
                                     float* __restrict__ delta, int B, int Sq, int H, int Sq_pad) {
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (b, s_pad, h)
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * Sq_pad * H) return;
  const int h = (int)(w % H);
  const int s = (int)((w / H) % Sq_pad);
  const int b = (int)(w / ((int64_t)H * Sq_pad));
  const size_t oi = ((size_t)b * H + h) * Sq_pad + s;
  if (s >= Sq) {
    if (lane == 0) { lse2[oi] = INFINITY; delta[oi] = 0.f; }
    return;
  }
  constexpr int kPer = kHd / 32;                         // channels per lane (4 or 2)
  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(o + b * o_sb + s * o_ss + h * o_sh + lane * kPer);
  const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(dout + b * d_sb + s * d_ss + h * d_sh + lane * kPer);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kPer / 2; ++i) {
    const float2 x = __bfloat1622float2(a2[i]), y = __bfloat1622float2(g2[i]);
    acc += x.x * y.x + x.y * y.y;
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    delta[oi] = acc;
    lse2[oi] = lse[((size_t)b * H + h) * Sq + s] * 1.4426950408889634f;
  }
}

// ---- dq: fp32 accumulator -> bf16 view
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, int64_t sb, int64_t ss, int64_t sh, int B, int Sq, int H, int kHd) {
  const int64_t n8 = (int64_t)B * Sq * H * (kHd / 8);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (kHd / 8));
    const int64_t row = i / (kHd / 8);                 // (b, s, h)
    const int h = (int)(row % H);
    const int s = (int)((row / H) % Sq);
    const int b = (int)(row / ((int64_t)H * Sq));
    const float4 x = *reinterpret_cast<const float4*>(acc + i * 8), y = *reinterpret_cast<const float4*>(acc + i * 8 + 4);
    uint4 v;
    v.x = pack_bf16x2(x.x, x.y); v.y = pack_bf16x2(x.z, x.w); v.z = pack_bf16x2(y.x, y.y); v.w = pack_bf16x2(y.z, y.w);
    *reinterpret_cast<uint4*>(dq + b * sb + s * ss + h * sh + c * 8) = v;
  }
}

bool view_ok(const AttnView& t) {
  return t.ptr != nullptr && (reinterpret_cast<uintptr_t>(t.ptr) % 16) == 0 && t.sb % 8 == 0 && t.ss % 8 == 0 && t.sh % 8 == 0;
}

bool make_map(CUtensorMap* m, const AttnView& t, int B, int S, int H, int D, uint32_t box_rows, bool* swapped) {
  const uint64_t inner = (uint64_t)(H - 1) * t.sh + D;
  return make_tmap_bshd(m, t.ptr, 1, inner, (uint64_t)S, (uint64_t)B, (uint64_t)t.ss * 2, (uint64_t)t.sb * 2, 64, box_rows, swapped);
}

}  // namespace

template <int kHd>
static cudaError_t attention_bwd_impl(const AttnView& q, const AttnView& k, const AttnView& v, const AttnView& out, const AttnView& dout, const float* lse,
                                      const AttnView& dq, const AttnView& dk, const AttnView& dv, float* dq_acc, float* lse2, float* delta, int B, int Sq,
                                      int Sk, int H, float scale, bool causal, AttnDropout drop, cudaStream_t st) {
  constexpr int D = kHd;
  using S = BwSmem<kHd>;
  const int Sq_pad = (Sq + kQt - 1) / kQt * kQt;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, (size_t)B * Sq * H * kHd * sizeof(float), st);
  if (e != cudaSuccess) return e;
  {
    const int64_t warps = (int64_t)B * Sq_pad * H;
    const int wpb = 8;
    attn_bwd_prep_kernel<kHd><<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        (const __nv_bfloat16*)out.ptr, out.sb, out.ss, out.sh, (const __nv_bfloat16*)dout.ptr, dout.sb, dout.ss, dout.sh, lse, lse2, delta, B, Sq, H, Sq_pad);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  CUtensorMap tq, tk, tv, tdo, tdq;
  bool sq = false, sk = false, sv = false, sdo = false;
  bool ok = make_map(&tq, q, B, Sq, H, D, kQt, &sq) && make_map(&tk, k, B, Sk, H, D, kKv, &sk) && make_map(&tv, v, B, Sk, H, D, kKv, &sv) &&
            make_map(&tdo, dout, B, Sq, H, D, kQt, &sdo) &&
            make_tmap_2d_plain(&tdq, dq_acc, 3, (uint64_t)H * kHd, (uint64_t)B * Sq, (uint64_t)H * kHd * 4, kHd, kQt);
  if (!ok) return cudaErrorInvalidValue;
  BwParams prm{};
  prm.lse2 = lse2; prm.delta = delta; prm.dq_acc = dq_acc;
  prm.dk = BwOut{const_cast<void*>(dk.ptr), dk.sb, dk.ss, dk.sh};
  prm.dv = BwOut{const_cast<void*>(dv.ptr), dv.sb, dv.ss, dv.sh};
  prm.B = B; prm.Sq = Sq; prm.Sk = Sk; prm.H = H; prm.Sq_pad = Sq_pad;
  prm.hs_q = (int)q.sh; prm.hs_k = (int)k.sh; prm.hs_v = (int)v.sh; prm.hs_do = (int)dout.sh;
  prm.swapped = (sq ? 1u : 0u) | (sk ? 2u : 0u) | (sv ? 4u : 0u) | (sdo ? 8u : 0u);
  prm.scale = scale; prm.scale_log2 = scale * 1.4426950408889634f;
  prm.causal = causal ? 1 : 0;
  prm.drop_thresh16 = drop.p > 0.f ? (uint32_t)(drop.p * 65536.f + 0.5f) : 0u;
  prm.inv_keep = drop.p > 0.f ? 1.f / (1.f - drop.p) : 1.f;
  prm.seed = drop.seed;
  static bool attr_set = false;
  if (!attr_set) {
    if ((e = cudaFuncSetAttribute(attention_bwd_kernel<kHd>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal)) != cudaSuccess) return e;
    attr_set = true;
  }
  const int n_kv = (Sk + kKv - 1) / kKv;
  attention_bwd_kernel<kHd><<<(unsigned)(n_kv * B * H), kBwThreads, S::kTotal, st>>>(tq, tk, tv, tdo, tdq, prm);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  attn_bwd_dq_convert_kernel<<<1184, 256, 0, st>>>(dq_acc, (__nv_bfloat16*)const_cast<void*>(dq.ptr), dq.sb, dq.ss, dq.sh, B, Sq, H, kHd);
  return cudaGetLastError();
}

cudaError_t attention_bwd(const AttnView& q, const AttnView& k, const AttnView& v, const AttnView& out, const AttnView& dout, const float* lse,
                          const AttnView& dq, const AttnView& dk, const AttnView& dv, float* dq_acc, float* lse2, float* delta, int B, int Sq,
                          int Sk, int H, int D, float scale, bool causal, AttnDropout drop, cudaStream_t st) {
  if ((D != 64 && D != 128) || B < 1 || Sq < 1 || Sk < 1 || (causal && Sk < Sq)) return cudaErrorInvalidValue;
  for (const AttnView* t : {&q, &k, &v, &out, &dout, &dq, &dk, &dv}) if (!view_ok(*t)) return cudaErrorInvalidValue;
  if (D == 128) return attention_bwd_impl<128>(q, k, v, out, dout, lse, dq, dk, dv, dq_acc, lse2, delta, B, Sq, Sk, H, scale, causal, drop, st);
  return attention_bwd_impl<64>(q, k, v, out, dout, lse, dq, dk, dv, dq_acc, lse2, delta, B, Sq, Sk, H, scale, causal, drop, st);
}

}  // namespace pfx

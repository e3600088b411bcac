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
//                          MMA consumes, running sum and the O accumulator (fp32, in registers) rescaled once per KV tile
// S for tile j+1 is issued as soon as the softmax warps have drained S_j from TMEM, so the tensor core computes the next
// scores while the softmax of the current tile is in its exp / store phase.  Causal tiles above the diagonal are never loaded.
// Output O (bf16) and the row-wise log-sum-exp (fp32, [B, H, S]) are written straight from registers.
//
// Used on the no-grad paths (prefill of generation, evaluation, vision / text encoders in inference).  Training keeps the
// library (cuDNN) kernels: a matching backward is future work.
// Reference call site: flash_attention in hybrid_model.py:284-301 (FlashAttention-2 library on Ampere mma.sync).
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
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = 1024 + kQBytes + kStages * kStageBytes + kPBytes + kBarBytes;
};

template <int kD, bool kCausal>
__global__ void __launch_bounds__(kFaThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                     __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int Sq, int Sk, int H, float scale_log2) {
  using S = FaSmem<kD>;
  constexpr int kPanels = kD / 64;                 // 64-element (128-byte) column panels of a D-wide tile
  constexpr int kTmemCols = 256;                   // S: [0,128), PV: [128, 128 + kD)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_q = smem_base;
  const uint32_t smem_kv = smem_q + S::kQBytes;
  const uint32_t smem_p = smem_kv + S::kStages * S::kStageBytes;
  const uint32_t smem_bar = smem_p + S::kPBytes;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t q_full = smem_bar;
  auto kv_full = [&](int s) { return smem_bar + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return smem_bar + 8u * (3 + s); };
  const uint32_t s_full = smem_bar + 8u * 5, s_free = smem_bar + 8u * 6, p_full = smem_bar + 8u * 7, pv_full = smem_bar + 8u * 8,
                 pv_free = smem_bar + 8u * 9;
  const uint32_t tmem_slot = smem_bar + 8u * 10;

  const uint32_t warp = warp_id(), lane = lane_id();
  const int n_q_tiles = (Sq + kFaTile - 1) / kFaTile;
  const int q_tile = n_q_tiles - 1 - (int)blockIdx.x;          // longest (most KV tiles under a causal mask) first
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * kFaTile;
  const int n_kv_all = (Sk + kFaTile - 1) / kFaTile;
  const int n_kv = kCausal ? min(n_kv_all, (q0 + kFaTile - 1 + (Sk - Sq)) / kFaTile + 1) : n_kv_all;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < S::kStages; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
      mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(pv_full, 1); mbar_init(pv_free, 8);
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
      mbar_arrive_expect_tx(q_full, S::kQBytes);
      for (int p = 0; p < kPanels; ++p) tma_load_2d(&tmap_q, q_full, smem_q + p * (kFaTile * 128), h * kD + p * 64, b * Sq + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int stage = j & 1;
        const uint32_t phase = (uint32_t)(j >> 1) & 1u;
        mbar_wait(kv_empty(stage), phase ^ 1u);
        const uint32_t sk = smem_kv + stage * S::kStageBytes, sv = sk + S::kKBytes;
        mbar_arrive_expect_tx(kv_full(stage), S::kStageBytes);
        const int key0 = b * Sk + j * kFaTile;
        for (int p = 0; p < kPanels; ++p) tma_load_2d(&tmap_k, kv_full(stage), sk + p * (kFaTile * 128), h * kD + p * 64, key0);
        for (int kb = 0; kb < 2; ++kb)            // V as the MN-major B operand: [64 keys x 64 channels] boxes, channel chunks 8 KB apart
          for (int nc = 0; nc < kPanels; ++nc)
            tma_load_2d(&tmap_v, kv_full(stage), sv + (kb * kPanels + nc) * 8192, h * kD + nc * 64, key0 + kb * 64);
      }
    }
  } else if (warp == 1) {
    // ======================================================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc(1, 1, 1, false, false, kFaTile, kFaTile);
      const uint32_t idesc_pv = umma_idesc(1, 1, 1, false, true, kFaTile, kD);
      constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);        // K-major, 128-byte swizzle
      constexpr uint64_t kDescMN = umma_desc_hi_lo(8192, 1024);     // MN-major: 64-channel chunks 8 KB apart
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {
        const int stage = j & 1;
        mbar_wait(kv_full(stage), (uint32_t)(j >> 1) & 1u);
        mbar_wait(s_free, ((uint32_t)j & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t sk = smem_kv + stage * S::kStageBytes;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t off = (k / 4) * (kFaTile * 128) + (k % 4) * 32;
          umma_f16<1>(tmem_s, umma_desc(smem_q + off, kDescK), umma_desc(sk + off, kDescK), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit<1>(s_full);
      };
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);          // next scores while the softmax of tile j runs
        const int stage = j & 1;
        mbar_wait(p_full, (uint32_t)j & 1u);
        mbar_wait(pv_free, ((uint32_t)j & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t sv = smem_kv + stage * S::kStageBytes + S::kKBytes;
#pragma unroll
        for (int k = 0; k < kFaTile / 16; ++k) {
          const uint32_t a_off = (k / 4) * (kFaTile * 128) + (k % 4) * 32;                 // P: two 64-key panels
          const uint32_t b_off = (k / 4) * (kPanels * 8192) + (k % 4) * 2048;              // V: 64-key blocks, 16 keys = 2 KB
          umma_f16<1>(tmem_pv, umma_desc(smem_p + a_off, kDescK), umma_desc(sv + b_off, kDescMN), idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit<1>(pv_full);
        umma_commit<1>(kv_empty(stage));
      }
    }
  } else {
    // ======================================================================================= softmax / epilogue
    // Two threads per query row: warps 2-5 own score columns [0, 64) and output channels [0, D/2), warps 6-9 the other halves (both
    // groups map onto the same four TMEM lane quarters).  The row maximum is exchanged through shared memory once per KV tile.
    __shared__ float s_xchg[2][2][kFaTile];
    const uint32_t q = warp & 3u;
    const int half = (int)((warp - 2u) >> 2);
    const int row = (int)(q * 32u + lane);               // query row inside the tile == TMEM lane
    const int row_g = q0 + row;
    const uint32_t lane_addr = (q * 32u) << 16;
    constexpr int kDH = kD / 2;
    float o[kDH];
#pragma unroll
    for (int i = 0; i < kDH; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    uint8_t* p_row = smem_gen + (smem_p - smem_base) + half * (kFaTile * 128) + (row / 8) * 1024 + (row % 8) * 128;
    for (int j = 0; j < n_kv; ++j) {
      const int col0 = j * kFaTile + half * 64;
      mbar_wait(s_full, (uint32_t)j & 1u);
      tcgen05_fence_after();
      const bool need_mask = (kCausal && col0 + 63 > row_g + (Sk - Sq)) || (col0 + 64 > Sk);
      // one pass over this thread's 64 scores: both 32-column loads are issued before the single wait
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(tmem_s + lane_addr + half * 64, r0);
      tmem_ld_32x32b_x32(tmem_s + lane_addr + half * 64 + 32, r1);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);            // S drained into registers: the next QK^T may overwrite it
      float sc[64];
      float m_part = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float v = __uint_as_float(i < 32 ? r0[i] : r1[i - 32]) * scale_log2;
        if (need_mask) {
          const int cg = col0 + i;
          if (cg >= Sk || (kCausal && cg > row_g + (Sk - Sq))) v = -INFINITY;
        }
        sc[i] = v;
        m_part = fmaxf(m_part, v);
      }
      s_xchg[j & 1][half][row] = m_part;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float m_new = fmaxf(m, fmaxf(m_part, s_xchg[j & 1][half ^ 1][row]));
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;        // fully masked row so far: keep everything at zero
      const float corr = exp2f(m - m_use);
      float l_tile = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {                  // eight 16-byte chunks (8 keys each) of this half's 64-key panel
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { pv[i] = exp2f(sc[g * 8 + i] - m_use); l_tile += pv[i]; }
        const int chunk = g ^ (row % 8);
        uint4 v;
        v.x = pack_bf16x2(pv[0], pv[1]); v.y = pack_bf16x2(pv[2], pv[3]); v.z = pack_bf16x2(pv[4], pv[5]); v.w = pack_bf16x2(pv[6], pv[7]);
        *reinterpret_cast<uint4*>(p_row + chunk * 16) = v;
      }
      fence_proxy_async_smem();                    // generic-proxy stores of P -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      l = l * corr + l_tile;
      m = m_new;
      // accumulate this thread's half of O = O * corr + P V
      mbar_wait(pv_full, (uint32_t)j & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < kDH; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_pv + lane_addr + half * kDH + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = o[c + i] * corr + __uint_as_float(r[i]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pv_free);
    }
    // row sum = sum of the two column halves
    s_xchg[n_kv & 1][half][row] = l;
    asm volatile("bar.sync 2, 256;" ::: "memory");
    const float l_all = l + s_xchg[n_kv & 1][half ^ 1][row];
    if (row_g < Sq) {
      const float inv = l_all > 0.f ? 1.f / l_all : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)(b * Sq + row_g) * H + h) * kD + half * kDH);
#pragma unroll
      for (int c = 0; c < kDH; c += 8) {
        uint4 v;
        v.x = pack_bf16x2(o[c + 0] * inv, o[c + 1] * inv); v.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        v.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv); v.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        dst[c / 8] = v;
      }
      if (lse != nullptr && half == 0) lse[((size_t)b * H + h) * Sq + row_g] = (l_all > 0.f) ? (m + log2f(l_all)) * 0.6931471805599453f : -INFINITY;
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
  const dim3 grid((Sq + kFaTile - 1) / kFaTile, H, B);
  kern<<<grid, kFaThreads, S::kTotal, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, Sq, Sk, H, scale * 1.4426950408889634f);
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

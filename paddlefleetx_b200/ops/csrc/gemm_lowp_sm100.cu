// 8-bit tcgen05 GEMMs for sm_100a: int8 x int8 -> int32 (kind::i8, the W8A8 inference GEMM behind the QAT /
// SmoothQuant export) and fp8 e4m3 x e4m3 -> fp32 (kind::f8f6f4, the scaled TP GEMM).
//
//   D[M,N] (bf16) = (A_q[M,K] * B_q[N,K]^T) * row_scale[M] * col_scale[N] (+ bias[N])
//
// Same pipeline as gemm_sm100.cu (TMA producer warp, single-thread MMA issuer, TMEM double-buffered accumulators,
// epilogue warps with swizzled smem staging + TMA store, persistent grouped raster, optional CTA pairs) with the
// 8-bit specifics: 128 elements per 128-byte swizzle row (BLOCK_K = 128), UMMA_K = 32, integer accumulators are
// converted with I2F in the epilogue and the per-token / per-channel dequantisation scales are applied there, so
// the int32 / fp32 product never touches HBM.  Reference: L23 in SURVEY §2.6 (Paddle-Inference / TensorRT int8
// passes — no in-tree kernel in the reference).
//
// kKind 3 = MX block-scaled fp8 (kind::mxf8f6f4.block_scale): every 32 K-elements of every row of A and B carry their own E8M0 scale
// (2^(e - 127)), applied by the tensor core itself.  The quantiser (quant_kernels.cu) writes the scale bytes in the layout the
// smem -> TMEM copy wants — one 512-byte atom per (128 rows, 128 K): byte (r % 32) * 16 + (r / 32) * 4 + k32 — so a stage needs one
// 512-byte bulk copy per operand, one tcgen05.cp (32 x 128 b, broadcast to the four lane quarters) into 4 TMEM columns, and the four
// UMMA K-steps of the stage select their scale byte through the instruction descriptor's a_sf_id / b_sf_id.  1-CTA 128 x 128 tiles.
#include "pfx_ptx.cuh"
#include "pfx_gemm.h"
#include <cudaTypedefs.h>

namespace pfx {
namespace lowp {

constexpr int kBlockM = 128;
constexpr int kBlockK = 128;      // 8-bit elements: 128 B = one swizzle span
constexpr int kUmmaK = 32;
constexpr int kNumThreads = 256;
constexpr int kGroupM = 16;
constexpr int kStoreCols = 64;

template <int kCG, int kBlockN, bool kMx = false>
struct Smem {
  static constexpr int kLoadN = kBlockN / kCG;
  static constexpr int kABytes = kBlockM * kBlockK;
  static constexpr int kBBytes = kLoadN * kBlockK;
  static constexpr int kSfBytes = 512;                      // one scale-factor atom (128 rows x 4 K-blocks of 32) per operand, MX only
  static constexpr int kStageBytes = kABytes + kBBytes + (kMx ? 2 * kSfBytes : 0);
  static constexpr int kEpiBytes = kBlockM * kStoreCols * 2;
  static constexpr int kBudget = 227 * 1024 - 1024 - 1024 - 2 * kEpiBytes;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kTotal = 1024 + kStages * kStageBytes + 2 * kEpiBytes + 1024;
};

__device__ __forceinline__ void tile_of(int tile, int num_m, int num_n, int& m_blk, int& n_blk) {
  const int per_group = kGroupM * num_n;
  const int group = tile / per_group;
  const int first = group * kGroupM;
  const int gm = min(kGroupM, num_m - first);
  const int in = tile - group * per_group;
  m_blk = first + in % gm;
  n_blk = in / gm;
}

// kKind: 1 = int8 (s32 accumulate), 2 = fp8 e4m3 (f32 accumulate)
template <int kCG, int kBlockN, int kKind>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_lowp_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_d, const float* __restrict__ row_scale, const float* __restrict__ col_scale,
                 const __nv_bfloat16* __restrict__ bias, int M, int N, int K, const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb) {
  constexpr bool kMx = kKind == 3;
  static_assert(!kMx || (kCG == 1 && kBlockN == 128), "MX path: 1-CTA 128 x 128 tiles");
  using S = Smem<kCG, kBlockN, kMx>;
  constexpr int kStages = S::kStages;
  constexpr int kLoadN = S::kLoadN;
  constexpr int kUmmaM = kBlockM * kCG;
  constexpr int kTmemCols = kMx ? 512 : 2 * kBlockN;       // MX: accumulators [0, 256), scale factors from column 256
  constexpr uint32_t kSfCol = 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_epi = smem_base + kStages * S::kStageBytes;
  const uint32_t smem_bar = smem_epi + 2 * S::kEpiBytes;
  auto full_bar = [&](int s) { return smem_bar + 8u * s; };
  auto empty_bar = [&](int s) { return smem_bar + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return smem_bar + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return smem_bar + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = smem_bar + 8u * (2 * kStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t warp = warp_id(), lane = lane_id();
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const int num_m = (M + kUmmaM - 1) / kUmmaM, num_n = (N + kBlockN - 1) / kBlockN;
  const int num_tiles = num_m * num_n, num_kb = (K + kBlockK - 1) / kBlockK;
  const int num_clusters = gridDim.x / kCG, cluster_id = blockIdx.x / kCG;

  if (warp == 0 && elect_one()) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_d); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4 * kCG); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc<kCG>(tmem_slot, kTmemCols); tmem_relinquish<kCG>(); }
  tcgen05_fence_before();
  if (kCG == 2) cluster_sync(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int m_blk, n_blk; tile_of(tile, num_m, num_n, m_blk, n_blk);
        const int m_idx = m_blk * kUmmaM + (int)cta_rank * kBlockM, n_idx = n_blk * kBlockN + (int)cta_rank * kLoadN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * S::kStageBytes, sb = sa + S::kABytes, fb = full_bar(stage);
          if (kCG == 1 || cta_rank == 0) mbar_arrive_expect_tx(fb, S::kStageBytes * kCG);
          if (kCG == 2) { tma_load_2d_2sm(&tmap_a, fb, sa, kb * kBlockK, m_idx); tma_load_2d_2sm(&tmap_b, fb, sb, kb * kBlockK, n_idx); }
          else { tma_load_2d(&tmap_a, fb, sa, kb * kBlockK, m_idx); tma_load_2d(&tmap_b, fb, sb, kb * kBlockK, n_idx); }
          if constexpr (kMx) {      // the stage's two scale-factor atoms (already in copy order in global memory)
            bulk_load_1d(sb + S::kBBytes, sfa + ((size_t)m_blk * num_kb + kb) * S::kSfBytes, S::kSfBytes, fb);
            bulk_load_1d(sb + S::kBBytes + S::kSfBytes, sfb + ((size_t)n_blk * num_kb + kb) * S::kSfBytes, S::kSfBytes, fb);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0 && elect_one()) {
      // int8: c = s32 (2), a/b = signed 8 bit (1).  fp8: c = f32 (1), a/b = e4m3 (0).
      const uint32_t idesc = kKind == 1 ? umma_idesc(2, 1, 1, false, false, kUmmaM, kBlockN) : umma_idesc(1, 0, 0, false, false, kUmmaM, kBlockN);
      constexpr uint64_t kDesc = umma_desc_hi_lo(16, 1024);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * S::kStageBytes, sb = sa + S::kABytes;
          uint32_t tsf = 0;
          if constexpr (kMx) {      // scale factors of this stage: smem -> 4 + 4 TMEM columns (two alternating sets); cp and mma run in issue order
            tsf = tmem_base + kSfCol + 8u * (uint32_t)(stage & 1);
            tmem_cp_32x128b_warpx4(tsf, umma_desc_noswizzle(sb + S::kBBytes, 128));
            tmem_cp_32x128b_warpx4(tsf + 4, umma_desc_noswizzle(sb + S::kBBytes + S::kSfBytes, 128));
          }
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = umma_desc(sa + k * kUmmaK, kDesc), db = umma_desc(sb + k * kUmmaK, kDesc);   // 32 B per UMMA_K step
            if constexpr (kKind == 1) umma_i8<kCG>(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else if constexpr (kKind == 2) umma_f8<kCG>(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_mxf8(d_tmem, da, db, umma_idesc_mxf8(kUmmaM, kBlockN, k, k), tsf, tsf + 4, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit<kCG>(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit<kCG>(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const uint32_t q = warp & 3u, row_in_cta = q * 32 + lane;
    const bool is_store_thread = (warp == 4) && (lane == 0);
    const uint32_t tempty_leader = mapa(tempty_bar(0), 0);
    int acc = 0; uint32_t acc_phase = 0, store_iter = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int m_blk, n_blk; tile_of(tile, num_m, num_n, m_blk, n_blk);
      const int row0 = m_blk * kUmmaM + (int)cta_rank * kBlockM, col_tile = n_blk * kBlockN;
      const int grow = row0 + (int)row_in_cta;
      const float rs = (row_scale && grow < M) ? row_scale[grow] : 1.f;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c = 0; c < kBlockN / kStoreCols; ++c) {
        const int col0 = col_tile + c * kStoreCols;
        uint32_t r[2][32];
        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * kBlockN + c * kStoreCols;
        tmem_ld_32x32b_x32(taddr, r[0]);
        tmem_ld_32x32b_x32(taddr + 32, r[1]);
        tmem_ld_wait();
        if (c == kBlockN / kStoreCols - 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * acc);
        }
        float v[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const uint32_t raw = r[i >> 5][i & 31];
          v[i] = (kKind == 1 ? __int2float_rn((int)raw) : __uint_as_float(raw)) * rs;
        }
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          if (col0 + i < N) {
            if (col_scale) {
              const float4 cs = __ldg(reinterpret_cast<const float4*>(col_scale + col0 + i));
              v[i] *= cs.x; v[i + 1] *= cs.y; v[i + 2] *= cs.z; v[i + 3] *= cs.w;
            }
            if (bias) {
              const uint2 bv = __ldg(reinterpret_cast<const uint2*>(bias + col0 + i));
              const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&bv);
              const float2 f0 = __bfloat1622float2(b2[0]), f1 = __bfloat1622float2(b2[1]);
              v[i] += f0.x; v[i + 1] += f0.y; v[i + 2] += f1.x; v[i + 3] += f1.y;
            }
          }
        }
        const uint32_t buf = store_iter & 1u;
        if (store_iter >= 2) {
          if (is_store_thread) tma_store_wait_read<1>();
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        const uint32_t sbase = smem_epi + buf * S::kEpiBytes + row_in_cta * 128u;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t dst = sbase + (((uint32_t)ch ^ (row_in_cta & 7u)) << 4);
          const uint32_t p0 = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]), p1 = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
          const uint32_t p2 = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]), p3 = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (is_store_thread) {
          if (col0 < N && row0 < M) tma_store_2d(&tmap_d, smem_epi + buf * S::kEpiBytes, col0, row0);
          tma_store_commit();
        }
        ++store_iter;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (is_store_thread) tma_store_wait<0>();
  }
  tcgen05_fence_before();
  if (kCG == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCG>(tmem_base, kTmemCols);
}

template <int kCG, int kBlockN, int kKind>
static cudaError_t launch(const LowpGemmArgs& g, cudaStream_t stream) {
  using S = Smem<kCG, kBlockN, kKind == 3>;
  if (kKind == 3 && (g.sfa == nullptr || g.sfb == nullptr || g.K % kBlockK != 0)) return cudaErrorInvalidValue;
  CUtensorMap ta, tb, td;
  bool ok = make_tmap_2d(&ta, g.a, 1, 2, g.K, g.M, (uint64_t)g.lda, kBlockK, kBlockM);
  ok &= make_tmap_2d(&tb, g.b, 1, 2, g.K, g.N, (uint64_t)g.ldb, kBlockK, S::kLoadN);
  ok &= make_tmap_2d(&td, g.d, 2, 1, g.N, g.M, (uint64_t)g.ldd * 2, kStoreCols, kBlockM);
  if (!ok) return cudaErrorInvalidValue;
  auto kern = gemm_lowp_kernel<kCG, kBlockN, kKind>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = ((g.M + kBlockM * kCG - 1) / (kBlockM * kCG)) * ((g.N + kBlockN - 1) / kBlockN);
  int clusters = g.num_sms / kCG;
  if (clusters > tiles) clusters = tiles;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * kCG); cfg.blockDim = dim3(kNumThreads); cfg.dynamicSmemBytes = S::kTotal; cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCG; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, td, g.row_scale, g.col_scale, reinterpret_cast<const __nv_bfloat16*>(g.bias), g.M, g.N, g.K,
                            reinterpret_cast<const uint8_t*>(g.sfa), reinterpret_cast<const uint8_t*>(g.sfb));
}

}  // namespace lowp

cudaError_t gemm_lowp_tcgen05(const LowpGemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return cudaSuccess;
  const long tiles_big = (long)((g.M + 255) / 256) * ((g.N + 255) / 256);
  const bool pair = g.config == 2 || (g.config == 0 && tiles_big >= g.num_sms / 2);
  if (g.kind == 3) return lowp::launch<1, 128, 3>(g, stream);
  if (g.kind == 1) return pair ? lowp::launch<2, 256, 1>(g, stream) : lowp::launch<1, 128, 1>(g, stream);
  return pair ? lowp::launch<2, 256, 2>(g, stream) : lowp::launch<1, 128, 2>(g, stream);
}

}  // namespace pfx

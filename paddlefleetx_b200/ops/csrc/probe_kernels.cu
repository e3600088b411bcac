// Hardware probes: tiny kernels that pin down an instruction's semantics on the real part before a production kernel depends on it.
//
//   probe_tmem_a   D[128, N] = A[128, K] . B[N, K]^T with the A operand of tcgen05.mma read from TENSOR MEMORY (written there by
//                  tcgen05.st: lane = row, 32-bit column c = elements (2c, 2c + 1) of the row) instead of shared memory.  This is the layout the
//                  attention kernels want for P / dS (produced in registers by the softmax threads, one row per thread): if it holds, the
//                  probabilities never touch shared memory.
#include "pfx_ptx.cuh"
#include "pfx_gemm.h"
#include "pfx_kernels.h"

namespace pfx {

namespace {


template <int kN, int kK>
__global__ void __launch_bounds__(128, 1) probe_tmem_a_kernel(const __nv_bfloat16* __restrict__ a, const __grid_constant__ CUtensorMap tmap_b,
                                                              float* __restrict__ d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kPanels = kK / 64;
  const uint32_t s_b = smem_base, s_bar = smem_base + kPanels * kN * 128;
  const uint32_t full = s_bar, done = s_bar + 8, slot = s_bar + 16;
  const uint32_t warp = warp_id(), lane = lane_id();
  if (warp == 0) {
    if (elect_one()) { mbar_init(full, 1); mbar_init(done, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<1>(slot, 256);
    tmem_relinquish<1>();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_gen + (slot - smem_base));
  const uint32_t t_d = tmem, t_a = tmem + 128;                // D: columns [0, kN), A: kK / 2 packed columns from 128
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(full, kPanels * kN * 128);
    for (int p = 0; p < kPanels; ++p) tma_load_2d(&tmap_b, full, s_b + p * kN * 128, p * 64, 0);
  }
  // every thread packs its row of A (kK bf16 = kK / 2 words) into tensor memory
  const int row = (int)(warp * 32u + lane);
  const uint32_t lane_addr = (warp * 32u) << 16;
  for (int c0 = 0; c0 < kK / 2; c0 += 32) {
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = reinterpret_cast<const uint32_t*>(a + (size_t)row * kK)[c0 + i];
    tmem_st_32x32b_x32(t_a + lane_addr + c0, r);
  }
  tmem_st_wait();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 0 && elect_one()) {
    mbar_wait(full, 0);
    tcgen05_fence_after();
    const uint32_t idesc = umma_idesc(1, 1, 1, false, false, 128, kN);
    constexpr uint64_t kDescK = umma_desc_hi_lo(16, 1024);
#pragma unroll
    for (int kk = 0; kk < kK / 16; ++kk)
      umma_f16_ts(t_d, t_a + kk * 8, umma_desc(s_b + (kk / 4) * (kN * 128) + (kk % 4) * 32, kDescK), idesc, kk != 0 ? 1u : 0u);
    umma_commit<1>(done);
  }
  mbar_wait(done, 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_d + lane_addr + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) d[(size_t)row * kN + c0 + i] = __uint_as_float(r[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem, 256);
}

}  // namespace

// a [128, 128] bf16, b [64, 128] bf16 (both row-major, K contiguous) -> d [128, 64] fp32
cudaError_t probe_tmem_a(const void* a, const void* b, float* d, cudaStream_t st) {
  constexpr int kN = 64, kK = 128;
  CUtensorMap tb;
  if (!make_tmap_2d(&tb, b, 2, 1, kK, kN, (uint64_t)kK * 2, 64, kN)) return cudaErrorInvalidValue;
  auto kern = probe_tmem_a_kernel<kN, kK>;
  const int smem = 1024 + (kK / 64) * kN * 128 + 64;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  kern<<<1, 128, smem, st>>>((const __nv_bfloat16*)a, tb, d);
  return cudaGetLastError();
}

}  // namespace pfx

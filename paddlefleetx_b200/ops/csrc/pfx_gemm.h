// Host-side C++ interface of the tcgen05 GEMM family (no torch headers: keeps .cu compiles fast).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pfx {

enum Epilogue : int { EPI_NONE = 0, EPI_BIAS = 1, EPI_BIAS_GELU = 2, EPI_GELU = 3 };

struct GemmArgs {
  const void* a;     // K-major: [M, K] row-major (lda = row stride, elements); MN-major: [K, M]
  const void* b;     // K-major: [N, K] row-major;                              MN-major: [K, N]
  void* d;           // [M, N] row-major, ldd = row stride (elements)
  const void* bias;  // [N] bf16 or nullptr
  int M, N, K;
  int lda, ldb, ldd;
  bool a_kmajor, b_kmajor;
  int out_mode;      // 0 = same dtype as inputs (TMA store), 1 = fp32 store, 2 = fp32 accumulate
  int epilogue;      // Epilogue
  int ab_format;     // 0 = fp16, 1 = bf16
  int num_sms;
  int config;        // 0 = auto
};

cudaError_t gemm_tcgen05(const GemmArgs& args, cudaStream_t stream);

bool make_tmap_2d(CUtensorMap* map, const void* ptr, int elem_bytes, int dtype_code, uint64_t inner, uint64_t outer,
                  uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

}  // namespace pfx

// Device-side helpers shared by the attention forward and backward kernels.
#pragma once
#include <cstdint>

namespace pfx {

// Counter-hash Bernoulli stream for attention dropout.  The forward kernel walks keys for a fixed query, the backward kernel
// walks queries for a fixed key, so the generator must be addressable per ELEMENT at a few instructions each (Philox would cost
// ~80 per call): a murmur3-style finaliser over (query, key-pair) with the per-(batch, head) key injected between the two
// multiplies; the two 16-bit halves of the result serve the even and the odd key of the pair.
__device__ __forceinline__ uint32_t attn_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t attn_rng_key(uint64_t seed, uint32_t bh) {
  return attn_mix((uint32_t)seed + 0x9E3779B9u * (bh + 1u)) ^ (uint32_t)(seed >> 32);
}
__device__ __forceinline__ uint32_t attn_rng_pair(uint32_t key, uint32_t q, uint32_t kpair, uint32_t pairs_per_row) {
  uint32_t x = q * pairs_per_row + kpair;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= key; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool attn_keep(uint32_t pair_bits, uint32_t k, uint32_t thresh16) {
  return ((k & 1u) ? (pair_bits >> 16) : (pair_bits & 0xFFFFu)) >= thresh16;
}

}  // namespace pfx

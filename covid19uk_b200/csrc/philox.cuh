// Counter-based Philox4x32-10 generator (Salmon et al., SC'11) for the device-side samplers.
// stream = (seed, global chain id); position = (sweep, purpose, index) => results do not depend on how
// chains are partitioned over GPUs.  The reference never seeds its RNG (inference.py:68,134,205),
// so no stream can be matched; only the proposal DISTRIBUTIONS are restated.
#pragma once
#include <stdint.h>

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
  c[0] = h1 ^ c[1] ^ k0;
  c[1] = l1;
  c[2] = h0 ^ c[3] ^ k1;
  c[3] = l0;
}

__device__ __forceinline__ void seir_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform in (0,1) with 53 random bits
__device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo) {
  const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// uniform integer in [0, n) from 64 random bits (multiply-high; bias < n / 2^64)
__device__ __forceinline__ uint32_t rand_below(uint32_t hi, uint32_t lo, uint32_t n) {
  const uint64_t bits = ((uint64_t)hi << 32) | lo;
  return (uint32_t)__umul64hi(bits, (uint64_t)n);
}

// Counter-based random numbers on the device: Philox4x32-10 (Salmon et al., SC'11; the generator behind torch's CUDA
// Philox streams), used for
//   * Normal.sample() noise of the rollout policy inside head_kernel (reference U/cleanrl/ppo.py:112-114),
//   * the minibatch permutation of every epoch (ppo.py:295 `torch.randperm`): a keyed Feistel bijection + cycle walking,
//   * the Bernoulli draws of the env-side MDP terms (U/mdp/commands.py:75,88, U/mdp/events.py:59-96) and the optional
//     stochastic termination mask.
// Every draw is a pure function of (seed, stream id, counter), so a CPU restatement (oracle/philox_oracle.py)
// reproduces the integer outputs bit for bit.  `rng_state` in the C ABI is {seed, offset}: two uint64 on the device.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace catb200 {

enum PhiloxStream : uint32_t {
  kStreamActionNoise = 0,
  kStreamPermutation = 1,
  kStreamBernoulli = 2,
  kStreamUniform = 3,
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10 on raw counter / key words (Random123 word order; pinned by its known-answer vectors)
__host__ __device__ __forceinline__ void philox4x32_10_raw(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// 4 x 32 random bits for counter (ctr, stream) under key `seed`
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t seed, uint32_t stream, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0u};
  philox4x32_10_raw(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform in (0, 1) with 24 random bits: (w >> 8) * 2^-24 + 2^-25
__host__ __device__ __forceinline__ float u24(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f; }

// standard normal for element `idx` of the draw that starts at `offset` (Box-Muller on words 0 and 1)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint32_t w[4];
  philox4x32_10(seed, kStreamActionNoise, offset + idx, w);
  const float r = sqrtf(-2.0f * logf(u24(w[0])));
  return r * cosf(6.283185307179586f * u24(w[1]));
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t stream, uint64_t ctr) {
  uint32_t w[4];
  philox4x32_10(seed, stream, ctr, w);
  return u24(w[0]);
}

// Keyed bijection of [0, 2^bits): unbalanced Feistel network, 6 rounds, round function = one multiply-xorshift of the
// half block with a Philox-derived round key.  With cycle walking (re-apply until the image is < n) it becomes a
// pseudo-random permutation of [0, n) that needs no sort and no global memory.
__host__ __device__ __forceinline__ uint32_t feistel_bijection(uint32_t x, int bits, const uint32_t (&rk)[8]) {
  int wl = bits / 2, wr = bits - wl;  // widths of the left / right halves (swap every round)
  uint32_t L = x >> wr, R = x & ((1u << wr) - 1u);
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    uint32_t f = (R ^ rk[r]) * 0x9E3779B1u;
    f ^= f >> 15;
    f *= 0x85EBCA77u;
    f ^= f >> 13;
    const uint32_t nl = R;                            // wr bits
    const uint32_t nr = (L ^ f) & ((1u << wl) - 1u);  // wl bits
    L = nl; R = nr;
    const int t = wl; wl = wr; wr = t;
  }
  return (L << wr) | R;
}

}  // namespace catb200

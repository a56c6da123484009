// Optimizer pieces shared by optim.cu (global-norm clip) and mlp.cu (Adam fused with the operand-copy refresh).
#pragma once

#include "common.cuh"

namespace catb200 {

struct OptScratch {  // 64 bytes of caller-provided zero-initialised device memory
  unsigned int ticket;
  float clip_coef, total_norm, step_size_scale, bc2_sqrt;
  float pad[3];
  double sumsq;
  double pad2[3];
};

// torch.optim.Adam (no weight decay, not amsgrad), one element; zeroes the gradient for the next minibatch
__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float clip, float step_size,
                                         float bc2_sqrt, float beta1, float beta2, float eps) {
  const float gs = g * clip;
  g = 0.0f;
  m = m + (gs - m) * (1.0f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.0f - beta2) * gs * gs;     // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;  // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - step_size * (m / denom);                // addcdiv_(exp_avg, denom, value = -step_size)
}

struct AdamState {
  float* params; float* grads; float* m; float* v;
  const float* lr; const OptScratch* sc;
  float beta1, beta2, eps, grad_scale;
};

// ---- peer-visible gradient block of the multi-GPU exchange (peer.cu, mlp.cu): [flags: kFlagWords x u32][arena 0][arena 1]
// flag row of rank r (64 words): [0, 8) "arena ready" epoch written by rank q into word q, [8, 16) "slice + partial norm
// delivered" epoch of the reduce-scatter path, bytes [64, 128) one double per rank: its partial squared norm.  Behind the
// two arenas sits a third region of n_pad floats: the summed gradient the reduce-scatter path assembles from the peers' slices.
constexpr int kPeerMax = 8;
constexpr int kFlagWords = 64;
constexpr int kFlag2Word = 8;
constexpr int kNormByte = 64;
__host__ __device__ inline size_t peer_n_pad(long long n_params) { return ((size_t)n_params + 63) / 64 * 64; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// relaxed system-scope flag accesses: one __threadfence_system() before a batch of flag stores / after a successful poll
// gives the release / acquire ordering once, instead of once per peer and once per poll iteration
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// mlp.cu: Adam over the whole flat parameter vector AND the operand-precision compute copies (W, W^T) of the six
// hidden matrices in one launch -- the tiled cast kernel with the update applied to every element on its way through.
int launch_adam_cast(const catb200_mlp_dims_t* dims, const AdamState& a, void* wc, cudaStream_t st);

}  // namespace catb200

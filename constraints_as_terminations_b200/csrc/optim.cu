// Global-norm gradient clipping + Adam on the flat parameter vector (sm_100a).
//
// Replaces nn.utils.clip_grad_norm_(agent.parameters(), max_norm) + optimizer.step() + optimizer.zero_grad()
// of the reference (U/cleanrl/ppo.py:351-354; torch.optim.Adam, eps=1e-5 set at ppo.py:168) with two
// launches over the 377k-element flat buffers:
//   grad_norm_kernel : sum of squares (double) -> total norm, clip coefficient, Adam bias corrections
//   adam_kernel      : scaled gradient -> moments -> parameter update (float4), zeroes the gradient for the
//                      next minibatch; then the tiled cast kernel refreshes the operand copies (W and W^T).
#include "common.cuh"
#include "mma.cuh"

namespace catb200 {

struct OptScratch {  // 64 bytes of caller-provided zero-initialised device memory
  unsigned int ticket;
  float clip_coef, total_norm, step_size_scale, bc2_sqrt;
  float pad[3];
  double sumsq;
  double pad2[3];
};

__global__ void __launch_bounds__(256)
grad_norm_kernel(const float* __restrict__ grads, long long n, float grad_scale, float max_norm, float beta1,
                 float beta2, int* __restrict__ step, float* __restrict__ grad_norm_out, OptScratch* __restrict__ sc) {
  pdl_launch_dependents();
  pdl_wait();
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const double g = (double)(grads[e] * grad_scale);
    s += g * g;
  }
  s = warp_sum(s);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(&sc->sumsq, t);
  }
  if (last_block_ticket(&sc->ticket, gridDim.x)) {
    if (threadIdx.x == 0) {
      const double tot = __longlong_as_double(atomicExch((unsigned long long*)&sc->sumsq, 0ull));
      const float norm = (float)sqrt(tot);
      // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
      sc->clip_coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
      sc->total_norm = norm;
      const int t = *step + 1;
      *step = t;
      const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
      sc->step_size_scale = (float)(1.0 / bc1);
      sc->bc2_sqrt = (float)sqrt(bc2);
      if (grad_norm_out) *grad_norm_out = norm;
    }
  }
}

struct AdamArgs {
  float* params; float* grads; float* m; float* v;
  const float* lr; const OptScratch* sc;
  long long n;
  float beta1, beta2, eps, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, float clip, float step_size,
                                         float bc2_sqrt, float beta1, float beta2, float eps) {
  const float gs = g * clip;
  g = 0.0f;
  m = m + (gs - m) * (1.0f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (1.0f - beta2) * gs * gs;     // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;  // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - step_size * (m / denom);                // addcdiv_(exp_avg, denom, value = -step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const float clip = a.sc->clip_coef * a.grad_scale;
  const float step_size = __ldg(a.lr) * a.sc->step_size_scale;
  const float bc2_sqrt = a.sc->bc2_sqrt;
  const long long n4 = a.n / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 p = reinterpret_cast<float4*>(a.params)[i], g = reinterpret_cast<float4*>(a.grads)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i], v = reinterpret_cast<float4*>(a.v)[i];
    adam_one(p.x, g.x, m.x, v.x, clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
    adam_one(p.y, g.y, m.y, v.y, clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
    adam_one(p.z, g.z, m.z, v.z, clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
    adam_one(p.w, g.w, m.w, v.w, clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
    reinterpret_cast<float4*>(a.params)[i] = p;
    reinterpret_cast<float4*>(a.grads)[i] = g;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  } else if (i < n4 + (a.n - n4 * 4)) {  // scalar tail
    const long long e = n4 * 4 + (i - n4);
    adam_one(a.params[e], a.grads[e], a.m[e], a.v[e], clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
  }
}

}  // namespace catb200

using namespace catb200;

extern "C" {

int catb200_adam_step(const catb200_mlp_dims_t* dims, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                      void* wc, const float* lr_dev, int32_t* step_dev, float max_grad_norm, float beta1, float beta2,
                      float eps, float grad_scale, float* grad_norm_out, void* opt_ws, void* stream) {
  if (!dims || !params || !grads || !exp_avg || !exp_avg_sq || !wc || !lr_dev || !step_dev || !opt_ws)
    return CATB200_ERR_INVALID_ARGUMENT;
  catb200_mlp_layout_t P;
  int rc = catb200_mlp_layout(dims, &P);
  if (rc != CATB200_OK) return rc;
  cudaStream_t st = as_stream(stream);
  OptScratch* sc = static_cast<OptScratch*>(opt_ws);
  const long long n = P.n_params;
  CATB200_CUDA_TRY(launch_pdl(grad_norm_kernel, dim3(kNumSMs), dim3(256), 0, st, (const float*)grads, n, grad_scale, max_grad_norm, beta1, beta2,
                              (int*)step_dev, grad_norm_out, sc));
  CATB200_LAUNCH_CHECK();
  AdamArgs a = {};
  a.params = params; a.grads = grads; a.m = exp_avg; a.v = exp_avg_sq;
  a.lr = lr_dev; a.sc = sc; a.n = n; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
  const long long threads = n / 4 + 4;
  CATB200_CUDA_TRY(launch_pdl(adam_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, st, a));
  CATB200_LAUNCH_CHECK();
  rc = launch_cast_weights(dims, params, wc, st);  // refresh the operand copies the tensor-core GEMMs read
  if (rc != CATB200_OK) return rc;
  return CATB200_OK;
}

}  // extern "C"

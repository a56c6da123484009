// Global-norm gradient clipping + Adam on the flat parameter vector (sm_100a).
//
// Replaces nn.utils.clip_grad_norm_(agent.parameters(), max_norm) + optimizer.step() + optimizer.zero_grad()
// of the reference (U/cleanrl/ppo.py:351-354; torch.optim.Adam, eps=1e-5 set at ppo.py:168) with two
// launches over the 377k-element flat buffers:
//   grad_norm_kernel : sum of squares (double) -> total norm, clip coefficient, Adam bias corrections
//   adam_kernel      : scaled gradient -> moments -> parameter update, zeroes the gradient for the next
//                      minibatch and refreshes the bf16 compute copies (W and W^T) of the hidden layers.
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

struct CastTarget {
  long long off;  // offset of the fp32 weight in the flat vector
  int rows, cols, cols_pad;
  long long dst, dst_t;  // bf16 offsets (dst_t < 0: no transposed copy)
};
struct AdamArgs {
  float* params; float* grads; float* m; float* v; bf16* w16;
  const float* lr; const OptScratch* sc;
  long long n;
  float beta1, beta2, eps, grad_scale;
  CastTarget cast[6];
};

__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamArgs a) {
  const float clip = a.sc->clip_coef * a.grad_scale;
  const float step_size = __ldg(a.lr) * a.sc->step_size_scale;
  const float bc2_sqrt = a.sc->bc2_sqrt;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a.n; e += (long long)gridDim.x * blockDim.x) {
    const float g = a.grads[e] * clip;
    a.grads[e] = 0.0f;
    const float m = a.m[e] + (g - a.m[e]) * (1.0f - a.beta1);          // exp_avg.lerp_(grad, 1 - beta1)
    const float v = a.v[e] * a.beta2 + (1.0f - a.beta2) * g * g;        // mul_(beta2).addcmul_(g, g, 1 - beta2)
    a.m[e] = m;
    a.v[e] = v;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    const float p = a.params[e] - step_size * (m / denom);
    a.params[e] = p;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const CastTarget& c = a.cast[s];
      const long long local = e - c.off;
      if (local >= 0 && local < (long long)c.rows * c.cols) {
        const int r = (int)(local / c.cols), k = (int)(local - (long long)r * c.cols);
        const bf16 pv = __float2bfloat16(p);
        a.w16[c.dst + (long long)r * c.cols_pad + k] = pv;
        if (c.dst_t >= 0) a.w16[c.dst_t + (long long)k * c.rows + r] = pv;
      }
    }
  }
}

}  // namespace catb200

using namespace catb200;

extern "C" {

int catb200_adam_step(const catb200_mlp_dims_t* dims, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                      void* w16, const float* lr_dev, int32_t* step_dev, float max_grad_norm, float beta1, float beta2,
                      float eps, float grad_scale, float* grad_norm_out, void* opt_ws, void* stream) {
  if (!dims || !params || !grads || !exp_avg || !exp_avg_sq || !w16 || !lr_dev || !step_dev || !opt_ws)
    return CATB200_ERR_INVALID_ARGUMENT;
  catb200_mlp_layout_t P;
  int rc = catb200_mlp_layout(dims, &P);
  if (rc != CATB200_OK) return rc;
  cudaStream_t st = as_stream(stream);
  OptScratch* sc = static_cast<OptScratch*>(opt_ws);
  const long long n = P.n_params;
  grad_norm_kernel<<<kNumSMs, 256, 0, st>>>(grads, n, grad_scale, max_grad_norm, beta1, beta2, step_dev, grad_norm_out, sc);
  CATB200_LAUNCH_CHECK();
  AdamArgs a = {};
  a.params = params; a.grads = grads; a.m = exp_avg; a.v = exp_avg_sq; a.w16 = static_cast<bf16*>(w16);
  a.lr = lr_dev; a.sc = sc; a.n = n; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
  const int in[3] = {dims->obs_dim, dims->h1, dims->h2}, in_pad[3] = {dims->obs_pad, dims->h1, dims->h2};
  const int out[3] = {dims->h1, dims->h2, dims->h3};
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      CastTarget& c = a.cast[z * 3 + l];
      c.off = P.w[z][l]; c.rows = out[l]; c.cols = in[l]; c.cols_pad = in_pad[l];
      c.dst = P.w16[z][l]; c.dst_t = P.wt16[z][l];
    }
  adam_kernel<<<kNumSMs * 4, 256, 0, st>>>(a);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

// Global-norm gradient clipping + Adam on the flat parameter vector (sm_100a).
//
// Replaces nn.utils.clip_grad_norm_(agent.parameters(), max_norm) + optimizer.step() + optimizer.zero_grad()
// of the reference (U/cleanrl/ppo.py:351-354; torch.optim.Adam, eps=1e-5 set at ppo.py:168) with two
// launches over the 377k-element flat buffers:
//   grad_norm_kernel : sum of squares (double) -> total norm, clip coefficient, Adam bias corrections
//   adam_cast_kernel : (mlp.cu) scaled gradient -> moments -> parameter update, zeroes the gradient for the next
//                      minibatch and writes the operand-precision copies (W and W^T) of the hidden matrices on the way
#include "common.cuh"
#include "mma.cuh"
#include "optim.cuh"

namespace catb200 {

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
  // Adam over the whole flat vector and the refresh of the operand copies (W, W^T) the tensor-core GEMMs read, one launch
  AdamState a = {params, grads, exp_avg, exp_avg_sq, lr_dev, sc, beta1, beta2, eps, grad_scale};
  rc = launch_adam_cast(dims, a, wc, st);
  if (rc != CATB200_OK) return rc;
  return CATB200_OK;
}

// The Adam half alone (clip coefficient and bias corrections already in opt_ws: written by catb200_grad_allreduce_norm)
int catb200_adam_apply(const catb200_mlp_dims_t* dims, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                       void* wc, const float* lr_dev, float beta1, float beta2, float eps, float grad_scale, void* opt_ws,
                       void* stream) {
  if (!dims || !params || !grads || !exp_avg || !exp_avg_sq || !wc || !lr_dev || !opt_ws) return CATB200_ERR_INVALID_ARGUMENT;
  AdamState a = {params, grads, exp_avg, exp_avg_sq, lr_dev, static_cast<OptScratch*>(opt_ws), beta1, beta2, eps, grad_scale};
  return launch_adam_cast(dims, a, wc, as_stream(stream));
}

}  // extern "C"

// Running mean / variance (RunningMeanStd) and rollout append on sm_100a.
//
// Replaces the ~8 eager kernels of RunningMeanStd.forward (U/cleanrl/ppo.py:21-62) with two launches:
//   rms_moments_kernel  : column sums / sums of squares of x[rows, dim] accumulated in double
//                         (per-thread -> per-CTA shared -> one double atomicAdd per column per CTA);
//                         the last CTA turns them into batch mean / biased variance and applies the
//                         Chan merge of update_mean_var_count_from_moments (ppo.py:48-62) in fp32 with
//                         the reference's operation order.
//   rms_normalize_kernel: out = (x - mean) / sqrt(var + eps), elementwise, IEEE division / sqrt.
// Thread mapping: a CTA uses floor(256/dim)*dim threads so that every thread keeps a fixed column while
// the CTA walks whole rows; consecutive threads touch consecutive addresses (coalesced), no modulo in
// the loop.  HBM-bound streaming work.
#include "common.cuh"
#include "mma.cuh"

namespace catb200 {

constexpr int kRmsThreads = 256;

struct RmsWorkspace {
  unsigned int* ticket;
  double* sums;  // [2*dim]: sum, then sum of squares
};

__host__ __device__ inline RmsWorkspace rms_carve(void* base, int dim) {
  RmsWorkspace w;
  w.ticket = reinterpret_cast<unsigned int*>(base);
  w.sums = reinterpret_cast<double*>(static_cast<char*>(base) + 256);
  (void)dim;
  return w;
}

// Chan et al. merge, fp32, same order of operations as ppo.py:51-62.
__device__ __forceinline__ void chan_merge(float& mean, float& var, float& count, float bmean, float bvar, float n) {
  const float delta = __fsub_rn(bmean, mean);
  const float tot = __fadd_rn(count, n);
  const float new_mean = __fadd_rn(mean, __fdiv_rn(__fmul_rn(delta, n), tot));
  const float m_a = __fmul_rn(var, count);
  const float m_b = __fmul_rn(bvar, n);
  const float cross = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), count), n), tot);
  const float m2 = __fadd_rn(__fadd_rn(m_a, m_b), cross);
  mean = new_mean;
  var = __fdiv_rn(m2, tot);
  count = tot;
}

__global__ void __launch_bounds__(kRmsThreads)
rms_moments_kernel(const float* __restrict__ x, long long rows, int dim, float* __restrict__ mean,
                   float* __restrict__ var, float* __restrict__ count, RmsWorkspace ws) {
  const int rows_per_pass = kRmsThreads / dim;
  const int active = rows_per_pass * dim;
  const int col = threadIdx.x % dim;
  const int slot = threadIdx.x / dim;
  double s = 0.0, q = 0.0;
  if (threadIdx.x < active) {
    const long long step = (long long)gridDim.x * rows_per_pass;
    long long r = (long long)blockIdx.x * rows_per_pass + slot;
    for (; r + 3 * step < rows; r += 4 * step) {  // 4 independent loads in flight per thread
      const float v0 = __ldg(x + r * dim + col), v1 = __ldg(x + (r + step) * dim + col);
      const float v2 = __ldg(x + (r + 2 * step) * dim + col), v3 = __ldg(x + (r + 3 * step) * dim + col);
      s += ((double)v0 + (double)v1) + ((double)v2 + (double)v3);
      q += ((double)v0 * v0 + (double)v1 * v1) + ((double)v2 * v2 + (double)v3 * v3);
    }
    for (; r < rows; r += step) {
      const double v = (double)__ldg(x + r * dim + col);
      s += v;
      q += v * v;
    }
  }
  __shared__ double sh_s[kRmsThreads], sh_q[kRmsThreads];
  sh_s[threadIdx.x] = s;
  sh_q[threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.x < dim) {
    double ts = 0.0, tq = 0.0;
    for (int k = 0; k < rows_per_pass; ++k) {
      ts += sh_s[k * dim + threadIdx.x];
      tq += sh_q[k * dim + threadIdx.x];
    }
    atomicAdd(&ws.sums[threadIdx.x], ts);
    atomicAdd(&ws.sums[dim + threadIdx.x], tq);
  }
  if (last_block_ticket(ws.ticket, gridDim.x)) {
    const float n = (float)rows;
    float cnt = *count;
    __syncthreads();  // everyone has read the old count before thread 0 overwrites it
    for (int c = threadIdx.x; c < dim; c += kRmsThreads) {
      const double ts = __longlong_as_double(atomicExch((unsigned long long*)&ws.sums[c], 0ull));
      const double tq = __longlong_as_double(atomicExch((unsigned long long*)&ws.sums[dim + c], 0ull));
      const double bm = ts / (double)rows;
      double bv = tq / (double)rows - bm * bm;  // biased variance (ppo.py:30, correction=0)
      if (bv < 0.0) bv = 0.0;
      float m = mean[c], v = var[c], k = cnt;
      chan_merge(m, v, k, (float)bm, (float)bv, n);
      mean[c] = m;
      var[c] = v;
      if (c == 0) *count = k;
    }
  }
}

__global__ void __launch_bounds__(kRmsThreads)
rms_normalize_kernel(const float* __restrict__ x, long long rows, int dim, const float* __restrict__ mean,
                     const float* __restrict__ var, float eps, float* __restrict__ out, void* __restrict__ out_op,
                     int pad_op, int op_prec) {
  const int rows_per_pass = kRmsThreads / dim;
  const int active = rows_per_pass * dim;
  if (threadIdx.x >= active) return;
  const int col = threadIdx.x % dim;
  const int slot = threadIdx.x / dim;
  const float m = mean[col];
  const float d = __fsqrt_rn(__fadd_rn(var[col], eps));  // ppo.py:25
  for (long long r = (long long)blockIdx.x * rows_per_pass + slot; r < rows;
       r += (long long)gridDim.x * rows_per_pass) {
    const long long k = r * dim + col;
    const float y = __fdiv_rn(__fsub_rn(x[k], m), d);
    out[k] = y;
    if (out_op != nullptr) {  // operand copy in the zero-padded row layout the first GEMM layer reads
      if (op_prec == CATB200_PREC_BF16) {
        bf16* o = static_cast<bf16*>(out_op);
        o[r * pad_op + col] = __float2bfloat16(y);
        if (col + dim < pad_op) o[r * pad_op + dim + col] = __float2bfloat16(0.0f);  // pad_op - dim <= dim
      } else {  // fp32 rounded to tf32 (cvt.rna), what tcgen05.mma kind::tf32 reads exactly
        float* o = static_cast<float*>(out_op);
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(t) : "f"(y));
        o[r * pad_op + col] = __uint_as_float(t);
        if (col + dim < pad_op) o[r * pad_op + dim + col] = 0.0f;
      }
    }
  }
}

__global__ void rollout_append_kernel(const float* __restrict__ reward, const float* __restrict__ done,
                                      const uint8_t* __restrict__ time_out, int n, float* __restrict__ rewards_t,
                                      float* __restrict__ dones_t1, float* __restrict__ true_dones_t1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rewards_t[i] = reward[i];
  dones_t1[i] = done[i];
  true_dones_t1[i] = time_out[i] ? 1.0f : 0.0f;
}

}  // namespace catb200

using namespace catb200;

extern "C" {

size_t catb200_rms_workspace_bytes(int32_t dim) { return dim > 0 ? 256 + sizeof(double) * 2 * (size_t)dim : 0; }

int catb200_rms_forward(const float* x, int64_t rows, int32_t dim, float* mean, float* var, float* count, float eps,
                        int32_t update, float* out, void* out_op, int32_t pad_op, int32_t op_prec, void* workspace,
                        size_t workspace_bytes, void* stream) {
  if (!x || rows <= 0 || dim <= 0 || dim > kRmsThreads || !mean || !var || !count) return CATB200_ERR_INVALID_ARGUMENT;
  if (out_op && (!out || pad_op < dim || pad_op - dim > dim)) return CATB200_ERR_INVALID_ARGUMENT;
  if (out_op && op_prec != CATB200_PREC_BF16 && op_prec != CATB200_PREC_TF32) return CATB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  const int rows_per_pass = kRmsThreads / dim;
  long long want = (rows + rows_per_pass - 1) / rows_per_pass;
  if (update) {
    if (!workspace) return CATB200_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < catb200_rms_workspace_bytes(dim)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
    // a few row passes per CTA keep the atomics per column low while still filling the SMs
    const int grid = (int)min((long long)kNumSMs * 4, max(1ll, (want + 3) / 4));
    rms_moments_kernel<<<grid, kRmsThreads, 0, st>>>(x, rows, dim, mean, var, count, rms_carve(workspace, dim));
    CATB200_LAUNCH_CHECK();
  }
  if (out) {
    const int grid = (int)min((long long)kNumSMs * 8, max(1ll, (want + 1) / 2));
    rms_normalize_kernel<<<grid, kRmsThreads, 0, st>>>(x, rows, dim, mean, var, eps, out, out_op, pad_op, op_prec);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

int catb200_rollout_append(const float* reward, const float* done, const uint8_t* time_out, int32_t num_envs,
                           float* rewards_t, float* dones_t1, float* true_dones_t1, void* stream) {
  if (!reward || !done || !time_out || num_envs <= 0 || !rewards_t || !dones_t1 || !true_dones_t1)
    return CATB200_ERR_INVALID_ARGUMENT;
  rollout_append_kernel<<<(num_envs + 255) / 256, 256, 0, as_stream(stream)>>>(reward, done, time_out, num_envs,
                                                                               rewards_t, dones_t1, true_dones_t1);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

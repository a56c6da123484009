// GAE with float dones + the value-normalisation statistics, one launch (sm_100a).
//
// Replaces the 24 sequential iterations x ~10 eager kernels of the reference's GAE loop
// (U/cleanrl/ppo.py:251-277) and the two full-buffer reductions of the value RunningMeanStd updates
// (ppo.py:287-288).
//
// Layout: all rollout buffers are time-major [T, N] with envs contiguous, so one thread per env reads
// and writes fully coalesced; the recurrence over t is sequential per env (24 dependent steps of 4
// flops) while the loads of different t are independent and are issued ahead of the dependent chain.
// Arithmetic is rounded exactly like the eager reference (separate multiplies, no FMA), so advantages
// and returns are bit-identical to torch's.  HBM-bound: 24*T*N + 12*N algorithmic bytes.
#include "common.cuh"

namespace catb200 {

constexpr int kGaeThreads = 64;
// time steps whose loads are issued together: 8 when the grid fills the machine many times over (registers stay
// low, occupancy hides the latency), 24 at small N (one DRAM round trip for the whole T = 24 horizon)
constexpr int kGaeChunkLargeN = 8;
constexpr int kGaeChunkSmallN = 24;

constexpr int kGaeGroups = 32;  // spread the 4 double accumulators: same-address atomics serialise in L2

struct GaeWorkspace {
  unsigned int* ticket;
  double* sums;  // [kGaeGroups][4]: sum v, sum v^2, sum ret, sum ret^2
};

__device__ __forceinline__ void chan_merge_scalar(float& mean, float& var, float& count, float bmean, float bvar,
                                                  float n) {
  // same operation order as update_mean_var_count_from_moments (ppo.py:51-62)
  const float delta = __fsub_rn(bmean, mean);
  const float tot = __fadd_rn(count, n);
  const float new_mean = __fadd_rn(mean, __fdiv_rn(__fmul_rn(delta, n), tot));
  const float m_a = __fmul_rn(var, count);
  const float m_b = __fmul_rn(bvar, n);
  const float cross = __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(delta, delta), count), n), tot);
  mean = new_mean;
  var = __fdiv_rn(__fadd_rn(__fadd_rn(m_a, m_b), cross), tot);
  count = tot;
}

template <int kGaeChunk>
__global__ void __launch_bounds__(kGaeThreads)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ dones,
           const float* __restrict__ true_dones, const float* __restrict__ next_value, int T, int N, float gamma,
           float gamma_lambda, float* __restrict__ advantages, float* __restrict__ returns,
           float* __restrict__ value_rms, float* __restrict__ norm_stats, GaeWorkspace ws) {
  const int i = blockIdx.x * kGaeThreads + threadIdx.x;
  double sv = 0.0, qv = 0.0, sr = 0.0, qr = 0.0;
  if (i < N) {
    float last = 0.0f;
    float nv = __ldg(next_value + i);
    int t = T - 1;
    while (t >= 0) {
      const int n = min(kGaeChunk, t + 1);
      float r[kGaeChunk], v[kGaeChunk], d[kGaeChunk], td[kGaeChunk];
#pragma unroll
      for (int k = 0; k < kGaeChunk; ++k) {
        if (k < n) {
          const size_t o = (size_t)(t - k) * N + i;
          r[k] = __ldcs(rewards + o);
          v[k] = __ldcs(values + o);
          d[k] = __ldcs(dones + o + N);        // dones[t+1]; slot T holds next_done
          td[k] = __ldcs(true_dones + o + N);  // true_dones[t+1]
        }
      }
#pragma unroll
      for (int k = 0; k < kGaeChunk; ++k) {
        if (k < n) {
          const float nnt = __fsub_rn(1.0f, d[k]);    // ppo.py:256,260
          const float tnnt = __fsub_rn(1.0f, td[k]);  // ppo.py:257,261
          const float boot = __fmul_rn(__fmul_rn(__fmul_rn(gamma, nv), nnt), tnnt);
          const float delta = __fsub_rn(__fadd_rn(r[k], boot), v[k]);  // ppo.py:264-268
          const float carry = __fmul_rn(__fmul_rn(__fmul_rn(gamma_lambda, nnt), tnnt), last);
          last = __fadd_rn(delta, carry);                               // ppo.py:269-276
          const float ret = __fadd_rn(last, v[k]);                      // ppo.py:277
          const size_t o = (size_t)(t - k) * N + i;
          __stcs(advantages + o, last);
          __stcs(returns + o, ret);
          nv = v[k];
          sv += (double)v[k];
          qv += (double)v[k] * (double)v[k];
          sr += (double)ret;
          qr += (double)ret * (double)ret;
        }
      }
      t -= n;
    }
  }
  if (value_rms == nullptr) return;

  // ---- value RunningMeanStd: update with all values, then with all returns (ppo.py:287-288)
  __shared__ double sh[4][kGaeThreads / 32];
  sv = warp_sum(sv); qv = warp_sum(qv); sr = warp_sum(sr); qr = warp_sum(qr);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[0][warp] = sv; sh[1][warp] = qv; sh[2][warp] = sr; sh[3][warp] = qr; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < kGaeThreads / 32; ++w) a += sh[threadIdx.x][w];
    atomicAdd(&ws.sums[(blockIdx.x & (kGaeGroups - 1)) * 4 + threadIdx.x], a);
  }
  if (last_block_ticket_grouped(ws.ticket, gridDim.x)) {
    if (threadIdx.x == 0) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int gi = 0; gi < kGaeGroups; ++gi)
        for (int k = 0; k < 4; ++k) s[k] += __longlong_as_double(atomicExch((unsigned long long*)&ws.sums[gi * 4 + k], 0ull));
      const double cnt = (double)T * (double)N;
      const float n = (float)((long long)T * (long long)N);
      float mean = value_rms[0], var = value_rms[1], count = value_rms[2];
      double bm = s[0] / cnt, bv = fmax(s[1] / cnt - bm * bm, 0.0);
      chan_merge_scalar(mean, var, count, (float)bm, (float)bv, n);
      norm_stats[0] = mean;
      norm_stats[1] = var;
      bm = s[2] / cnt;
      bv = fmax(s[3] / cnt - bm * bm, 0.0);
      chan_merge_scalar(mean, var, count, (float)bm, (float)bv, n);
      norm_stats[2] = mean;
      norm_stats[3] = var;
      value_rms[0] = mean;
      value_rms[1] = var;
      value_rms[2] = count;
    }
  }
}

// ---- single-`dones` GAE of the rl_games / skrl front-ends ----------------------------
// VARIANT 0 (rl_games A2CBase.discount_values as CaTA2CAgent calls it, U/rl_games/cat_common.py:96-103, with the
//   float `dones` of CaTExperienceBuffer, U/rl_games/cat_experience.py:27-33): dones[t] is the flag observed BEFORE
//   step t, so step t bootstraps through dones[t+1] (last step: last_dones):
//     nnt = 1 - d;  delta = r + gamma * nv * nnt - v;  adv = last = delta + (gamma*tau) * nnt * last
// VARIANT 1 (skrl compute_gae with the CaT change `not_dones = 1 - dones`, U/skrl/ppo.py:397-442): dones[t] is the
//   termination probability of step t itself:
//     adv = (r - v) + (gamma * (1 - d[t])) * (nv + lambda * adv)
// Both: returns = adv + values.  Rounded operation by operation like the eager code (no FMA), bit-identical.
// Optional (skrl :440): advantages = (adv - mean) / (std + 1e-8) over all T*N entries, unbiased std; the
// statistics are reduced in double precision here and applied by adv_normalize_kernel.
struct GaeFdWorkspace {
  unsigned int* ticket;
  double* sums;   // [kGaeGroups][2]: sum adv, sum adv^2
  float* stats;   // mean, std + 1e-8
};

template <int VARIANT, int kGaeChunk>
__global__ void __launch_bounds__(kGaeThreads)
gae_fd_kernel(const float* __restrict__ rewards, const float* __restrict__ values, const float* __restrict__ dones,
              const float* __restrict__ last_dones, const float* __restrict__ last_values, int T, int N, float gamma,
              float coef, float* __restrict__ advantages, float* __restrict__ returns, int normalize,
              GaeFdWorkspace ws) {
  const int i = blockIdx.x * kGaeThreads + threadIdx.x;
  double sa = 0.0, qa = 0.0;
  if (i < N) {
    float last = 0.0f;
    float nv = __ldg(last_values + i);
    int t = T - 1;
    while (t >= 0) {
      const int n = min(kGaeChunk, t + 1);
      float r[kGaeChunk], v[kGaeChunk], d[kGaeChunk];
#pragma unroll
      for (int k = 0; k < kGaeChunk; ++k) {
        if (k < n) {
          const size_t o = (size_t)(t - k) * N + i;
          r[k] = __ldcs(rewards + o);
          v[k] = __ldcs(values + o);
          if (VARIANT == 0) d[k] = (t - k == T - 1) ? __ldg(last_dones + i) : __ldcs(dones + o + N);
          else d[k] = __ldcs(dones + o);
        }
      }
#pragma unroll
      for (int k = 0; k < kGaeChunk; ++k) {
        if (k < n) {
          const float nnt = __fsub_rn(1.0f, d[k]);
          if (VARIANT == 0) {
            const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, nv), nnt)), v[k]);
            last = __fadd_rn(delta, __fmul_rn(__fmul_rn(coef, nnt), last));  // coef = fl32(gamma * tau)
          } else {
            const float inner = __fadd_rn(nv, __fmul_rn(coef, last));        // coef = fl32(lambda)
            last = __fadd_rn(__fsub_rn(r[k], v[k]), __fmul_rn(__fmul_rn(gamma, nnt), inner));
          }
          const size_t o = (size_t)(t - k) * N + i;
          __stcs(returns + o, __fadd_rn(last, v[k]));
          if (normalize) advantages[o] = last;  // re-read by the normalisation pass: keep it in L2
          else __stcs(advantages + o, last);
          nv = v[k];
          sa += (double)last;
          qa += (double)last * (double)last;
        }
      }
      t -= n;
    }
  }
  if (!normalize) return;
  __shared__ double sh[2][kGaeThreads / 32];
  sa = warp_sum(sa);
  qa = warp_sum(qa);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sh[0][warp] = sa; sh[1][warp] = qa; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < kGaeThreads / 32; ++w) a += sh[threadIdx.x][w];
    atomicAdd(&ws.sums[(blockIdx.x & (kGaeGroups - 1)) * 2 + threadIdx.x], a);
  }
  if (last_block_ticket_grouped(ws.ticket, gridDim.x)) {
    if (threadIdx.x == 0) {
      double s[2] = {0.0, 0.0};
      for (int gi = 0; gi < kGaeGroups; ++gi)
        for (int k = 0; k < 2; ++k) s[k] += __longlong_as_double(atomicExch((unsigned long long*)&ws.sums[gi * 2 + k], 0ull));
      const double cnt = (double)T * (double)N;
      const double mean = s[0] / cnt;
      const double var = cnt > 1.0 ? fmax((s[1] - cnt * mean * mean) / (cnt - 1.0), 0.0) : __longlong_as_double(0x7ff8000000000000ull);
      ws.stats[0] = (float)mean;
      ws.stats[1] = __fadd_rn((float)sqrt(var), 1e-8f);
    }
  }
}

__global__ void __launch_bounds__(256)
adv_normalize_kernel(float* __restrict__ adv, size_t n, const float* __restrict__ stats) {
  const float mean = stats[0], denom = stats[1];
  for (size_t k = (size_t)blockIdx.x * 256 + threadIdx.x; k < n; k += (size_t)gridDim.x * 256)
    adv[k] = __fdiv_rn(__fsub_rn(adv[k], mean), denom);
}

}  // namespace catb200

using namespace catb200;

extern "C" {

size_t catb200_gae_workspace_bytes(void) { return 256 + kGaeGroups * 4 * sizeof(double); }

int catb200_gae(const float* rewards, const float* values, const float* dones, const float* true_dones,
                const float* next_value, int32_t T, int32_t num_envs, float gamma, float gamma_lambda,
                float* advantages, float* returns, float* value_rms, float* norm_stats, void* workspace,
                size_t workspace_bytes, void* stream) {
  if (!rewards || !values || !dones || !true_dones || !next_value || T <= 0 || num_envs <= 0 || !advantages || !returns)
    return CATB200_ERR_INVALID_ARGUMENT;
  GaeWorkspace ws = {nullptr, nullptr};
  if (value_rms) {
    if (!norm_stats || !workspace) return CATB200_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < catb200_gae_workspace_bytes()) return CATB200_ERR_WORKSPACE_TOO_SMALL;
    ws.ticket = reinterpret_cast<unsigned int*>(workspace);
    ws.sums = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  }
  const int grid = (num_envs + kGaeThreads - 1) / kGaeThreads;
  if (grid <= 16 * kNumSMs)
    gae_kernel<kGaeChunkSmallN><<<grid, kGaeThreads, 0, as_stream(stream)>>>(
        rewards, values, dones, true_dones, next_value, T, num_envs, gamma, gamma_lambda, advantages, returns, value_rms,
        norm_stats, ws);
  else
    gae_kernel<kGaeChunkLargeN><<<grid, kGaeThreads, 0, as_stream(stream)>>>(
        rewards, values, dones, true_dones, next_value, T, num_envs, gamma, gamma_lambda, advantages, returns, value_rms,
        norm_stats, ws);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

size_t catb200_gae_float_dones_workspace_bytes(void) { return 256 + kGaeGroups * 2 * sizeof(double) + 64; }

int catb200_gae_float_dones(int32_t variant, const float* rewards, const float* values, const float* dones,
                            const float* last_dones, const float* last_values, int32_t T, int32_t num_envs,
                            float gamma, float coef, float* advantages, float* returns, int32_t normalize,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!rewards || !values || !dones || !last_values || T <= 0 || num_envs <= 0 || !advantages || !returns)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (variant != CATB200_GAE_RLGAMES && variant != CATB200_GAE_SKRL) return CATB200_ERR_UNSUPPORTED;
  if (variant == CATB200_GAE_RLGAMES && !last_dones) return CATB200_ERR_INVALID_ARGUMENT;
  GaeFdWorkspace ws = {nullptr, nullptr, nullptr};
  if (normalize) {
    if (!workspace) return CATB200_ERR_INVALID_ARGUMENT;
    if (workspace_bytes < catb200_gae_float_dones_workspace_bytes()) return CATB200_ERR_WORKSPACE_TOO_SMALL;
    char* p = static_cast<char*>(workspace);
    ws.ticket = reinterpret_cast<unsigned int*>(p);
    ws.sums = reinterpret_cast<double*>(p + 256);
    ws.stats = reinterpret_cast<float*>(p + 256 + kGaeGroups * 2 * sizeof(double));
  }
  cudaStream_t st = as_stream(stream);
  const int grid = (num_envs + kGaeThreads - 1) / kGaeThreads;
  const bool small = grid <= 16 * kNumSMs;
#define CATB200_GAE_FD(V, C)                                                                                          \
  gae_fd_kernel<V, C><<<grid, kGaeThreads, 0, st>>>(rewards, values, dones, last_dones, last_values, T, num_envs, gamma, \
                                                    coef, advantages, returns, normalize, ws)
  if (variant == CATB200_GAE_RLGAMES) {
    if (small) CATB200_GAE_FD(0, kGaeChunkSmallN); else CATB200_GAE_FD(0, kGaeChunkLargeN);
  } else {
    if (small) CATB200_GAE_FD(1, kGaeChunkSmallN); else CATB200_GAE_FD(1, kGaeChunkLargeN);
  }
#undef CATB200_GAE_FD
  CATB200_LAUNCH_CHECK();
  if (normalize) {
    const size_t n = (size_t)T * (size_t)num_envs;
    const int ngrid = (int)min((size_t)(kNumSMs * 8), (n + 255) / 256);
    adv_normalize_kernel<<<ngrid, 256, 0, st>>>(advantages, n, ws.stats);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

}  // extern "C"

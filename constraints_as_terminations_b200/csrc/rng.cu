// Device-side random draws of the trainer and of the env-side MDP terms (sm_100a), all from Philox4x32-10 (philox.cuh).
//
//   catb200_random_permutation : the per-epoch minibatch shuffle (reference U/cleanrl/ppo.py:295 `torch.randperm`) as a
//                                keyed Feistel bijection with cycle walking -- no sort, one pass, int64 indices
//   catb200_bernoulli_mask     : optional stochastic termination mask, mask[i] = u_i < p_i, plus the ascending index list
//   catb200_command_update     : UniformVelocityCommandWithDeadzone._update_command (U/mdp/commands.py:39-93)
//   catb200_push_select        : push_by_setting_velocity_with_random_envs (U/mdp/events.py:59-96): Bernoulli selection
//                                + uniform velocities for the selected envs
// `rng_state` = {seed, offset}: two uint64 in device memory; every call advances offset on the device (graph-safe).
#include "common.cuh"
#include "philox.cuh"

namespace catb200 {

struct RoundKeys {
  uint32_t rk[8];
};

__global__ void __launch_bounds__(256)
permutation_kernel(long long n, int bits, const unsigned long long* __restrict__ rng_state, long long* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  // round keys of this draw: two Philox blocks at counter offset, offset + 1 of the permutation stream
  __shared__ uint32_t rk_s[8];
  if (threadIdx.x < 2) {
    uint32_t w[4];
    philox4x32_10(rng_state[0], kStreamPermutation, rng_state[1] + threadIdx.x, w);
#pragma unroll
    for (int i = 0; i < 4; ++i) rk_s[threadIdx.x * 4 + i] = w[i];
  }
  __syncthreads();
  uint32_t rk[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rk[i] = rk_s[i];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i;
    do {
      x = feistel_bijection(x, bits, rk);
    } while ((long long)x >= n);  // cycle walking: a bijection of [0, 2^bits) restricted to [0, n) stays a bijection
    out[i] = (long long)x;
  }
}

__global__ void rng_advance_kernel(unsigned long long* rng_state, unsigned long long n) { rng_state[1] += n; }

// mask[i] = u_i < p[i] with u_i = uniform(seed, kStreamBernoulli, offset + i)
__global__ void __launch_bounds__(256)
bernoulli_mask_kernel(const float* __restrict__ p, int n, const unsigned long long* __restrict__ rng_state,
                      uint8_t* __restrict__ mask) {
  const unsigned long long seed = rng_state[0], off = rng_state[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    mask[i] = philox_uniform(seed, kStreamBernoulli, off + i) < p[i] ? 1 : 0;
}

// ascending list of the set positions of mask[0..n) (what `mask.nonzero().flatten()` returns) + their count; one CTA
// walks the mask in chunks of 1024 with a ballot / prefix-count scan, so the order is deterministic
__global__ void __launch_bounds__(1024)
compact_mask_kernel(const uint8_t* __restrict__ mask, int n, long long* __restrict__ ids, int* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + threadIdx.x;
    const bool set = i < n && mask[i] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, set);
    if (lane == 0) warp_tot[warp] = __popc(b);
    __syncthreads();
    int before = base_s;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    if (set) ids[before + __popc(b & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_tot[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base_s;
}

// ---- velocity command post-processing (U/mdp/commands.py:39-93) --------------------------------------------
// One thread per env.  Draw k of env i uses counter offset + 8 * i + k of the uniform stream:
//   0 resample Bernoulli, 1..3 new lin_x / lin_y / ang_z, 4 new heading, 5 is_heading, 6 is_standing, 7 yaw-flip Bernoulli
// `u_ext` (N x 8, optional) replaces the Philox draws (parity tests feed the oracle's numbers).
__global__ void __launch_bounds__(256)
command_update_kernel(const catb200_command_cfg_t c, int n, float* __restrict__ cmd, float* __restrict__ heading_target,
                      const float* __restrict__ heading_w, uint8_t* __restrict__ is_heading, uint8_t* __restrict__ is_standing,
                      const float* __restrict__ u_ext, const unsigned long long* __restrict__ rng_state,
                      uint8_t* __restrict__ resampled) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float u[8];
  if (u_ext) {
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = u_ext[(size_t)i * 8 + k];
  } else {
    const unsigned long long seed = rng_state[0], off = rng_state[1] + 8ull * i;
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = philox_uniform(seed, kStreamUniform, off + k);
  }
  float x = cmd[i * 3], y = cmd[i * 3 + 1], w = cmd[i * 3 + 2];
  // angular velocity from the heading error (commands.py:46-58): wrap_to_pi, stiffness, clip to the ang_vel_z range
  if (c.heading_command && is_heading[i]) {
    const float kPi = 3.14159265358979323846f;
    float err = heading_target[i] - heading_w[i];
    // isaaclab.utils.math.wrap_to_pi: ((a + pi) mod 2 pi) - pi with +pi kept at +pi
    float wrapped = fmodf(err + kPi, 2.0f * kPi);
    if (wrapped < 0.0f) wrapped += 2.0f * kPi;
    wrapped -= kPi;
    if (wrapped == -kPi && err > 0.0f) wrapped = kPi;
    w = fminf(fmaxf(c.heading_control_stiffness * wrapped, c.ang_vel_z[0]), c.ang_vel_z[1]);
  }
  // small commands to zero (commands.py:60-65)
  const bool any_big = fabsf(x) > c.velocity_deadzone || fabsf(y) > c.velocity_deadzone || fabsf(w) > c.velocity_deadzone;
  if (!any_big) { x *= 0.0f; y *= 0.0f; w *= 0.0f; }
  // random resampling (commands.py:67-78): p = 0.01 for still commands, dt / T_episode otherwise
  const float nrm = sqrtf(fmaf(w, w, fmaf(y, y, x * x)));
  const float no_vel = nrm < c.velocity_deadzone ? 1.0f : 0.0f;
  const float p = __fadd_rn(__fmul_rn(0.01f, no_vel), __fmul_rn(c.p_step, 1.0f - no_vel));
  const bool res = u[0] < p;
  if (res) {  // UniformVelocityCommand._resample_command (Isaac Lab, third party): uniform draws in the cfg ranges
    // lo + (hi - lo) * u with separately rounded multiply and add (no FMA contraction), like the eager tensor ops
    x = __fadd_rn(c.lin_vel_x[0], __fmul_rn(c.lin_vel_x[1] - c.lin_vel_x[0], u[1]));
    y = __fadd_rn(c.lin_vel_y[0], __fmul_rn(c.lin_vel_y[1] - c.lin_vel_y[0], u[2]));
    w = __fadd_rn(c.ang_vel_z[0], __fmul_rn(c.ang_vel_z[1] - c.ang_vel_z[0], u[3]));
    if (c.heading_command) {
      heading_target[i] = __fadd_rn(c.heading[0], __fmul_rn(c.heading[1] - c.heading[0], u[4]));
      is_heading[i] = u[5] <= c.rel_heading_envs ? 1 : 0;
    }
    is_standing[i] = u[6] <= c.rel_standing_envs ? 1 : 0;
  }
  // random yaw-rate inversion (commands.py:80-93)
  if (u[7] < c.p_step) w *= -1.0f;
  cmd[i * 3] = x; cmd[i * 3 + 1] = y; cmd[i * 3 + 2] = w;
  if (resampled) resampled[i] = res ? 1 : 0;
}

// ---- random push (U/mdp/events.py:59-96) -----------------------------------------------------------------------
// push[i] = u_{i,0} < p_push; selected envs get root_vel_w[i, k] = lo_k + (hi_k - lo_k) * u_{i,1+k}, k = 0..5
__global__ void __launch_bounds__(256)
push_select_kernel(int n, float p_push, const float* __restrict__ lo, const float* __restrict__ hi,
                   float* __restrict__ root_vel_w, const float* __restrict__ u_ext,
                   const unsigned long long* __restrict__ rng_state, uint8_t* __restrict__ pushed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float u[7];
  if (u_ext) {
#pragma unroll
    for (int k = 0; k < 7; ++k) u[k] = u_ext[(size_t)i * 7 + k];
  } else {
    const unsigned long long seed = rng_state[0], off = rng_state[1] + 8ull * i;
#pragma unroll
    for (int k = 0; k < 7; ++k) u[k] = philox_uniform(seed, kStreamUniform, off + k);
  }
  const bool push = u[0] < p_push;
  if (push) {
#pragma unroll
    for (int k = 0; k < 6; ++k) root_vel_w[(size_t)i * 6 + k] = __fadd_rn(lo[k], __fmul_rn(hi[k] - lo[k], u[1 + k]));
  }
  pushed[i] = push ? 1 : 0;
}

// ---- policy observation assembly (S12/cat_flat_env_cfg.py:137-172 through Isaac Lab's ObservationManager) ---------
__global__ void __launch_bounds__(256)
obs_assemble_kernel(const __grid_constant__ catb200_obs_plan_t plan, int n, float* __restrict__ out,
                    const float* __restrict__ u_ext, const unsigned long long* __restrict__ rng_state) {
  const long long total = (long long)n * plan.n_cols;
  unsigned long long seed = 0, off = 0;
  if (!u_ext && rng_state) { seed = rng_state[0]; off = rng_state[1]; }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / plan.n_cols), c = (int)(e - (long long)i * plan.n_cols);
    const catb200_obs_term_t& t = plan.terms[plan.col_term[c]];
    const int k = plan.col_idx[c];
    float v = __ldg(t.src + (size_t)i * t.row_stride + t.ids[k]);
    if (t.n_max > t.n_min) {  // data + rand * (n_max - n_min) + n_min, each operation rounded separately like the eager ops
      const float u = u_ext ? u_ext[e] : philox_uniform(seed, kStreamUniform, off + (unsigned long long)e);
      v = __fadd_rn(__fadd_rn(v, __fmul_rn(u, t.noise_span)), t.n_min);
    }
    out[e] = __fmul_rn(v, t.scale[k]);
  }
}

}  // namespace catb200

using namespace catb200;

extern "C" {

// host-side evaluation of the same Philox4x32-10 code the kernels inline (no GPU needed): known-answer tests
int catb200_philox4x32_10(const uint32_t* counter4, const uint32_t* key2, uint32_t* out4) {
  if (!counter4 || !key2 || !out4) return CATB200_ERR_INVALID_ARGUMENT;
  uint32_t c[4] = {counter4[0], counter4[1], counter4[2], counter4[3]};
  philox4x32_10_raw(c, key2[0], key2[1]);
  for (int i = 0; i < 4; ++i) out4[i] = c[i];
  return CATB200_OK;
}

// host-side evaluation of the permutation the kernel computes for (n, seed, offset): out[0..n)
int catb200_random_permutation_host(int64_t n, uint64_t seed, uint64_t offset, int64_t* out) {
  if (n <= 0 || n > (1ll << 31) || !out) return CATB200_ERR_INVALID_ARGUMENT;
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  uint32_t rk[8];
  for (int b = 0; b < 2; ++b) {
    uint32_t w[4];
    philox4x32_10(seed, kStreamPermutation, offset + b, w);
    for (int i = 0; i < 4; ++i) rk[b * 4 + i] = w[i];
  }
  for (int64_t i = 0; i < n; ++i) {
    uint32_t x = (uint32_t)i;
    do {
      x = feistel_bijection(x, bits, rk);
    } while ((int64_t)x >= n);
    out[i] = (int64_t)x;
  }
  return CATB200_OK;
}

int catb200_random_permutation(int64_t n, uint64_t* rng_state, int64_t* out, void* stream) {
  if (n <= 0 || n > (1ll << 31) || !rng_state || !out) return CATB200_ERR_INVALID_ARGUMENT;
  int bits = 1;
  while ((1ll << bits) < n) ++bits;
  cudaStream_t st = as_stream(stream);
  const int grid = (int)min((long long)(n + 255) / 256, (long long)kNumSMs * 8);
  CATB200_CUDA_TRY(launch_pdl(permutation_kernel, dim3(grid), dim3(256), 0, st, (long long)n, bits,
                              (const unsigned long long*)rng_state, (long long*)out));
  CATB200_LAUNCH_CHECK();
  rng_advance_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state, 2ull);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_bernoulli_mask(const float* p, int32_t n, uint64_t* rng_state, uint8_t* mask, int64_t* ids, int32_t* count,
                           void* stream) {
  if (!p || n <= 0 || !rng_state || !mask) return CATB200_ERR_INVALID_ARGUMENT;
  if ((ids == nullptr) != (count == nullptr)) return CATB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  bernoulli_mask_kernel<<<min((n + 255) / 256, kNumSMs * 8), 256, 0, st>>>(p, n, (const unsigned long long*)rng_state, mask);
  CATB200_LAUNCH_CHECK();
  if (ids) {
    compact_mask_kernel<<<1, 1024, 0, st>>>(mask, n, (long long*)ids, count);
    CATB200_LAUNCH_CHECK();
  }
  rng_advance_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state, (unsigned long long)n);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_command_update(const catb200_command_cfg_t* cfg, int32_t num_envs, float* vel_command_b, float* heading_target,
                           const float* heading_w, uint8_t* is_heading_env, uint8_t* is_standing_env, const float* u_ext,
                           uint64_t* rng_state, uint8_t* resampled, void* stream) {
  if (!cfg || num_envs <= 0 || !vel_command_b || !is_standing_env) return CATB200_ERR_INVALID_ARGUMENT;
  if (cfg->heading_command && (!heading_target || !heading_w || !is_heading_env)) return CATB200_ERR_INVALID_ARGUMENT;
  if (!u_ext && !rng_state) return CATB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  command_update_kernel<<<(num_envs + 255) / 256, 256, 0, st>>>(*cfg, num_envs, vel_command_b, heading_target, heading_w,
                                                                 is_heading_env, is_standing_env, u_ext,
                                                                 (const unsigned long long*)rng_state, resampled);
  CATB200_LAUNCH_CHECK();
  if (!u_ext) {
    rng_advance_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state, 8ull * num_envs);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

int catb200_obs_assemble(catb200_obs_plan_t* plan, int32_t num_envs, float* obs_out, const float* u_ext,
                         uint64_t* rng_state, void* stream) {
  if (!plan || num_envs <= 0 || !obs_out) return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_terms <= 0 || plan->n_terms > CATB200_OBS_MAX_TERMS) return CATB200_ERR_INVALID_ARGUMENT;
  int col = 0;
  bool noisy = false;
  for (int t = 0; t < plan->n_terms; ++t) {
    const catb200_obs_term_t& term = plan->terms[t];
    if (!term.src || term.n_cols <= 0 || term.n_cols > 32 || col + term.n_cols > CATB200_OBS_MAX_COLS) return CATB200_ERR_INVALID_ARGUMENT;
    noisy = noisy || term.n_max > term.n_min;
    for (int k = 0; k < term.n_cols; ++k, ++col) {
      plan->col_term[col] = (uint8_t)t;
      plan->col_idx[col] = (uint8_t)k;
    }
  }
  plan->n_cols = col;
  if (noisy && !u_ext && !rng_state) return CATB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)num_envs * col;
  obs_assemble_kernel<<<(int)min((total + 255) / 256, (long long)kNumSMs * 16), 256, 0, st>>>(*plan, num_envs, obs_out, u_ext,
                                                                                             (const unsigned long long*)rng_state);
  CATB200_LAUNCH_CHECK();
  if (noisy && !u_ext) {
    rng_advance_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state, (unsigned long long)total);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

int catb200_push_select(int32_t num_envs, float p_push, const float* range_lo, const float* range_hi, float* root_vel_w,
                        const float* u_ext, uint64_t* rng_state, uint8_t* pushed, void* stream) {
  if (num_envs <= 0 || !range_lo || !range_hi || !root_vel_w || !pushed) return CATB200_ERR_INVALID_ARGUMENT;
  if (!u_ext && !rng_state) return CATB200_ERR_INVALID_ARGUMENT;
  cudaStream_t st = as_stream(stream);
  push_select_kernel<<<(num_envs + 255) / 256, 256, 0, st>>>(num_envs, p_push, range_lo, range_hi, root_vel_w, u_ext,
                                                              (const unsigned long long*)rng_state, pushed);
  CATB200_LAUNCH_CHECK();
  if (!u_ext) {
    rng_advance_kernel<<<1, 1, 0, st>>>((unsigned long long*)rng_state, 8ull * num_envs);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

}  // extern "C"

// Per-step CaT path on sm_100a: constraint terms -> termination probabilities.
//
// Replaces, in two launches, the ~250-300 eager kernels + 13 host syncs of the reference's
// ConstraintManager.compute() (U/cat/constraint_manager.py:213-229 driving constraints.py:23-235 and
// CaT.add/get_probs :39-82) and the reward/dones lines of CaTEnv.step (U/cat/cat_env.py:102-121).
//
// Data flow (N envs, K constraint columns, S statistics slots):
//   cat_eval_kernel : one CTA per 32-env tile.  Every source tensor row block of the tile is staged
//                     into shared memory with coalesced loads (rows padded to an odd pitch -> lane r
//                     reading row r is bank-conflict free).  Warp w evaluates columns w, w+4, ... for
//                     the tile's 32 envs (one env per lane, warp-uniform op dispatch), stores the raw
//                     constraint column-major into the workspace (C_T[K][N], coalesced) and folds the
//                     column max over envs with redux.sync + one atomicMax per column per CTA.  The last
//                     CTA to finish applies the clamp + Polyak update to running_max[K] (:55-61).
//   cat_apply_kernel: one thread per env.  Reads its K constraint values back (coalesced, L2 hits),
//                     maps violations to probabilities (:64-72), takes the per-term and overall row
//                     max (:82,:225), updates the two per-term episode statistics (:226-227) and writes
//                     cstr_prob plus, optionally, the scaled reward and float dones.
//
// The cross-env column max is a true global dependency (probability of env i depends on the max over
// all envs of this step), hence two phases.  HBM-bound streaming work: no tensor cores involved.
#include "common.cuh"

namespace catb200 {

thread_local cudaError_t g_last_cuda_error = cudaSuccess;
unsigned long long g_launch_count = 0;

constexpr int kTile = 32;          // envs per CTA in the eval kernel (one per lane)
constexpr int kEvalWarps = 4;      // warps per CTA; warp w owns columns w, w+4, ...
constexpr int kEvalThreads = kEvalWarps * 32;
constexpr int kApplyThreads = 64;

__device__ __forceinline__ int pitch_of(const catb200_source_t& s) { return s.row_len | 1; }

// ---- staged-source accessors ----------------------------------------------------------------------
struct TileView {
  const float* smem;
  const catb200_plan_t* plan;
  int row;  // env within the tile == lane
  __device__ __forceinline__ float at(int src, int e) const {
    const catb200_source_t& s = plan->sources[src];
    return smem[s.smem_off * kTile + row * pitch_of(s) + e];
  }
};

// sqrt(x^2 + y^2 + z^2) the way torch.norm reduces a short contiguous dim on CPU and CUDA:
// sequential fused multiply-adds from a zero accumulator, then a correctly rounded sqrt.
__device__ __forceinline__ float norm3(float x, float y, float z) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  acc = __fmaf_rn(z, z, acc);
  return __fsqrt_rn(acc);
}
__device__ __forceinline__ float norm2(float x, float y) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  return __fsqrt_rn(acc);
}

// max over the history axis of |F[h, body, :]| for one body (constraints.py:102-107,151-158,207-209)
__device__ __forceinline__ float force_peak(const TileView& v, int src, int body) {
  const catb200_source_t& s = v.plan->sources[src];
  const int B = s.aux;
  const int H = s.row_len / (3 * B);
  float peak = -INFINITY;
  for (int h = 0; h < H; ++h) {
    const int e = (h * B + body) * 3;
    peak = fmaxf(peak, norm3(v.at(src, e), v.at(src, e + 1), v.at(src, e + 2)));
  }
  return peak;
}

__device__ __forceinline__ float command_norm(const TileView& v, int src) {
  return norm3(v.at(src, 0), v.at(src, 1), v.at(src, 2));
}

// Value of column `lc` of term `t` for the env of this lane.  Operation order follows the cited
// reference lines; every intermediate is rounded to fp32 exactly where torch materialises a tensor.
__device__ float eval_column(const TileView& v, const catb200_term_t& t, int lc) {
  switch (t.op) {
    case CATB200_OP_GENERIC:
      return v.at(t.src0, t.ids[lc]);
    case CATB200_OP_ABS_MINUS:  // constraints.py:30,64,75,85
      return __fsub_rn(fabsf(v.at(t.src0, t.ids[lc])), t.p0);
    case CATB200_OP_ABSDIFF_MINUS:  // constraints.py:176-181
      return __fsub_rn(fabsf(__fsub_rn(v.at(t.src0, t.ids[lc]), v.at(t.src1, t.ids[lc]))), t.p0);
    case CATB200_OP_ABSDIFF_MINUS_GATE_Y: {  // constraints.py:42-53
      float c = __fsub_rn(fabsf(__fsub_rn(v.at(t.src0, t.ids[lc]), v.at(t.src1, t.ids[lc]))), t.p0);
      float gate = fabsf(v.at(t.src2, 1)) < t.p1 ? 1.0f : 0.0f;
      return __fmul_rn(c, gate);
    }
    case CATB200_OP_ACTION_RATE: {  // constraints.py:191-198 (true division by step_dt)
      float d = fabsf(__fsub_rn(v.at(t.src0, t.ids[lc]), v.at(t.src1, t.ids[lc])));
      return __fsub_rn(__fdiv_rn(d, t.p1), t.p0);
    }
    case CATB200_OP_COMPONENT_GT:  // constraints.py:94
      return v.at(t.src0, t.ids[lc]) > t.p0 ? 1.0f : 0.0f;
    case CATB200_OP_CONTACT_ANY: {  // constraints.py:103-110
      bool any = false;
      for (int b = 0; b < t.n_ids; ++b) any |= force_peak(v, t.src0, t.ids[b]) > t.p0;
      return any ? 1.0f : 0.0f;
    }
    case CATB200_OP_NORM2_MINUS:  // constraints.py:119
      return __fsub_rn(norm2(v.at(t.src0, 0), v.at(t.src0, 1)), t.p0);
    case CATB200_OP_AIR_TIME: {  // constraints.py:129-141
      float td = v.at(t.src1, t.ids[lc]) != 0.0f ? 1.0f : 0.0f;
      float moving = command_norm(v, t.src2) > t.p1 ? 1.0f : 0.0f;
      float c = __fsub_rn(t.p0, v.at(t.src0, t.ids[lc]));
      return __fmul_rn(__fmul_rn(c, td), moving);
    }
    case CATB200_OP_N_CONTACT: {  // constraints.py:151-168
      int n = 0;
      for (int b = 0; b < t.n_ids; ++b) n += force_peak(v, t.src0, t.ids[b]) > t.p2 ? 1 : 0;
      float miss = fabsf((float)n - t.p0);
      float moving = command_norm(v, t.src2) > t.p1 ? 1.0f : 0.0f;
      return __fmul_rn(miss, moving);
    }
    case CATB200_OP_FORCE_PEAK_MINUS:  // constraints.py:207-210
      return __fsub_rn(force_peak(v, t.src0, t.ids[lc]), t.p0);
    case CATB200_OP_LIMIT_MINUS:  // constraints.py:220
      return __fsub_rn(t.p0, v.at(t.src0, t.ids[lc]));
    case CATB200_OP_ABS_MINUS_GATE_STILL: {  // constraints.py:231-235
      float c = __fsub_rn(fabsf(v.at(t.src0, t.ids[lc])), t.p0);
      float still = command_norm(v, t.src2) < t.p1 ? 1.0f : 0.0f;
      return __fmul_rn(c, still);
    }
    default:
      return 0.0f;
  }
}

struct CatWorkspace {
  // layout inside the caller's workspace (all 256-byte aligned)
  unsigned int* ticket;   // 1 word (padded)
  uint32_t* colmax;       // [CATB200_MAX_COLS] ordered-float column maxima, 0 between launches
  float* c_t;             // [K][N] raw constraints, column-major
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ __device__ inline CatWorkspace carve(void* base, int num_envs) {
  CatWorkspace w;
  char* p = static_cast<char*>(base);
  w.ticket = reinterpret_cast<unsigned int*>(p);
  w.colmax = reinterpret_cast<uint32_t*>(p + 256);
  w.c_t = reinterpret_cast<float*>(p + 256 + align256(sizeof(uint32_t) * CATB200_MAX_COLS));
  (void)num_envs;
  return w;
}

enum EvalMode { kEvalStep = 0, kEvalRowMajor = 1 };

template <int MODE>
__global__ void __launch_bounds__(kEvalThreads)
cat_eval_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                int num_envs, float* __restrict__ running_max, int* __restrict__ rm_init,
                CatWorkspace ws, float* __restrict__ out_rowmajor) {
  extern __shared__ float smem[];
  const int tile0 = blockIdx.x * kTile;
  const int rows = min(kTile, num_envs - tile0);

  // ---- stage every source row block of this tile (coalesced: consecutive threads, consecutive elements)
  for (int s = 0; s < plan.n_sources; ++s) {
    const catb200_source_t& src = plan.sources[s];
    const int pitch = pitch_of(src);
    float* dst = smem + src.smem_off * kTile;
    const int total = rows * src.row_len;
    if (src.dtype == CATB200_F32) {
      const float* g = static_cast<const float*>(src.ptr);
      for (int f = threadIdx.x; f < total; f += kEvalThreads) {
        const int r = src.row_len == 1 ? f : (int)__umulhi((unsigned)f, src.magic);
        const int e = f - r * src.row_len;
        dst[r * pitch + e] = __ldg(g + (size_t)(tile0 + r) * src.row_stride + e);
      }
    } else {
      const uint8_t* g = static_cast<const uint8_t*>(src.ptr);
      for (int f = threadIdx.x; f < total; f += kEvalThreads) {
        const int r = src.row_len == 1 ? f : (int)__umulhi((unsigned)f, src.magic);
        const int e = f - r * src.row_len;
        dst[r * pitch + e] = g[(size_t)(tile0 + r) * src.row_stride + e] ? 1.0f : 0.0f;
      }
    }
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool live = lane < rows;
  TileView view{smem, &plan, live ? lane : 0};
  for (int col = warp; col < plan.n_cols; col += kEvalWarps) {
    const catb200_term_t& t = plan.terms[plan.col_term[col]];
    const float c = eval_column(view, t, col - t.col_offset);
    if (MODE == kEvalRowMajor) {
      if (live) out_rowmajor[(size_t)(tile0 + lane) * plan.n_cols + col] = c;
    } else {
      if (live) ws.c_t[(size_t)col * num_envs + tile0 + lane] = c;
      const uint32_t key = live ? float_to_ordered(c) : 0u;
      const uint32_t m = __reduce_max_sync(0xffffffffu, key);
      if (lane == 0) atomicMax(&ws.colmax[col], m);
    }
  }

  if (MODE == kEvalStep) {
    // ---- the last CTA folds the column maxima into the Polyak running max (constraint_manager.py:55-61)
    if (last_block_ticket(ws.ticket, gridDim.x)) {
      for (int col = threadIdx.x; col < plan.n_cols; col += kEvalThreads) {
        const uint32_t key = atomicExch(&ws.colmax[col], 0u);
        float cmax = fmaxf(ordered_to_float(key), prm.floor_max);
        float rm;
        if (rm_init[col]) {
          rm = __fadd_rn(__fmul_rn(running_max[col], prm.tau), __fmul_rn(prm.one_minus_tau, cmax));
        } else {
          rm = cmax;
          rm_init[col] = 1;
        }
        running_max[col] = rm;
      }
    }
  }
}

// probability of one column value (constraint_manager.py:64-72)
__device__ __forceinline__ float violation_prob(float c, float rm, float min_p, float span) {
  if (!(c > 0.0f)) return 0.0f;
  float x = __fdiv_rn(c, rm);
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return __fadd_rn(min_p, __fmul_rn(x, span));
}

__global__ void __launch_bounds__(kApplyThreads)
cat_apply_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ c_t,
                 float* __restrict__ episode_sums, float* __restrict__ mean_values,
                 float* __restrict__ cstr_prob, const float* __restrict__ raw_reward,
                 const uint8_t* __restrict__ reset_buf, float* __restrict__ reward_out,
                 float* __restrict__ dones_out) {
  __shared__ float s_rm[CATB200_MAX_COLS];
  for (int c = threadIdx.x; c < plan.n_cols; c += kApplyThreads) s_rm[c] = running_max[c];
  __syncthreads();
  const int i = blockIdx.x * kApplyThreads + threadIdx.x;
  if (i >= num_envs) return;

  // All loads of an env are independent (K constraint values, 2 statistics per slot): the statistics of a
  // slot are requested before its columns and the columns four at a time, so several loads are in flight
  // per thread instead of one round trip each.
  float overall = -INFINITY;
#pragma unroll 1
  for (int slot = 0; slot < plan.n_slots; ++slot) {
    const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
    const float span = prm.span[slot];
    const size_t k = (size_t)slot * num_envs + i;
    const float es = episode_sums[k], mv = mean_values[k];
    float tmax = -INFINITY;
    int col = c0;
    for (; col + 4 <= c1; col += 4) {
      const float v0 = __ldcs(c_t + (size_t)col * num_envs + i);
      const float v1 = __ldcs(c_t + (size_t)(col + 1) * num_envs + i);
      const float v2 = __ldcs(c_t + (size_t)(col + 2) * num_envs + i);
      const float v3 = __ldcs(c_t + (size_t)(col + 3) * num_envs + i);
      tmax = fmaxf(tmax, fmaxf(fmaxf(violation_prob(v0, s_rm[col], prm.min_p, span), violation_prob(v1, s_rm[col + 1], prm.min_p, span)),
                               fmaxf(violation_prob(v2, s_rm[col + 2], prm.min_p, span), violation_prob(v3, s_rm[col + 3], prm.min_p, span))));
    }
    for (; col < c1; ++col) tmax = fmaxf(tmax, violation_prob(__ldcs(c_t + (size_t)col * num_envs + i), s_rm[col], prm.min_p, span));
    episode_sums[k] = __fadd_rn(es, tmax > 0.0f ? 1.0f : 0.0f);  // :226
    mean_values[k] = __fadd_rn(mv, tmax);                          // :227
    overall = fmaxf(overall, tmax);
  }
  cstr_prob[i] = overall;
  if (raw_reward != nullptr) {
    // cat_env.py:102-107: reward = clip(reward * (1 - p), min=0); dones = p; :121 dones[reset] = 1
    reward_out[i] = fmaxf(__fmul_rn(raw_reward[i], __fsub_rn(1.0f, overall)), 0.0f);
    dones_out[i] = (reset_buf != nullptr && reset_buf[i]) ? 1.0f : overall;
  }
}

__global__ void __launch_bounds__(kApplyThreads)
cat_probs_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ c_t,
                 float* __restrict__ probs_out) {
  const int i = blockIdx.x * kApplyThreads + threadIdx.x;
  if (i >= num_envs) return;
  for (int slot = 0; slot < plan.n_slots; ++slot) {
    const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
    for (int col = c0; col < c1; ++col) {
      const float c = c_t[(size_t)col * num_envs + i];
      probs_out[(size_t)i * plan.n_cols + col] = violation_prob(c, running_max[col], prm.min_p, prm.span[slot]);
    }
  }
}

// ---- ConstraintManager.reset (constraint_manager.py:190-211) -----------------------------------------
// grid = n_slots CTAs; each reduces its statistics row over the selected envs in double precision
// (torch's own fp32 reduction order is implementation defined; parity tolerance 1e-5 relative) and
// then zeroes the selected entries.
constexpr int kResetThreads = 256;

__global__ void __launch_bounds__(kResetThreads)
cat_reset_kernel(const int64_t* __restrict__ env_ids, int n_ids, const uint8_t* __restrict__ mask,
                 const int64_t* __restrict__ episode_length, int num_envs, float* __restrict__ episode_sums,
                 float* __restrict__ mean_values, float* __restrict__ out) {
  const int slot = blockIdx.x;
  float* sums = episode_sums + (size_t)slot * num_envs;
  float* means = mean_values + (size_t)slot * num_envs;
  double acc_v = 0.0, acc_p = 0.0;
  long long cnt = 0;
  const int total = env_ids ? n_ids : num_envs;
  for (int k = threadIdx.x; k < total; k += kResetThreads) {
    int i = k;
    if (env_ids) {
      i = (int)env_ids[k];
    } else if (mask && !mask[k]) {
      continue;
    }
    const float len = (float)episode_length[i];  // int64 -> float like torch's float / long promotion
    acc_v += (double)__fdiv_rn(sums[i], len);
    acc_p += (double)__fdiv_rn(means[i], len);
    cnt += 1;
    sums[i] = 0.0f;
    means[i] = 0.0f;
  }
  __shared__ double s_v[kResetThreads / 32], s_p[kResetThreads / 32];
  __shared__ long long s_c[kResetThreads / 32];
  acc_v = warp_sum(acc_v);
  acc_p = warp_sum(acc_p);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    s_v[threadIdx.x >> 5] = acc_v;
    s_p[threadIdx.x >> 5] = acc_p;
    s_c[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0, p = 0.0;
    long long c = 0;
    for (int w = 0; w < kResetThreads / 32; ++w) {
      v += s_v[w];
      p += s_p[w];
      c += s_c[w];
    }
    // empty selection -> mean of nothing = NaN, like torch
    const float mv = (float)(v / (double)c), mp = (float)(p / (double)c);
    out[2 * slot] = __fmul_rn(mv, 100.0f);
    out[2 * slot + 1] = mp;
  }
}

}  // namespace catb200

using namespace catb200;

extern "C" {

int catb200_version(void) { return CATB200_VERSION; }

uint64_t catb200_launch_count(void) { return g_launch_count; }

const char* catb200_error_string(int status) {
  switch (status) {
    case CATB200_OK: return "ok";
    case CATB200_ERR_INVALID_ARGUMENT: return "invalid argument";
    case CATB200_ERR_UNSUPPORTED: return "unsupported configuration";
    case CATB200_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
    case CATB200_ERR_CUDA: return cudaGetErrorString(g_last_cuda_error);
    default: return "unknown status";
  }
}

int catb200_cat_plan_finalize(catb200_plan_t* plan) {
  if (!plan) return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_sources < 0 || plan->n_sources > CATB200_MAX_SOURCES) return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_terms < 0 || plan->n_terms > CATB200_MAX_TERMS) return CATB200_ERR_INVALID_ARGUMENT;
  int off = 0;
  for (int s = 0; s < plan->n_sources; ++s) {
    catb200_source_t& src = plan->sources[s];
    if (src.row_len <= 0 || src.row_len > 4096 || src.row_stride < src.row_len) return CATB200_ERR_INVALID_ARGUMENT;
    if (src.dtype != CATB200_F32 && src.dtype != CATB200_U8) return CATB200_ERR_UNSUPPORTED;
    if (src.aux < 0 || (src.aux > 0 && src.row_len % (3 * src.aux) != 0)) return CATB200_ERR_INVALID_ARGUMENT;
    src.smem_off = off;
    off += src.row_len | 1;
    // exact floor(f / row_len) for f < 2^17 via umulhi (f * ceil(2^32 / d)) -- tile * row_len <= 131072
    src.magic = (uint32_t)((0x100000000ull + (uint64_t)src.row_len - 1) / (uint64_t)src.row_len);
    if (src.row_len == 1) src.magic = 0u;  // row_len 1 is special-cased in the kernel (2^32 does not fit)
  }
  plan->smem_floats_per_env = off;
  int col = 0, slots = 0, last_slot = -1;
  for (int t = 0; t < plan->n_terms; ++t) {
    catb200_term_t& term = plan->terms[t];
    if (term.op > CATB200_OP_ABS_MINUS_GATE_STILL) return CATB200_ERR_UNSUPPORTED;
    if (term.n_cols == 0 || term.n_ids > CATB200_MAX_IDS || term.src0 >= plan->n_sources)
      return CATB200_ERR_INVALID_ARGUMENT;
    if ((term.src1 != 0xff && term.src1 >= plan->n_sources) || (term.src2 != 0xff && term.src2 >= plan->n_sources))
      return CATB200_ERR_INVALID_ARGUMENT;
    if (col + term.n_cols > CATB200_MAX_COLS) return CATB200_ERR_UNSUPPORTED;
    // per-op sanity: sources that must exist, ids inside the rows they index
    const catb200_source_t& s0 = plan->sources[term.src0];
    const bool contact_op = term.op == CATB200_OP_CONTACT_ANY || term.op == CATB200_OP_N_CONTACT ||
                            term.op == CATB200_OP_FORCE_PEAK_MINUS;
    if (contact_op && s0.aux <= 0) return CATB200_ERR_INVALID_ARGUMENT;
    for (int k = 0; k < term.n_ids; ++k) {
      const int bound = contact_op ? s0.aux : s0.row_len;
      if (term.ids[k] >= bound) return CATB200_ERR_INVALID_ARGUMENT;
    }
    const bool per_id = !(term.op == CATB200_OP_CONTACT_ANY || term.op == CATB200_OP_N_CONTACT ||
                          term.op == CATB200_OP_NORM2_MINUS);
    if (per_id && term.n_ids != term.n_cols) return CATB200_ERR_INVALID_ARGUMENT;
    if (!per_id && term.n_cols != 1) return CATB200_ERR_INVALID_ARGUMENT;
    const bool needs_src1 = term.op == CATB200_OP_ABSDIFF_MINUS || term.op == CATB200_OP_ABSDIFF_MINUS_GATE_Y ||
                            term.op == CATB200_OP_ACTION_RATE || term.op == CATB200_OP_AIR_TIME;
    const bool needs_cmd = term.op == CATB200_OP_ABSDIFF_MINUS_GATE_Y || term.op == CATB200_OP_AIR_TIME ||
                           term.op == CATB200_OP_N_CONTACT || term.op == CATB200_OP_ABS_MINUS_GATE_STILL;
    if (needs_src1 && term.src1 == 0xff) return CATB200_ERR_INVALID_ARGUMENT;
    if (needs_cmd && (term.src2 == 0xff || plan->sources[term.src2].row_len < 3)) return CATB200_ERR_INVALID_ARGUMENT;
    if (term.stat_slot != last_slot) {
      if (term.stat_slot != slots) return CATB200_ERR_INVALID_ARGUMENT;  // slots must be 0,1,2,... in order
      plan->slot_col_begin[slots] = (uint16_t)col;
      last_slot = term.stat_slot;
      ++slots;
    }
    term.col_offset = (uint16_t)col;
    for (int k = 0; k < term.n_cols; ++k) plan->col_term[col + k] = (uint8_t)t;
    col += term.n_cols;
  }
  plan->slot_col_begin[slots] = (uint16_t)col;
  plan->n_cols = col;
  plan->n_slots = slots;
  if ((size_t)off * kTile * sizeof(float) > 200 * 1024) return CATB200_ERR_UNSUPPORTED;
  return CATB200_OK;
}

size_t catb200_cat_workspace_bytes(int32_t num_envs, int32_t n_cols) {
  if (num_envs < 0 || n_cols < 0) return 0;
  return 256 + align256(sizeof(uint32_t) * CATB200_MAX_COLS) + align256(sizeof(float) * (size_t)num_envs * n_cols);
}

static int launch_eval(const catb200_plan_t* plan, const catb200_cat_params_t* prm, int num_envs, float* running_max,
                       int* rm_init, CatWorkspace ws, float* out_rowmajor, int mode, cudaStream_t stream) {
  const size_t smem = (size_t)plan->smem_floats_per_env * kTile * sizeof(float);
  const int grid = (num_envs + kTile - 1) / kTile;
  if (mode == kEvalStep) {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalStep>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalStep><<<grid, kEvalThreads, smem, stream>>>(*plan, *prm, num_envs, running_max, rm_init, ws, nullptr);
  } else {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalRowMajor>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalRowMajor><<<grid, kEvalThreads, smem, stream>>>(*plan, *prm, num_envs, nullptr, nullptr, ws, out_rowmajor);
  }
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_step(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                     float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                     float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf, float* reward_out,
                     float* dones_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !rm_init || !episode_sums || !mean_values || !cstr_prob ||
      !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_cols <= 0 || plan->smem_floats_per_env <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  if (raw_reward && (!reward_out || !dones_out)) return CATB200_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < catb200_cat_workspace_bytes(num_envs, plan->n_cols)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  CatWorkspace ws = carve(workspace, num_envs);
  int rc = launch_eval(plan, params, num_envs, running_max, rm_init, ws, nullptr, kEvalStep, st);
  if (rc != CATB200_OK) return rc;
  const int grid = (num_envs + kApplyThreads - 1) / kApplyThreads;
  cat_apply_kernel<<<grid, kApplyThreads, 0, st>>>(*plan, *params, num_envs, running_max, ws.c_t, episode_sums,
                                                   mean_values, cstr_prob, raw_reward, reset_buf, reward_out,
                                                   dones_out);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_eval_terms(const catb200_plan_t* plan, int32_t num_envs, float* out, void* stream) {
  if (!plan || num_envs <= 0 || !out || plan->n_cols <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  catb200_cat_params_t dummy = {};
  CatWorkspace ws = {};
  return launch_eval(plan, &dummy, num_envs, nullptr, nullptr, ws, out, kEvalRowMajor, as_stream(stream));
}

int catb200_cat_probs(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                      const float* running_max, float* probs_out, const void* workspace, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !probs_out || !workspace) return CATB200_ERR_INVALID_ARGUMENT;
  CatWorkspace ws = carve(const_cast<void*>(workspace), num_envs);
  const int grid = (num_envs + kApplyThreads - 1) / kApplyThreads;
  cat_probs_kernel<<<grid, kApplyThreads, 0, as_stream(stream)>>>(*plan, *params, num_envs, running_max, ws.c_t, probs_out);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_reset_stats(const int64_t* env_ids, int32_t n_ids, const uint8_t* mask, const int64_t* episode_length,
                            int32_t num_envs, int32_t n_slots, float* episode_sums, float* mean_values, float* out,
                            void* stream) {
  if (!episode_length || num_envs <= 0 || n_slots <= 0 || !episode_sums || !mean_values || !out)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (env_ids && n_ids < 0) return CATB200_ERR_INVALID_ARGUMENT;
  cat_reset_kernel<<<n_slots, kResetThreads, 0, as_stream(stream)>>>(env_ids, n_ids, mask, episode_length, num_envs,
                                                                    episode_sums, mean_values, out);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

// Actor-critic MLP forward / backward and the PPO-clip loss on sm_100a.
//
// Reference: Agent (U/cleanrl/ppo.py:71-123) and the minibatch body of PPO() (ppo.py:298-352).
//
// Structure of one minibatch (M rows, both nets batched into every launch: 0 = critic, 1 = actor):
//   gather_kernel        : X[M, obs_pad] <- padded operand rows of the rollout at mb_inds + advantage mean / unbiased std
//   mlp_gemm<fwd> (x3)   : H_l = ELU(H_{l-1} W_l^T + b_l) on tcgen05 (tc_gemm.cu), activations kept for the backward pass
//   head_mma_kernel      : heads (h3 -> act_dim / 1) as mma.sync tf32 fragments, 16 samples per warp (fp32-grade forward:
//                          hi + lo weight terms), Normal log-prob, PPO-clip + clipped value loss + entropy, their
//                          gradients w.r.t. the head weights / biases / log-std (per-CTA partial rows) and
//                          dZ3 = dH3 * ELU'(H3) for both nets.  head_kernel<train> (one warp per sample, fp32 FMAs) is the
//                          older variant (CATB200_HEAD=warp); head_kernel<rollout> serves the rollout policy
//   mlp_wgrad (x3)       : dW_l += dZ_l^T H_{l-1}, db_l += dZ_l^T 1 on tcgen05, red.global.add into padded accumulators
//   mlp_gemm<dgrad> (x2) : dZ_{l-1} = (dZ_l W_l) * ELU'(H_{l-1})
//   opt_step_kernel      : padded weight-gradient accumulators + head partial rows -> flat gradient (reference order),
//                          squared norm -- grid barrier -- clip, Adam, operand copies W / W^T: the optimizer step as ONE
//                          launch (opt_step_peer_kernel: the same with the NVLink gradient exchange in the middle;
//                          fold_grads_kernel + optim.cu's grad_norm_kernel + adam_cast_kernel: the separate launches)
// Operand precision (dims->prec): tf32 = fp32 storage rounded to tf32, the reference's GPU numerics
// (scripts/clean_rl/train.py:86-87); bf16 = half the operand bytes.  Heads, loss and optimizer math are fp32 either way.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mma.cuh"
#include "optim.cuh"
#include "philox.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace catb200 {

// Backward-pass overlap: the weight-gradient GEMM of layer l and the data-gradient GEMM that produces dZ_{l-1} both
// only read dZ_l, so they can run concurrently -- wgrad on a library-owned side stream forked from / joined back into
// the caller's stream with events (inside a CUDA graph capture these become plain graph edges).  All work is still
// complete when the caller's stream reaches the end of the call.  Opt-in (CATB200_BWD_OVERLAP=1).
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t dz_ready[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t done = nullptr;
};
static SideStream* side_stream() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("CATB200_BWD_OVERLAP");
    enabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (!enabled) return nullptr;
  static SideStream per_device[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& s = per_device[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    bool ok = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < 3; ++i) ok = ok && cudaEventCreateWithFlags(&s.dz_ready[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return nullptr;
  }
  return &s;
}

constexpr int kHeadThreads = 256;
constexpr int kMaxAct = 16;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;
// layout of one head-kernel CTA's partial row: [value][lane]; values 0..63 = gW4a[j][f] (j = v/4, f = v%4)
constexpr int kHvW4c = 64, kHvB4a = 68, kHvLogstd = 69, kHvScalars = 70;
constexpr int kHeadValues = 76;
constexpr int kHeadSmem = (kHeadThreads / 32) * kHeadValues * 32 * 4;  // 76 KiB

// Folds the padded weight-gradient accumulators (mlp_wgrad_kernel's red.global.add target, [N, Kpad] with 16-byte
// aligned rows) into the flat gradient (reference parameter order, [N, Ktrue] rows, arbitrary alignment) and zeroes
// them for the next minibatch.  blockIdx.y selects the segment; the last segment is the head kernel's per-CTA rows.
struct FoldArgs {
  float* acc[6];
  float* grad[6];
  int N[6], Kpad[6], Ktrue[6];
  int n_segments;
  // head segment
  const float* head_part; int head_rows;
  float* gW4c; float* gb4c; float* gW4a; float* gb4a; float* glogstd;
  const float* logstd; float* loss_acc;
  int A, h3, M;
  float ent_coef, vf_coef;
};

// One 16-byte quad of a padded accumulator: added into the flat gradient, accumulator zeroed; returns the sum of squares of
// the (up to four) final gradient values.  The gradient values are fetched before the accumulator is zeroed so that both
// loads are in flight together (the pointers may alias as far as the compiler knows).
__device__ __forceinline__ float fold_weight_quad(const FoldArgs& r, int seg, int q) {
  const int Kpad = r.Kpad[seg], Kt = r.Ktrue[seg];
  float4* __restrict__ acc4 = reinterpret_cast<float4*>(r.acc[seg]);
  const int qpr = Kpad / 4;
  const int n = q / qpr, k = (q - n * qpr) * 4;
  const float4 v = acc4[q];
  float* gp = r.grad[seg] + (size_t)n * Kt + k;
  const float vv[4] = {v.x, v.y, v.z, v.w};
  float g[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = k + i < Kt ? gp[i] : 0.0f;
  acc4[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  float ss = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (k + i < Kt) {  // the padded input columns of layer 0 are dropped
      const float t = g[i] + vv[i];
      gp[i] = t;
      ss = fmaf(t, t, ss);
    }
  return ss;
}

// Head segment, one CTA (256 threads) per 32 consecutive values of the head kernel's per-CTA rows: the 8 warps split the
// rows (a single thread walking all ~148 rows is a chain of dependent-latency loads: 20 us per launch), combine through
// shared memory, and warp 0 routes every value to its gradient slot / loss accumulator.  Returns (in warp 0) the square of
// the final gradient value this thread wrote, 0 elsewhere.
__device__ __forceinline__ float fold_head_group(const FoldArgs& r, int group, float (*part_s)[32]) {
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
  const int e = group * 32 + ln;
  float acc = 0.0f;
  for (int c0 = warp; c0 < r.head_rows; c0 += 64) {  // eight independent loads in flight per thread
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = c0 + 8 * u;
      t[u] = c < r.head_rows ? r.head_part[(size_t)c * kHeadValues * 32 + e] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += t[u];
  }
  part_s[warp][ln] = acc;
  __syncthreads();
  if (warp != 0) return 0.0f;
  float t = 0.0f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += part_s[w][ln];
  const int v = e >> 5, lane = e & 31;
  const float inv_M = 1.0f / (float)r.M;
  float fin = 0.0f;
  if (v < kHvW4c) {
    const int j = v >> 2, f = v & 3;
    if (j < r.A) { float* p = r.gW4a + j * r.h3 + lane * 4 + f; fin = *p + t; *p = fin; }
  } else if (v < kHvB4a) {
    float* p = r.gW4c + lane * 4 + (v - kHvW4c); fin = *p + t; *p = fin;
  } else if (v == kHvB4a) {  // per-action-dim scalars live in the even lane of pair (2j, 2j+1)
    if ((lane & 1) == 0 && (lane >> 1) < r.A) { float* p = r.gb4a + (lane >> 1); fin = *p + t; *p = fin; }
  } else if (v == kHvLogstd) {
    // policy part + d(-ent_coef * mean entropy)/d logstd_j = -ent_coef (entropy is sample independent)
    if ((lane & 1) == 0 && (lane >> 1) < r.A) { float* p = r.glogstd + (lane >> 1); fin = *p + (t - r.ent_coef); *p = fin; }
  } else if (lane == 0) {
    const int k = v - kHvScalars;  // g_b4c, pg, v, kl, clip, old_kl
    if (k == 0) {
      fin = r.gb4c[0] + t; r.gb4c[0] = fin;
    } else if (k == 1) {
      r.loss_acc[0] += t * inv_M;
      // entropy = sum_j (0.5 + 0.5 log(2 pi) + logstd_j); also the per-minibatch counter and total loss
      float ent = 0.0f;
      for (int j = 0; j < r.A; ++j) ent += 0.5f + kLogSqrt2Pi + r.logstd[j];
      r.loss_acc[2] += ent;
      atomicAdd(r.loss_acc + 6, t * inv_M - r.ent_coef * ent);
      r.loss_acc[7] += 1.0f;
    } else if (k == 2) {
      r.loss_acc[1] += t * inv_M;
      atomicAdd(r.loss_acc + 6, r.vf_coef * t * inv_M);
    } else if (k == 3) {
      r.loss_acc[3] += t * inv_M;
    } else if (k == 4) {
      r.loss_acc[4] += t * inv_M;
    } else if (k == 5) {
      r.loss_acc[5] += t * inv_M;
    }
  }
  return fin * fin;
}

__global__ void __launch_bounds__(256) fold_grads_kernel(const __grid_constant__ FoldArgs r) {
  pdl_launch_dependents();
  pdl_wait();
  const int seg = blockIdx.y;
  if (seg < r.n_segments) {
    const int quads = r.N[seg] * r.Kpad[seg] / 4;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) fold_weight_quad(r, seg, q);
    return;
  }
  __shared__ float part_s[8][32];
  if ((int)blockIdx.x < kHeadValues) fold_head_group(r, blockIdx.x, part_s);
}

// ---- minibatch gather + advantage statistics -----------------------------------------------------------
struct MbStats {
  float adv_mean, adv_std;  // unbiased std (ppo.py:316-318)
  unsigned int ticket, pad;
  double sum, sumsq;
};

// obs_all / X: operand rows of `row_bytes` bytes (obs_pad bf16 or fp32 elements, a multiple of 16 bytes)
__global__ void __launch_bounds__(256)
gather_kernel(const int64_t* __restrict__ mb_inds, int M, const uint4* __restrict__ obs_all, int row_bytes,
              const float* __restrict__ adv_all, const float* __restrict__ logp_all, const float* __restrict__ ret_all,
              const float* __restrict__ val_all, const float* __restrict__ act_all, int A, uint4* __restrict__ X,
              float4* __restrict__ scal_mb, float* __restrict__ act_mb, MbStats* __restrict__ st) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = row_bytes / 16;  // 16-byte chunks per row
  const int total = M * chunks;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int m = e / chunks, c = e - m * chunks;
    const int64_t src = mb_inds[m];
    X[e] = __ldg(obs_all + (size_t)src * chunks + c);
  }
  // per-sample scalars {old log-prob, advantage, return, old value} and actions, packed contiguously in
  // minibatch order so that the head kernel streams them instead of chasing mb_inds
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < M * A; e += gridDim.x * blockDim.x) {
    const int m = e / A, j = e - m * A;
    act_mb[e] = __ldg(act_all + (size_t)mb_inds[m] * A + j);
  }
  double s = 0.0, q = 0.0;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
    const int64_t src = mb_inds[m];
    const float adv = __ldg(adv_all + src);
    scal_mb[m] = make_float4(__ldg(logp_all + src), adv, __ldg(ret_all + src), __ldg(val_all + src));
    const double a = (double)adv;
    s += a;
    q += a * a;
  }
  // Only the CTAs that own a sample in the loop above have anything to add; the others (15 of 16 at M = 16384) skip the
  // reduction and the ticket: ~3000 fewer same-address atomics queueing in the L2 per launch.
  const int stat_ctas = min((M + (int)blockDim.x - 1) / (int)blockDim.x, (int)gridDim.x);
  if ((int)blockIdx.x >= stat_ctas) return;
  s = warp_sum(s);
  q = warp_sum(q);
  __shared__ double sh_s[8], sh_q[8];
  if ((threadIdx.x & 31) == 0) {
    sh_s[threadIdx.x >> 5] = s;
    sh_q[threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < 8; ++w) {
      ts += sh_s[w];
      tq += sh_q[w];
    }
    atomicAdd(&st->sum, ts);
    atomicAdd(&st->sumsq, tq);
  }
  if (last_block_ticket(&st->ticket, stat_ctas)) {
    if (threadIdx.x == 0) {
      const double ts = __longlong_as_double(atomicExch((unsigned long long*)&st->sum, 0ull));
      const double tq = __longlong_as_double(atomicExch((unsigned long long*)&st->sumsq, 0ull));
      const double mean = ts / M;
      const double var = M > 1 ? fmax((tq - ts * mean) / (double)(M - 1), 0.0) : 0.0;
      st->adv_mean = (float)mean;
      st->adv_std = (float)sqrt(var);
    }
  }
}

// ---- heads, loss and their gradients -----------------------------------------------------------------

struct HeadArgs {
  const void* H3[2];   // [M, h3] activations of the last hidden layer (0 critic, 1 actor), bf16 or fp32
  void* dZ3[2];        // [M, h3] out (training): gradient w.r.t. the pre-activation of that layer
  const float* W4c; const float* b4c;  // critic head [1, h3], [1]
  const float* W4a; const float* b4a;  // actor head  [A, h3], [A]
  const float* logstd;                 // [A]
  int M, h3, A;
  // rollout outputs / inputs
  const float* noise; const float* action_in; float* action; float* logprob; float* value; float* mean_out;
  // rollout: device-side Normal.sample() noise (Philox4x32-10 + Box-Muller) when rng_state != NULL; the counter
  // (rng_state[1]) is advanced by the host-side wrapper's trailing bump kernel
  const unsigned long long* rng_state;
  // training inputs, packed in minibatch order by gather_kernel
  const float4* scal_mb;  // {old log-prob, advantage, return, old value}
  const float* act_mb;    // [M, A]
  const float* norm_stats; const MbStats* mb;
  catb200_ppo_hparams_t hp;
  // training outputs
  float* head_part;  // [gridDim.x][kHeadValues][32] per-CTA partial sums (training)
  int reverse;       // walk the samples from the last to the first (row order of the minibatch launches, tc_gemm.cu)
};

template <int PREC>
__device__ __forceinline__ void load_row4(const void* base, size_t row, int h3, int lane, float (&h)[4]);
template <>
__device__ __forceinline__ void load_row4<kPrecBf16>(const void* base, size_t row, int h3, int lane, float (&h)[4]) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(static_cast<const bf16*>(base) + row * h3) + lane);
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
  h[0] = __low2float(p[0]); h[1] = __high2float(p[0]); h[2] = __low2float(p[1]); h[3] = __high2float(p[1]);
}
template <>
__device__ __forceinline__ void load_row4<kPrecTf32>(const void* base, size_t row, int h3, int lane, float (&h)[4]) {
  const float4 r = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(base) + row * h3) + lane);
  h[0] = r.x; h[1] = r.y; h[2] = r.z; h[3] = r.w;
}
template <int PREC>
__device__ __forceinline__ void store_row4(void* base, size_t row, int h3, int lane, const float (&d)[4]) {
  if (PREC == kPrecBf16) {
    uint2 o;
    o.x = pack_bf16x2(d[0], d[1]); o.y = pack_bf16x2(d[2], d[3]);
    reinterpret_cast<uint2*>(static_cast<bf16*>(base) + row * h3)[lane] = o;
  } else {  // dZ3 is a tensor-core operand of dgrad / wgrad: stored rounded to tf32
    reinterpret_cast<float4*>(static_cast<float*>(base) + row * h3)[lane] =
        make_float4(round_tf32(d[0]), round_tf32(d[1]), round_tf32(d[2]), round_tf32(d[3]));
  }
}

// One warp per sample; lane owns features lane*4 .. lane*4+3 of the 128-wide h3.
template <bool TRAIN, int PREC>
__global__ void __launch_bounds__(kHeadThreads)
head_kernel(const __grid_constant__ HeadArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int AP = kMaxAct;  // action dims carried through the unrolled loops (weights beyond A are zero)
  constexpr int F = 4;  // features per lane (h3 == 128)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = kHeadThreads / 32;
  const int A = a.A;
  float w4a[kMaxAct][F], w4c[F];
#pragma unroll
  for (int f = 0; f < F; ++f) w4c[f] = __ldg(a.W4c + lane * F + f);
#pragma unroll
  for (int j = 0; j < kMaxAct; ++j)
#pragma unroll
    for (int f = 0; f < F; ++f) w4a[j][f] = j < A ? __ldg(a.W4a + j * a.h3 + lane * F + f) : 0.0f;
  const float b4c = __ldg(a.b4c);
  // Action dim j lives in the lane pair (2j, 2j+1): that is where the recursive-halving reduction of the 16
  // head dot products leaves its sums.  Only the even lane of a pair is "owner" (contributes to sums / stores).
  const int aj = lane >> 1;
  const bool owner = (lane & 1) == 0 && aj < A;
  float my_b4a = 0.0f, my_logstd = 0.0f;
  if (aj < A) {
    my_b4a = __ldg(a.b4a + aj);
    my_logstd = __ldg(a.logstd + aj);
  }
  const float my_std = expf(my_logstd);
  const float my_inv_var = 1.0f / (my_std * my_std);

  // training accumulators (per lane): head weight grads for its features, per-action scalars in lane j
  float gw4a[kMaxAct][F], gw4c[F];
  float g_b4a = 0.0f, g_logstd = 0.0f, g_b4c = 0.0f;
  float l_pg = 0.0f, l_v = 0.0f, l_kl = 0.0f, l_clip = 0.0f, l_oldkl = 0.0f;
  if (TRAIN) {
#pragma unroll
    for (int f = 0; f < F; ++f) {
      gw4c[f] = 0.0f;
#pragma unroll
      for (int j = 0; j < kMaxAct; ++j) gw4a[j][f] = 0.0f;
    }
  }
  float adv_mean = 0.0f, adv_std = 1.0f, m1 = 0.0f, v1 = 1.0f, m2 = 0.0f, v2 = 1.0f;
  if (TRAIN) {
    adv_mean = a.mb->adv_mean;
    adv_std = a.mb->adv_std;
    m1 = a.norm_stats[0]; v1 = a.norm_stats[1]; m2 = a.norm_stats[2]; v2 = a.norm_stats[3];
  }
  const float inv_sd1 = 1.0f / sqrtf(v1 + 1e-8f), inv_sd2 = 1.0f / sqrtf(v2 + 1e-8f);
  const float inv_M = 1.0f / (float)a.M;
  unsigned long long rng_seed = 0, rng_offset = 0;
  if (!TRAIN && a.rng_state) {
    rng_seed = a.rng_state[0];
    rng_offset = a.rng_state[1];
  }

  // software pipeline: the inputs of the next sample are in flight while the current one is processed
  const int m_stride = gridDim.x * warps;
  int m = blockIdx.x * warps + warp;
  float n_hc[F] = {0.f, 0.f, 0.f, 0.f}, n_ha[F] = {0.f, 0.f, 0.f, 0.f};
  float4 n_sc = make_float4(0.f, 0.f, 0.f, 0.f);
  float n_act = 0.0f;
  auto fetch = [&](int mi) {
    const int mm = a.reverse ? a.M - 1 - mi : mi;
    load_row4<PREC>(a.H3[0], (size_t)mm, a.h3, lane, n_hc);
    load_row4<PREC>(a.H3[1], (size_t)mm, a.h3, lane, n_ha);
    if (TRAIN) {
      n_sc = __ldg(a.scal_mb + mm);
      if (owner) n_act = __ldg(a.act_mb + (size_t)mm * A + aj);
    }
  };
  if (m < a.M) fetch(m);
  for (int mi = m; mi < a.M; mi += m_stride) {
    m = a.reverse ? a.M - 1 - mi : mi;  // the sample this iteration works on
    // ---- the two 128-wide activation rows (coalesced) + scalars of this sample
    float hc[F], ha[F];
    const float4 sc = n_sc;
    const float act_in = n_act;
#pragma unroll
    for (int f = 0; f < F; ++f) {
      hc[f] = n_hc[f];
      ha[f] = n_ha[f];
    }
    if (mi + m_stride < a.M) fetch(mi + m_stride);
    // ---- heads: value and action mean (fp32), warp all-reduce of the per-lane partial dot products
    float v = 0.0f;
#pragma unroll
    for (int f = 0; f < F; ++f) v = fmaf(hc[f], w4c[f], v);
    v = warp_sum(v) + b4c;
    float mean_j = 0.0f;  // lane j keeps mean_j
    {
      float p[AP];
#pragma unroll
      for (int j = 0; j < AP; ++j) {
        p[j] = 0.0f;
#pragma unroll
        for (int f = 0; f < F; ++f) p[j] = fmaf(ha[f], w4a[j][f], p[j]);
      }
      // recursive halving: 8 + 4 + 2 + 1 exchanges leave, in lane l, the 16-lane partial sum of dim (l >> 1);
      // one more exchange with the pair partner completes the 32-lane sum (16 shuffles instead of 80)
#pragma unroll
      for (int o = 16; o >= 2; o >>= 1) {
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int j = 0; j < o / 2; ++j) {
          const float send = upper ? p[j] : p[j + o / 2];
          const float keep = upper ? p[j + o / 2] : p[j];
          p[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      mean_j = p[0] + __shfl_xor_sync(0xffffffffu, p[0], 1) + my_b4a;
    }
    if (!TRAIN) {
      // ---- rollout: sample, log-prob, store (ppo.py:104-119)
      float act = 0.0f, lp = 0.0f;
      if (owner) {
        if (a.action_in) {
          act = __ldg(a.action_in + (size_t)m * A + aj);  // evaluate a given action (ppo.py:110 `action is not None`)
        } else {
          float eps = 0.0f;
          if (a.rng_state) eps = philox_normal(rng_seed, rng_offset, (unsigned long long)m * A + aj);
          else if (a.noise) eps = __ldg(a.noise + (size_t)m * A + aj);
          act = fmaf(my_std, eps, mean_j);
        }
        const float d = act - mean_j;
        lp = -(d * d) * 0.5f * my_inv_var - my_logstd - kLogSqrt2Pi;
        if (a.action) a.action[(size_t)m * A + aj] = act;
        if (a.mean_out) a.mean_out[(size_t)m * A + aj] = mean_j;
      }
      lp = warp_sum(lp);
      if (lane == 0) {
        if (a.logprob) a.logprob[m] = lp;
        if (a.value) a.value[m] = v;
      }
      continue;
    }
    // ---- training: PPO-clip loss and its gradient for this sample (ppo.py:300-344)
    float lp = 0.0f, dmu = 0.0f, dls = 0.0f;  // owner lane of dim j: d logp / d mean_j, d logp / d logstd_j
    if (owner) {
      const float d = act_in - mean_j;
      lp = -(d * d) * 0.5f * my_inv_var - my_logstd - kLogSqrt2Pi;
      dmu = d * my_inv_var;
      dls = d * d * my_inv_var - 1.0f;
    }
    const float newlogp = warp_sum(lp);
    const float logratio = newlogp - sc.x;
    const float ratio = expf(logratio);
    float adv = sc.y;
    if (a.hp.norm_adv) adv = (adv - adv_mean) / (adv_std + 1e-8f);
    const float clipped = fminf(fmaxf(ratio, 1.0f - a.hp.clip_coef), 1.0f + a.hp.clip_coef);
    const float pg1 = -adv * ratio, pg2 = -adv * clipped;
    const float pg = fmaxf(pg1, pg2);
    // d max(pg1, pg2) / d ratio: -adv through pg1 when it is the larger (or tied, unclipped) branch
    float dpg_dratio;
    if (pg1 > pg2) dpg_dratio = -adv;
    else if (pg1 < pg2) dpg_dratio = (clipped == ratio) ? -adv : 0.0f;
    else dpg_dratio = (clipped == ratio) ? -adv : -0.5f * adv;
    const float dL_dlogp = dpg_dratio * ratio * inv_M;
    // value loss on normalised values (ppo.py:328-341)
    const float nv = (v - m2) * inv_sd2;  // value_rms(newvalue, update=False): statistics after both updates
    const float ret_n = (sc.z - m2) * inv_sd2;
    const float val_n = (sc.w - m1) * inv_sd1;
    float vl, dvl_dnv;
    const float e_u = nv - ret_n;
    if (a.hp.clip_vloss) {
      const float diff = nv - val_n;
      const float dclip = fminf(fmaxf(diff, -a.hp.clip_coef), a.hp.clip_coef);
      const float e_c = val_n + dclip - ret_n;
      const float lu = e_u * e_u, lc = e_c * e_c;
      const float pass = (dclip == diff) ? 1.0f : 0.0f;
      if (lu > lc) { vl = lu; dvl_dnv = 2.0f * e_u; }
      else if (lu < lc) { vl = lc; dvl_dnv = 2.0f * e_c * pass; }
      else { vl = lu; dvl_dnv = e_u + e_c * pass; }
    } else {
      vl = e_u * e_u;
      dvl_dnv = 2.0f * e_u;
    }
    const float dL_dv = a.hp.vf_coef * 0.5f * dvl_dnv * inv_sd2 * inv_M;
    l_pg += pg; l_v += 0.5f * vl; l_kl += (ratio - 1.0f) - logratio; l_oldkl += -logratio;
    l_clip += fabsf(ratio - 1.0f) > a.hp.clip_coef ? 1.0f : 0.0f;
    // gradients w.r.t. the head outputs: lane j holds dL/dmean_j
    const float dmean = dL_dlogp * dmu;
    g_b4a += dmean;
    g_logstd += dL_dlogp * dls;  // entropy term added once at the end (it does not depend on the sample)
    g_b4c += dL_dv;
    // back through the heads: dH3 = dmean . W4a (actor), dv * W4c (critic); times ELU' -> dZ3
    float dha[F] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int j = 0; j < kMaxAct; ++j) {
      if (j < A) {
        const float dmj = __shfl_sync(0xffffffffu, dmean, 2 * j);
#pragma unroll
        for (int f = 0; f < F; ++f) {
          dha[f] = fmaf(dmj, w4a[j][f], dha[f]);
          gw4a[j][f] = fmaf(dmj, ha[f], gw4a[j][f]);
        }
      }
    }
    float dza[F], dzc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      dza[f] = dha[f] * elu_grad_from_output(ha[f]);
      dzc[f] = dL_dv * w4c[f] * elu_grad_from_output(hc[f]);
      gw4c[f] = fmaf(dL_dv, hc[f], gw4c[f]);
    }
    // the bias gradient of the last hidden layer (column sums of dZ3) comes out of mlp_wgrad_kernel's ones-MMA
    store_row4<PREC>(a.dZ3[1], (size_t)m, a.h3, lane, dza);
    store_row4<PREC>(a.dZ3[0], (size_t)m, a.h3, lane, dzc);
  }
  if (!TRAIN) return;

  // ---- CTA-level reduction: every warp parks its accumulators in shared memory ([value][lane] rows),
  // one barrier, then the CTA sums over its warps and writes ONE partial row per CTA (plain coalesced
  // stores, no atomics).  fold_grads_kernel folds the rows into the gradient.
  extern __shared__ float hsm[];  // [warps][kHeadValues][32]
  float* mine = hsm + (size_t)warp * kHeadValues * 32;
#pragma unroll
  for (int j = 0; j < kMaxAct; ++j)
#pragma unroll
    for (int f = 0; f < F; ++f) mine[(j * F + f) * 32 + lane] = gw4a[j][f];
#pragma unroll
  for (int f = 0; f < F; ++f) mine[(kHvW4c + f) * 32 + lane] = gw4c[f];
  mine[kHvB4a * 32 + lane] = g_b4a;        // lane j: d/d b4a[j]
  mine[kHvLogstd * 32 + lane] = g_logstd;  // lane j: d/d logstd[j] (policy part)
  // warp-uniform scalars: keep lane 0's copy only
  const float scal[6] = {g_b4c, l_pg, l_v, l_kl, l_clip, l_oldkl};
#pragma unroll
  for (int k = 0; k < 6; ++k) mine[(kHvScalars + k) * 32 + lane] = lane == 0 ? scal[k] : 0.0f;
  __syncthreads();
  float* __restrict__ row = a.head_part + (size_t)blockIdx.x * kHeadValues * 32;
  for (int o = threadIdx.x; o < kHeadValues * 32; o += kHeadThreads) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kHeadThreads / 32; ++w) t += hsm[(size_t)w * kHeadValues * 32 + o];
    row[o] = t;
  }
}

// ---- heads + loss on warp-level tensor cores (training) ------------------------------------------------
// head_kernel<true> above gives every sample a whole warp: 594 warp instructions per sample, two warps per scheduler,
// issue bound (ncu: 22 us for 16384 samples).  head_mma_kernel gives a warp 16 samples at a time and runs the three
// small matrix products of the heads as mma.sync.m16n8k8 tf32 fragments out of a shared-memory copy of the tile:
//   mean[16, A]   = Ha[16, 128] W4a^T      forward: the weights are split hi + lo (two tf32 terms), the activations
//   value[16]     = Hc[16, 128] w4c^T      are tf32-exact already -> fp32-grade head outputs, as in head_kernel
//   dHa[16, 128]  = dmean[16, A] W4a       (tf32 operands, like every other backward GEMM of the step)
//   gW4a[A, 128] += dmean^T[A, 16] Ha[16, 128]
// The accumulator fragment of one product is the A fragment of the next: mma.sync leaves C[row g][cols 2t, 2t+1] in
// thread (g = lane / 4, t = lane % 4) and wants A[row g][k-slots t, t + 4]; the reduction index of the consumer is
// free to be permuted, so slot t is bound to column 2t and slot t + 4 to column 2t + 1 and the B fragments are fetched
// through the same permutation -- no shuffles, no shared-memory round trip (only gW4a needs dmean transposed).
// Loss arithmetic per sample is head_kernel's, executed redundantly by the four threads that share a row.
// Output format (per-CTA partial rows for fold_grads / opt_step) is head_kernel's.
constexpr int kHmWarps = 8;
constexpr int kHmStride = 132;                       // floats per tile row: conflict-free fragment loads (132 % 32 == 4)
constexpr int kHmTile = 16 * kHmStride;              // floats per 16 x 128 tile
constexpr int kHmTrans = 16 * 17;                    // dmean^T staging, per warp
constexpr int kHmWeights = 2 * kHmTile + 3 * 128;    // W4a hi, lo; w4c hi, lo, full
constexpr int kHmPerWarp = 2 * kHmTile + kHmTrans;   // Ha, Hc, dmean^T
constexpr int kHmSmem = (kHmWeights + kHmWarps * kHmPerWarp) * 4;
constexpr int kHmPark = 33;                          // row stride of the parked per-warp values: conflict-free scatter
static_assert(kHmWarps * kHmPerWarp >= kHmWarps * kHeadValues * kHmPark, "the final per-warp rows alias the tiles");

__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int PREC>
__device__ __forceinline__ void store_pair(void* base, size_t row, int h3, int col, float x, float y) {
  if (PREC == kPrecBf16) {
    *reinterpret_cast<uint32_t*>(static_cast<bf16*>(base) + row * h3 + col) = pack_bf16x2(x, y);
  } else {  // tensor-core operand of dgrad / wgrad: stored rounded to tf32
    *reinterpret_cast<float2*>(static_cast<float*>(base) + row * h3 + col) = make_float2(round_tf32(x), round_tf32(y));
  }
}

template <int PREC>
__global__ void __launch_bounds__(kHmWarps * 32, 1)
head_mma_kernel(const __grid_constant__ HeadArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float hm[];
  float* wa_hi = hm;
  float* wa_lo = hm + kHmTile;
  float* wc_hi = hm + 2 * kHmTile;
  float* wc_lo = wc_hi + 128;
  float* wc_full = wc_lo + 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* tiles = hm + kHmWeights + warp * kHmPerWarp;
  float* sHa = tiles;
  float* sHc = tiles + kHmTile;
  float* sT = tiles + 2 * kHmTile;
  const int A = a.A, M = a.M;

  // The activation tiles of a pass travel global -> shared memory: with fp32 storage as 16-byte cp.async copies (no
  // registers, one wait for all 32 of a lane; rows beyond M are zero-filled), with bf16 storage through registers.
  const int passes = (M + 15) / 16;
  const int pass_stride = gridDim.x * kHmWarps;
  auto issue_tiles = [&](int pi) {
    const int row0 = (a.reverse ? passes - 1 - pi : pi) * 16;
    if (PREC == kPrecTf32) {
      const float* ga = static_cast<const float*>(a.H3[1]);
      const float* gc = static_cast<const float*>(a.H3[0]);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int row = row0 + r;
        const bool ok = row < M;
        const size_t off = (size_t)(ok ? row : 0) * 128 + lane * 4;
        cp_async16(smem_u32(sHa + r * kHmStride + lane * 4), ga + off, ok);
        cp_async16(smem_u32(sHc + r * kHmStride + lane * 4), gc + off, ok);
      }
      cp_async_commit();
    } else {
#pragma unroll
      for (int r0 = 0; r0 < 16; r0 += 4) {
        float ha[4][4], hc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = row0 + r0 + r;
          if (row < M) {
            load_row4<PREC>(a.H3[1], (size_t)row, 128, lane, ha[r]);
            load_row4<PREC>(a.H3[0], (size_t)row, 128, lane, hc[r]);
          } else {
#pragma unroll
            for (int f = 0; f < 4; ++f) ha[r][f] = hc[r][f] = 0.0f;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          *reinterpret_cast<float4*>(sHa + (r0 + r) * kHmStride + lane * 4) = make_float4(ha[r][0], ha[r][1], ha[r][2], ha[r][3]);
          *reinterpret_cast<float4*>(sHc + (r0 + r) * kHmStride + lane * 4) = make_float4(hc[r][0], hc[r][1], hc[r][2], hc[r][3]);
        }
      }
    }
  };
  const int first_pass = blockIdx.x + gridDim.x * warp;
  if (first_pass < passes) issue_tiles(first_pass);  // in flight while the weights are split below

  // this thread's four action slots: js[n][e] = 8n + 2t + e (the C-fragment columns of n-tile n); the scalar loads are
  // issued here so that they are in flight, like the tiles, while the weights are split
  float s_b4a[2][2], s_logstd[2][2], s_inv_var[2][2];
  bool s_ok[2][2];
#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 8 * n + 2 * t + e;
      s_ok[n][e] = j < A;
      s_b4a[n][e] = s_ok[n][e] ? __ldg(a.b4a + j) : 0.0f;
      s_logstd[n][e] = s_ok[n][e] ? __ldg(a.logstd + j) : 0.0f;
    }
  const float b4c = __ldg(a.b4c);
  const float adv_mean = a.mb->adv_mean, adv_std = a.mb->adv_std;
  const float m1 = a.norm_stats[0], v1 = a.norm_stats[1], m2 = a.norm_stats[2], v2 = a.norm_stats[3];

  // head weights -> shared memory, split into two tf32 terms (rows beyond A are zero)
  for (int i = threadIdx.x; i < 16 * 128; i += blockDim.x) {
    const int j = i >> 7, f = i & 127;
    const float w = j < A ? __ldg(a.W4a + j * 128 + f) : 0.0f;
    const float hi = round_tf32(w);
    wa_hi[j * kHmStride + f] = hi;
    wa_lo[j * kHmStride + f] = round_tf32(w - hi);
  }
  for (int f = threadIdx.x; f < 128; f += blockDim.x) {
    const float w = __ldg(a.W4c + f);
    const float hi = round_tf32(w);
    wc_full[f] = w;
    wc_hi[f] = hi;
    wc_lo[f] = round_tf32(w - hi);
  }
  __syncthreads();

#pragma unroll
  for (int n = 0; n < 2; ++n)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float sd = expf(s_logstd[n][e]);
      s_inv_var[n][e] = 1.0f / (sd * sd);
    }
  const float inv_sd1 = 1.0f / sqrtf(v1 + 1e-8f), inv_sd2 = 1.0f / sqrtf(v2 + 1e-8f);
  const float inv_M = 1.0f / (float)M;

  float gwa[16][4];  // gW4a fragments: [n-tile of 8 features][(action g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1)]
  float gwc[16][2];  // gW4c partial over this thread's two rows: features 8nn + 2t + e
#pragma unroll
  for (int nn = 0; nn < 16; ++nn) {
    gwa[nn][0] = gwa[nn][1] = gwa[nn][2] = gwa[nn][3] = 0.0f;
    gwc[nn][0] = gwc[nn][1] = 0.0f;
  }
  float g_b4a[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, g_logstd[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float g_b4c = 0.0f, l_pg = 0.0f, l_v = 0.0f, l_kl = 0.0f, l_clip = 0.0f, l_oldkl = 0.0f;

  for (int pi = first_pass; pi < passes; pi += pass_stride) {
    const int p = a.reverse ? passes - 1 - pi : pi;
    const int row0 = p * 16;
    if (pi != first_pass) {
      __syncwarp();  // every lane is done reading the previous pass's tiles
      issue_tiles(pi);
    }
    // per-sample scalars of this thread's two rows (the four threads of a quad read the same values)
    const int rowA = row0 + g, rowB = row0 + g + 8;
    const bool okA = rowA < M, okB = rowB < M;
    const float4 scA = okA ? __ldg(a.scal_mb + rowA) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 scB = okB ? __ldg(a.scal_mb + rowB) : make_float4(0.f, 0.f, 0.f, 0.f);
    float act[2][2][2];  // [row A / B][n][e]
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * n + 2 * t + e;
        act[0][n][e] = (okA && s_ok[n][e]) ? __ldg(a.act_mb + (size_t)rowA * A + j) : 0.0f;
        act[1][n][e] = (okB && s_ok[n][e]) ? __ldg(a.act_mb + (size_t)rowB * A + j) : 0.0f;
      }
    if (PREC == kPrecTf32) cp_async_wait<0>();
    __syncwarp();

    // ---- forward: action means (two n-tiles of 8 actions) and the value (column 0 of a third n-tile)
    float cm[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, cv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int c0 = 8 * k + t;
      uint32_t fa[4], fc[4];
      fa[0] = __float_as_uint(sHa[g * kHmStride + c0]);
      fa[1] = __float_as_uint(sHa[(g + 8) * kHmStride + c0]);
      fa[2] = __float_as_uint(sHa[g * kHmStride + c0 + 4]);
      fa[3] = __float_as_uint(sHa[(g + 8) * kHmStride + c0 + 4]);
      fc[0] = __float_as_uint(sHc[g * kHmStride + c0]);
      fc[1] = __float_as_uint(sHc[(g + 8) * kHmStride + c0]);
      fc[2] = __float_as_uint(sHc[g * kHmStride + c0 + 4]);
      fc[3] = __float_as_uint(sHc[(g + 8) * kHmStride + c0 + 4]);
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const float* wh = wa_hi + (8 * n + g) * kHmStride + c0;
        const float* wl = wa_lo + (8 * n + g) * kHmStride + c0;
        mma_tf32_1688(cm[n], fa, __float_as_uint(wh[0]), __float_as_uint(wh[4]));
        mma_tf32_1688(cm[n], fa, __float_as_uint(wl[0]), __float_as_uint(wl[4]));
      }
      const uint32_t vh0 = g == 0 ? __float_as_uint(wc_hi[c0]) : 0u, vh1 = g == 0 ? __float_as_uint(wc_hi[c0 + 4]) : 0u;
      const uint32_t vl0 = g == 0 ? __float_as_uint(wc_lo[c0]) : 0u, vl1 = g == 0 ? __float_as_uint(wc_lo[c0 + 4]) : 0u;
      mma_tf32_1688(cv, fc, vh0, vh1);
      mma_tf32_1688(cv, fc, vl0, vl1);
    }
    // value of row g / g + 8 sits in column 0 = thread t == 0 of the quad
    const float valA = __shfl_sync(0xffffffffu, cv[0], lane & ~3) + b4c;
    const float valB = __shfl_sync(0xffffffffu, cv[2], lane & ~3) + b4c;

    // ---- PPO-clip loss and its gradient (ppo.py:300-344), rows A and B
    float dm[2][4];  // dL / d mean in C-fragment layout: [n][(row A, 2t) (row A, 2t+1) (row B, 2t) (row B, 2t+1)]
    float dvv[2];    // dL / d value of rows A, B
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const bool ok = rr ? okB : okA;
      const float4 sc = rr ? scB : scA;
      const float v = rr ? valB : valA;
      float lp = 0.0f, dmu[2][2], dls[2][2];
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float mean_j = cm[n][2 * rr + e] + s_b4a[n][e];
          const float d = act[rr][n][e] - mean_j;
          const float term = -(d * d) * 0.5f * s_inv_var[n][e] - s_logstd[n][e] - kLogSqrt2Pi;
          lp += s_ok[n][e] ? term : 0.0f;
          dmu[n][e] = s_ok[n][e] ? d * s_inv_var[n][e] : 0.0f;
          dls[n][e] = s_ok[n][e] ? d * d * s_inv_var[n][e] - 1.0f : 0.0f;
        }
      lp += __shfl_xor_sync(0xffffffffu, lp, 1);
      lp += __shfl_xor_sync(0xffffffffu, lp, 2);
      const float logratio = lp - sc.x;
      const float ratio = expf(logratio);
      float adv = sc.y;
      if (a.hp.norm_adv) adv = (adv - adv_mean) / (adv_std + 1e-8f);
      const float clipped = fminf(fmaxf(ratio, 1.0f - a.hp.clip_coef), 1.0f + a.hp.clip_coef);
      const float pg1 = -adv * ratio, pg2 = -adv * clipped;
      const float pg = fmaxf(pg1, pg2);
      float dpg_dratio;
      if (pg1 > pg2) dpg_dratio = -adv;
      else if (pg1 < pg2) dpg_dratio = (clipped == ratio) ? -adv : 0.0f;
      else dpg_dratio = (clipped == ratio) ? -adv : -0.5f * adv;
      const float dL_dlogp = ok ? dpg_dratio * ratio * inv_M : 0.0f;
      const float nv = (v - m2) * inv_sd2;
      const float ret_n = (sc.z - m2) * inv_sd2;
      const float val_n = (sc.w - m1) * inv_sd1;
      float vl, dvl_dnv;
      const float e_u = nv - ret_n;
      if (a.hp.clip_vloss) {
        const float diff = nv - val_n;
        const float dclip = fminf(fmaxf(diff, -a.hp.clip_coef), a.hp.clip_coef);
        const float e_c = val_n + dclip - ret_n;
        const float lu = e_u * e_u, lc = e_c * e_c;
        const float pass = (dclip == diff) ? 1.0f : 0.0f;
        if (lu > lc) { vl = lu; dvl_dnv = 2.0f * e_u; }
        else if (lu < lc) { vl = lc; dvl_dnv = 2.0f * e_c * pass; }
        else { vl = lu; dvl_dnv = e_u + e_c * pass; }
      } else {
        vl = e_u * e_u;
        dvl_dnv = 2.0f * e_u;
      }
      const float dL_dv = ok ? a.hp.vf_coef * 0.5f * dvl_dnv * inv_sd2 * inv_M : 0.0f;
      dvv[rr] = dL_dv;
      if (ok && t == 0) {  // one thread of the quad keeps the per-sample scalars
        l_pg += pg; l_v += 0.5f * vl; l_kl += (ratio - 1.0f) - logratio; l_oldkl += -logratio;
        l_clip += fabsf(ratio - 1.0f) > a.hp.clip_coef ? 1.0f : 0.0f;
        g_b4c += dL_dv;
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float dmean = dL_dlogp * dmu[n][e];
          dm[n][2 * rr + e] = round_tf32(dmean);
          g_b4a[n][e] += dmean;
          g_logstd[n][e] += dL_dlogp * dls[n][e];
        }
    }
    // dmean^T for the weight-gradient product: T[sample][action]
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sT[g * 17 + 8 * n + 2 * t + e] = dm[n][e];
        sT[(g + 8) * 17 + 8 * n + 2 * t + e] = dm[n][2 + e];
      }
    __syncwarp();

    // ---- backward through the heads, 8 features (one n-tile) at a time
    // A fragments of dHa = dmean W4a: k-step s covers actions 8s .. 8s+7, slot t = action 8s + 2t, slot t + 4 = 8s + 2t + 1
    uint32_t fdm[2][4];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      fdm[s2][0] = __float_as_uint(dm[s2][0]);
      fdm[s2][1] = __float_as_uint(dm[s2][2]);
      fdm[s2][2] = __float_as_uint(dm[s2][1]);
      fdm[s2][3] = __float_as_uint(dm[s2][3]);
    }
    // A fragments of gW4a = dmean^T Ha: rows = actions, k-step s covers samples 8s .. 8s+7 under the same slot permutation
    uint32_t fdt[2][4];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      fdt[s2][0] = __float_as_uint(sT[(8 * s2 + 2 * t) * 17 + g]);
      fdt[s2][1] = __float_as_uint(sT[(8 * s2 + 2 * t) * 17 + g + 8]);
      fdt[s2][2] = __float_as_uint(sT[(8 * s2 + 2 * t + 1) * 17 + g]);
      fdt[s2][3] = __float_as_uint(sT[(8 * s2 + 2 * t + 1) * 17 + g + 8]);
    }
#pragma unroll
    for (int nn = 0; nn < 16; ++nn) {
      const int fcol = 8 * nn + 2 * t;  // this thread's two output features of the n-tile
      float dh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const float* w = wa_hi + (8 * s2 + 2 * t) * kHmStride + 8 * nn + g;
        mma_tf32_1688(dh, fdm[s2], __float_as_uint(w[0]), __float_as_uint(w[kHmStride]));
        const float* h = sHa + (8 * s2 + 2 * t) * kHmStride + 8 * nn + g;
        mma_tf32_1688(gwa[nn], fdt[s2], __float_as_uint(h[0]), __float_as_uint(h[kHmStride]));
      }
      const float2 haA = *reinterpret_cast<const float2*>(sHa + g * kHmStride + fcol);
      const float2 haB = *reinterpret_cast<const float2*>(sHa + (g + 8) * kHmStride + fcol);
      const float2 hcA = *reinterpret_cast<const float2*>(sHc + g * kHmStride + fcol);
      const float2 hcB = *reinterpret_cast<const float2*>(sHc + (g + 8) * kHmStride + fcol);
      const float2 wc = *reinterpret_cast<const float2*>(wc_full + fcol);
      if (okA) {
        store_pair<PREC>(a.dZ3[1], (size_t)rowA, 128, fcol, dh[0] * elu_grad_from_output(haA.x), dh[1] * elu_grad_from_output(haA.y));
        store_pair<PREC>(a.dZ3[0], (size_t)rowA, 128, fcol, dvv[0] * wc.x * elu_grad_from_output(hcA.x), dvv[0] * wc.y * elu_grad_from_output(hcA.y));
      }
      if (okB) {
        store_pair<PREC>(a.dZ3[1], (size_t)rowB, 128, fcol, dh[2] * elu_grad_from_output(haB.x), dh[3] * elu_grad_from_output(haB.y));
        store_pair<PREC>(a.dZ3[0], (size_t)rowB, 128, fcol, dvv[1] * wc.x * elu_grad_from_output(hcB.x), dvv[1] * wc.y * elu_grad_from_output(hcB.y));
      }
      gwc[nn][0] = fmaf(dvv[0], hcA.x, fmaf(dvv[1], hcB.x, gwc[nn][0]));
      gwc[nn][1] = fmaf(dvv[0], hcA.y, fmaf(dvv[1], hcB.y, gwc[nn][1]));
    }
  }

  // ---- CTA-level reduction in head_kernel's format: every warp parks its values in a [kHeadValues][32] row set (aliasing
  //      the tiles), the CTA sums over its warps in a fixed order and writes ONE partial row per CTA
  // sums over the 8 row groups (lanes with the same t): gW4c, bias / log-std gradients, scalars
#pragma unroll
  for (int o = 4; o <= 16; o <<= 1) {
#pragma unroll
    for (int nn = 0; nn < 16; ++nn) {
      gwc[nn][0] += __shfl_xor_sync(0xffffffffu, gwc[nn][0], o);
      gwc[nn][1] += __shfl_xor_sync(0xffffffffu, gwc[nn][1], o);
    }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        g_b4a[n][e] += __shfl_xor_sync(0xffffffffu, g_b4a[n][e], o);
        g_logstd[n][e] += __shfl_xor_sync(0xffffffffu, g_logstd[n][e], o);
      }
    g_b4c += __shfl_xor_sync(0xffffffffu, g_b4c, o);
    l_pg += __shfl_xor_sync(0xffffffffu, l_pg, o);
    l_v += __shfl_xor_sync(0xffffffffu, l_v, o);
    l_kl += __shfl_xor_sync(0xffffffffu, l_kl, o);
    l_clip += __shfl_xor_sync(0xffffffffu, l_clip, o);
    l_oldkl += __shfl_xor_sync(0xffffffffu, l_oldkl, o);
  }
  __syncthreads();  // every warp is done with its tiles
  float* rows_all = hm + kHmWeights;
  float* mine = rows_all + (size_t)warp * kHeadValues * kHmPark;  // [value][lane'] rows of 33 floats (scatter below: 32 banks)
  for (int o = lane; o < (kHeadValues - kHvB4a) * kHmPark; o += 32) mine[kHvB4a * kHmPark + o] = 0.0f;  // sparse rows: zero first
  __syncwarp();
#pragma unroll
  for (int nn = 0; nn < 16; ++nn)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = g + 8 * (c >> 1), f = 8 * nn + 2 * t + (c & 1);
      mine[(j * 4 + (f & 3)) * kHmPark + (f >> 2)] = gwa[nn][c];  // value j*4 + f%4, lane f/4 (head_kernel's lane owns 4 features)
    }
  if (g == 0) {
#pragma unroll
    for (int nn = 0; nn < 16; ++nn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int f = 8 * nn + 2 * t + e;
        mine[(kHvW4c + (f & 3)) * kHmPark + (f >> 2)] = gwc[nn][e];
      }
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 8 * n + 2 * t + e;
        if (j < A) {  // per-action scalars live in the even lane of pair (2j, 2j+1)
          mine[kHvB4a * kHmPark + 2 * j] = g_b4a[n][e];
          mine[kHvLogstd * kHmPark + 2 * j] = g_logstd[n][e];
        }
      }
    if (t == 0) {
      const float scal[6] = {g_b4c, l_pg, l_v, l_kl, l_clip, l_oldkl};
#pragma unroll
      for (int k = 0; k < 6; ++k) mine[(kHvScalars + k) * kHmPark] = scal[k];
    }
  }
  __syncthreads();
  float* __restrict__ row = a.head_part + (size_t)blockIdx.x * kHeadValues * 32;
  for (int o = threadIdx.x; o < kHeadValues * 32; o += blockDim.x) {
    const int po = (o >> 5) * kHmPark + (o & 31);
    float tsum = 0.0f;
#pragma unroll
    for (int w = 0; w < kHmWarps; ++w) tsum += rows_all[(size_t)w * kHeadValues * kHmPark + po];
    row[o] = tsum;
  }
}

// ---- small utility kernels ------------------------------------------------------------------------------
template <int PREC>
__global__ void obs_to_operand_kernel(const float* __restrict__ obs, long long rows, int dim, int pad,
                                      typename PrecT<PREC>::T* __restrict__ out) {
  const long long total = rows * pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / pad;
    const int c = (int)(e - r * pad);
    const float v = c < dim ? obs[r * dim + c] : 0.0f;
    if (PREC == kPrecBf16) reinterpret_cast<bf16*>(out)[e] = __float2bfloat16(v);
    else reinterpret_cast<float*>(out)[e] = round_tf32(v);
  }
}

struct CastSeg {
  const float* src; void* dst; void* dst_t;
  int rows, cols, cols_pad;  // src [rows, cols] -> dst [rows, cols_pad] and dst_t [cols, rows]
};
struct CastArgs { CastSeg seg[6]; };

// 32x32 tiles: coalesced fp32 reads, coalesced writes of W, and a shared-memory transpose for W^T.
// grid = (max tiles over segments, 6 segments), block = (32, 8).
template <int PREC>
__global__ void __launch_bounds__(256) cast_weights_kernel(const __grid_constant__ CastArgs c) {
  using T = typename PrecT<PREC>::T;
  pdl_launch_dependents();
  pdl_wait();
  const CastSeg& s = c.seg[blockIdx.y];
  const int tiles_c = (s.cols_pad + 31) / 32, tiles_r = s.rows / 32;
  if ((int)blockIdx.x >= tiles_c * tiles_r) return;
  const int tr = blockIdx.x / tiles_c, tc = blockIdx.x % tiles_c;
  __shared__ float tile[32][33];
  T* dst = static_cast<T*>(s.dst);
  T* dst_t = static_cast<T*>(s.dst_t);
  auto conv = [](float v) -> T {
    if (PREC == kPrecBf16) return T(__float2bfloat16(v));
    return T(round_tf32(v));
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 32 + threadIdx.y + i * 8, k = tc * 32 + threadIdx.x;
    const float v = k < s.cols ? s.src[(size_t)r * s.cols + k] : 0.0f;
    tile[threadIdx.y + i * 8][threadIdx.x] = v;
    if (k < s.cols_pad) dst[(size_t)r * s.cols_pad + k] = conv(v);
  }
  if (dst_t == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = tc * 32 + threadIdx.y + i * 8, r = tr * 32 + threadIdx.x;
    if (k < s.cols) dst_t[(size_t)k * s.rows + r] = conv(tile[threadIdx.x][threadIdx.y + i * 8]);
  }
}

// Adam fused with the operand-copy refresh.  blockIdx.y < 6: one hidden weight matrix, 32 x 32 tiles as in
// cast_weights_kernel (update every element, write it back, emit W [rows, cols_pad] and W^T [cols, rows] in operand
// precision); blockIdx.y == 6: everything else of the flat vector (biases, heads, log-std: `n_rest` elements in at most
// 8 contiguous ranges), one thread per element.
struct AdamCastArgs {
  CastSeg seg[6];
  long long seg_off[6];      // flat offset of each hidden weight matrix
  long long rest_begin[8];   // flat ranges [begin, begin + len) outside the hidden matrices
  int rest_len[8];
  int n_rest_ranges;
  AdamState a;
};

// one 32 x 32 tile of hidden matrix `seg`: Adam on every element, W [rows, cols_pad] and W^T [cols, rows] in operand
// precision on the way (tx = lane, ty = warp of a 256-thread CTA; `tile` is the CTA's transpose buffer)
template <int PREC>
__device__ __forceinline__ void adam_cast_tile(const AdamCastArgs& c, int seg, int t, int tx, int ty, float clip, float step_size,
                                               float bc2_sqrt, float (*tile)[33]) {
  using T = typename PrecT<PREC>::T;
  const AdamState& a = c.a;
  const CastSeg& s = c.seg[seg];
  const long long base = c.seg_off[seg];
  const int tiles_c = (s.cols_pad + 31) / 32;
  const int tr = t / tiles_c, tc = t % tiles_c;
  T* dst = static_cast<T*>(s.dst);
  T* dst_t = static_cast<T*>(s.dst_t);
  auto conv = [](float v) -> T {
    if (PREC == kPrecBf16) return T(__float2bfloat16(v));
    return T(round_tf32(v));
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 32 + ty + i * 8, k = tc * 32 + tx;
    float v = 0.0f;
    if (k < s.cols) {
      const long long e = base + (long long)r * s.cols + k;
      v = a.params[e];
      float g = a.grads[e], m = a.m[e], vv = a.v[e];
      adam_one(v, g, m, vv, clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
      a.params[e] = v; a.grads[e] = g; a.m[e] = m; a.v[e] = vv;
    }
    tile[ty + i * 8][tx] = v;
    if (k < s.cols_pad) dst[(size_t)r * s.cols_pad + k] = conv(v);
  }
  if (dst_t == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = tc * 32 + ty + i * 8, r = tr * 32 + tx;
    if (k < s.cols) dst_t[(size_t)k * s.rows + r] = conv(tile[tx][ty + i * 8]);
  }
}

// element e of "everything else" (biases, heads, log-std)
__device__ __forceinline__ void adam_rest(const AdamCastArgs& c, int e, float clip, float step_size, float bc2_sqrt) {
  const AdamState& a = c.a;
  for (int r = 0; r < c.n_rest_ranges; ++r) {
    if (e < c.rest_len[r]) {
      const long long i = c.rest_begin[r] + e;
      adam_one(a.params[i], a.grads[i], a.m[i], a.v[i], clip, step_size, bc2_sqrt, a.beta1, a.beta2, a.eps);
      return;
    }
    e -= c.rest_len[r];
  }
}

template <int PREC>
__global__ void __launch_bounds__(256) adam_cast_kernel(const __grid_constant__ AdamCastArgs c) {
  pdl_launch_dependents();
  pdl_wait();
  const AdamState& a = c.a;
  const float clip = a.sc->clip_coef * a.grad_scale;
  const float step_size = __ldg(a.lr) * a.sc->step_size_scale;
  const float bc2_sqrt = a.sc->bc2_sqrt;
  if (blockIdx.y == 6) {
    adam_rest(c, blockIdx.x * 256 + threadIdx.y * 32 + threadIdx.x, clip, step_size, bc2_sqrt);
    return;
  }
  const CastSeg& s = c.seg[blockIdx.y];
  if ((int)blockIdx.x >= ((s.cols_pad + 31) / 32) * (s.rows / 32)) return;
  __shared__ float tile[32][33];
  adam_cast_tile<PREC>(c, blockIdx.y, blockIdx.x, threadIdx.x, threadIdx.y, clip, step_size, bc2_sqrt, tile);
}

// ---- the whole optimizer step of one minibatch as ONE launch (single GPU) --------------------------------------------
// fold (padded accumulators + head rows -> flat gradient) | squared norm of the flat gradient  -- grid barrier --
// clip coefficient, bias corrections | Adam on every element + operand copies.  Replaces fold_grads_kernel +
// grad_norm_kernel + adam_cast_kernel (11 + 6 + 6 us of three latency-bound launches over 1.5 MB).  Up to 4 CTAs per SM, all
// co-resident (256 threads, 40 registers, < 6 KiB of shared memory; the launch asks the occupancy calculator): the barrier is an arrival counter in the
// optimizer scratch that only ever grows (arrival ticket / grid = generation), the sum of squares is reset by the last CTA
// that has read it (ticket), so the scratch is clean for the next launch and for catb200_adam_step.
struct OptStepArgs {
  FoldArgs f;
  AdamCastArgs c;
  int quad_prefix[7];          // fold segments: prefix sums of their 16-byte quads
  int tile_prefix[7];          // hidden matrices: prefix sums of their 32 x 32 tiles
  long long bias_begin[6];     // hidden-layer bias gradients (written by mlp_wgrad_kernel directly): flat offsets
  int bias_len[6];
  int n_rest;
  float max_norm;
  int* step;
  float* grad_norm_out;
  OptScratch* sc;
};

template <int PREC>
__global__ void __launch_bounds__(256) opt_step_kernel(const __grid_constant__ OptStepArgs r) {
  __shared__ float part_s[8][32];
  __shared__ float tile[32][33];
  __shared__ double red_s[8];
  __shared__ float coef_s[3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gtid = blockIdx.x * 256 + tid, gthreads = gridDim.x * 256;
  const float gscale = r.c.a.grad_scale;
  OptScratch* sc = r.sc;
  const int step0 = *r.step;  // CTA 0 advances it only after the barrier below

  // ---- phase A: fold + sum of squares of every element of the flat gradient
  double ss = 0.0;
  const int total_quads = r.quad_prefix[r.f.n_segments];
  for (int q = gtid; q < total_quads; q += gthreads) {
    int seg = 0;
    while (q >= r.quad_prefix[seg + 1]) ++seg;
    ss += (double)fold_weight_quad(r.f, seg, q - r.quad_prefix[seg]);
  }
  if ((int)blockIdx.x < kHeadValues) ss += (double)fold_head_group(r.f, blockIdx.x, part_s);
  {  // hidden-layer biases, dealt from the far end of the grid (the first 76 CTAs carry the head rows)
    int e = (gridDim.x - 1 - blockIdx.x) * 256 + tid;
    for (int b = 0; b < 6; ++b) {
      if (e < r.bias_len[b]) {
        const float g = r.c.a.grads[r.bias_begin[b] + e];
        ss += (double)g * (double)g;
        break;
      }
      e -= r.bias_len[b];
    }
  }
  ss *= (double)gscale * (double)gscale;
  ss = warp_sum(ss);
  if (lane == 0) red_s[warp] = ss;
  __syncthreads();
  // ---- grid barrier (release: every thread's writes of this CTA are ordered before thread 0's fence + arrival)
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red_s[w];
    atomicAdd(&sc->sumsq, t);
    __threadfence();
    unsigned int* bar = reinterpret_cast<unsigned int*>(&sc->pad[0]);
    const unsigned int ticket = atomicAdd(bar, 1u);
    const unsigned int target = (ticket / gridDim.x + 1u) * gridDim.x;
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<volatile unsigned int*>(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000ll) __trap();  // a CTA that never arrived (not co-resident): fail, do not hang
    }
    __threadfence();
    const double tot = *reinterpret_cast<volatile double*>(&sc->sumsq);
    const float norm = (float)sqrt(tot);
    // torch.nn.utils.clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float clip_coef = fminf(r.max_norm / (norm + 1e-6f), 1.0f);
    const int t1 = step0 + 1;
    const double bc1 = 1.0 - pow((double)r.c.a.beta1, (double)t1), bc2 = 1.0 - pow((double)r.c.a.beta2, (double)t1);
    coef_s[0] = clip_coef * gscale;
    coef_s[1] = __ldg(r.c.a.lr) * (float)(1.0 / bc1);
    coef_s[2] = (float)sqrt(bc2);
    if (blockIdx.x == 0) {
      sc->clip_coef = clip_coef;
      sc->total_norm = norm;
      sc->step_size_scale = (float)(1.0 / bc1);
      sc->bc2_sqrt = (float)sqrt(bc2);
      *r.step = t1;
      if (r.grad_norm_out) *r.grad_norm_out = norm;
    }
    // the last CTA to have read the sum resets it (measured: behind phase C instead, the launch takes 19.0 instead of 17.8 us)
    __threadfence();
    if (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1) {
      sc->ticket = 0u;
      sc->sumsq = 0.0;
    }
  }
  __syncthreads();
  const float clip = coef_s[0], step_size = coef_s[1], bc2_sqrt = coef_s[2];

  // ---- phase C: Adam + operand copies
  const int n_tiles = r.tile_prefix[6], n_items = n_tiles + (r.n_rest + 255) / 256;
  for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    if (w < n_tiles) {
      int seg = 0;
      while (w >= r.tile_prefix[seg + 1]) ++seg;
      __syncthreads();  // the previous tile's transpose reads are done
      adam_cast_tile<PREC>(r.c, seg, w - r.tile_prefix[seg], lane, warp, clip, step_size, bc2_sqrt, tile);
    } else {
      adam_rest(r.c, (w - n_tiles) * 256 + tid, clip, step_size, bc2_sqrt);
    }
  }
}

// ---- the optimizer step of one minibatch on several GPUs as ONE launch --------------------------------------------------
// opt_step_kernel with the gradient exchange of peer.cu in the middle (catb200_ppo_minibatch_update_peer):
//   A  fold the padded accumulators + head rows into this rank's peer-visible arena of the minibatch's parity
//      -- local grid barrier; its last arriver announces the arena to every peer (st.release.sys into their flag rows) --
//      every CTA waits until all peers have announced theirs (bounded)
//   B  rank-ordered sum of the `world` arenas over NVLink -> private gradient + squared norm; the OTHER arena is zeroed
//      -- local grid barrier: clip coefficient, bias corrections --
//   C  Adam + operand copies from the private sum.
// Replaces fold_grads_kernel + grad_allreduce_norm_kernel + adam_cast_kernel (10 + 17 + 7 us per optimizer step at 2 GPUs).
struct OptStepPeerArgs {
  OptStepArgs o;                  // o.f.grad[] point into this rank's arena, o.c.a.grads is the private sum
  const float* arena[kPeerMax];   // arena of this minibatch's parity on every rank (index = rank)
  uint32_t* flags[kPeerMax];      // flag row of every rank
  float* zero_arena;              // this rank's other arena
  float* out;                     // private summed gradient [n]
  long long n;
  int rank, world;
  unsigned int* epoch;            // device-local count of completed exchanges
  int* err;                       // device-local error flag (1: a peer did not arrive, 2: parity out of step)
  unsigned int parity;
  // reduce-scatter / all-gather variant (world > 2): every rank sums ONE slice of the vector over all arenas and writes it
  // into every rank's summed-gradient region, so a rank moves 2 (W-1)/W of the vector over NVLink instead of pulling W-1
  // whole arenas (10.5 MB per step at 8 ranks: 48 us); costs a second handshake, which also carries the partial norms
  int rs;
  float* gsum[kPeerMax];          // summed-gradient region of every rank's peer block (index = rank)
};

template <int PREC>
__global__ void __launch_bounds__(256) opt_step_peer_kernel(const __grid_constant__ OptStepPeerArgs p) {
  const OptStepArgs& r = p.o;
  __shared__ float part_s[8][32];
  __shared__ float tile[32][33];
  __shared__ double red_s[8];
  __shared__ float coef_s[3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gtid = blockIdx.x * 256 + tid, gthreads = gridDim.x * 256;
  const float gscale = r.c.a.grad_scale;
  OptScratch* sc = r.sc;
  const int step0 = *r.step;               // CTA 0 advances it only after the second barrier
  const unsigned int e = *p.epoch + 1u;    // likewise
  unsigned int* bar = reinterpret_cast<unsigned int*>(&sc->pad[0]);
  auto grid_barrier = [&](bool announce) {  // thread 0 only; arrival counter that only ever grows (ticket / grid = generation)
    const unsigned int ticket = atomicAdd(bar, 1u);
    const unsigned int target = (ticket / gridDim.x + 1u) * gridDim.x;
    if (announce && ticket == target - 1u) {
      // last arriver: every CTA fenced its arena writes (system scope) before its arrival, which this thread has observed
      __threadfence_system();  // one release fence for all the flag stores (a st.release per peer would serialise W - 1 fences)
      for (int q = 0; q < p.world; ++q)
        if (q != p.rank) st_relaxed_sys(p.flags[q] + p.rank, e);
    }
    const long long t0 = clock64();
    while ((int)(*reinterpret_cast<volatile unsigned int*>(bar) - target) < 0) {
      if (clock64() - t0 > 4000000000ll) __trap();  // a CTA that never arrived (not co-resident): fail, do not hang
    }
    __threadfence();
  };

  // ---- phase A: fold into the arena
  const int total_quads = r.quad_prefix[r.f.n_segments];
  for (int q = gtid; q < total_quads; q += gthreads) {
    int seg = 0;
    while (q >= r.quad_prefix[seg + 1]) ++seg;
    fold_weight_quad(r.f, seg, q - r.quad_prefix[seg]);
  }
  if ((int)blockIdx.x < kHeadValues) fold_head_group(r.f, blockIdx.x, part_s);
  __syncthreads();
  if (tid == 0) {
    if (((*p.epoch) & 1u) != p.parity) atomicExch(p.err, 2);
    __threadfence_system();
    grid_barrier(true);
  }
  __syncthreads();
  auto wait_flags = [&](int word0, bool with_self) {  // threads < world: every rank's epoch-e flag has arrived in this rank's row
    if (tid < p.world && (with_self || tid != p.rank)) {
      const uint32_t* mine = p.flags[p.rank] + word0 + tid;
      const long long t0 = clock64();
      while (ld_acquire_sys(mine) < e) {  // (relaxed polls + one fence.sys behind the loop measured slower: 38.6 vs 33.9 us at 2 ranks)
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: give up loudly, do not hang the GPU
          atomicExch(p.err, 1);
          break;
        }
        __nanosleep(200);
      }
    }
    __syncthreads();
  };
  auto publish_coefficients = [&](double tot) {  // thread 0: clip coefficient + Adam bias corrections from the squared norm
    const float norm = (float)sqrt(tot);
    const float clip_coef = fminf(r.max_norm / (norm + 1e-6f), 1.0f);  // torch.nn.utils.clip_grad_norm_
    const int t1 = step0 + 1;
    const double bc1 = 1.0 - pow((double)r.c.a.beta1, (double)t1), bc2 = 1.0 - pow((double)r.c.a.beta2, (double)t1);
    coef_s[0] = clip_coef * gscale;
    coef_s[1] = __ldg(r.c.a.lr) * (float)(1.0 / bc1);
    coef_s[2] = (float)sqrt(bc2);
    if (blockIdx.x == 0) {
      sc->clip_coef = clip_coef;
      sc->total_norm = norm;
      sc->step_size_scale = (float)(1.0 / bc1);
      sc->bc2_sqrt = (float)sqrt(bc2);
      *r.step = t1;
      *p.epoch = e;
      if (r.grad_norm_out) *r.grad_norm_out = norm;
    }
  };

  wait_flags(0, false);  // every peer's arena of this parity is complete

  double ss = 0.0;
  if (p.rs) {
    // ---- phase B (reduce-scatter / all-gather): this rank's slice of the vector, summed over the arenas in rank order
    //      (so the slice is bit-identical wherever it lands), goes to every rank's summed-gradient region
    const long long n4 = (long long)(peer_n_pad(p.n) / 4);  // whole 16-byte quads; the padding stays zero
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo = per * p.rank, hi = min(n4, lo + per);
    for (long long i = lo + gtid; i < hi; i += gthreads) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int q = 0; q < p.world; ++q) {
        const float4 v = __ldcv(reinterpret_cast<const float4*>(p.arena[q]) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
#pragma unroll 8
      for (int q = 0; q < p.world; ++q) reinterpret_cast<float4*>(p.gsum[q])[i] = acc;
      const double x = (double)(acc.x * gscale), y = (double)(acc.y * gscale), z = (double)(acc.z * gscale), w = (double)(acc.w * gscale);
      ss += x * x + y * y + z * z + w * w;
    }
    for (long long i = gtid; i < n4; i += gthreads) reinterpret_cast<float4*>(p.zero_arena)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ss = warp_sum(ss);
    if (lane == 0) red_s[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red_s[w];
      atomicAdd(&sc->sumsq, t);
      __threadfence_system();  // this CTA's slice stores (remote ones included) are ordered before its arrival
      const unsigned int ticket = atomicAdd(bar, 1u);
      const unsigned int target = (ticket / gridDim.x + 1u) * gridDim.x;
      if (ticket == target - 1u) {
        // last arriver: the slice is complete everywhere.  Its squared norm goes into slot `rank` of every rank's row (own
        // one included), then the second flag.  Nobody waits on the arrival counter itself: every CTA waits for the flags.
        __threadfence();
        const double part = *reinterpret_cast<volatile double*>(&sc->sumsq);
        sc->sumsq = 0.0;
        for (int q = 0; q < p.world; ++q)
          reinterpret_cast<double*>(reinterpret_cast<char*>(p.flags[q]) + kNormByte)[p.rank] = part;
        __threadfence_system();
        for (int q = 0; q < p.world; ++q) st_relaxed_sys(p.flags[q] + kFlag2Word + p.rank, e);
      }
    }
    wait_flags(kFlag2Word, true);  // every rank's slice and partial norm have landed here
    if (tid == 0) {
      double tot = 0.0;
      const volatile double* parts = reinterpret_cast<const volatile double*>(reinterpret_cast<char*>(p.flags[p.rank]) + kNormByte);
      for (int q = 0; q < p.world; ++q) tot += parts[q];  // rank order: the same total on every rank
      publish_coefficients(tot);
    }
  } else {
    // ---- phase B (pull): rank-ordered sum of the whole arenas (identical on every rank), squared norm, zero the other arena
    const long long n4 = p.n / 4;
    for (long long i = gtid; i < n4; i += gthreads) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int q = 0; q < p.world; ++q) {
        const float4 v = __ldcv(reinterpret_cast<const float4*>(p.arena[q]) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(p.out)[i] = acc;
      reinterpret_cast<float4*>(p.zero_arena)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const double x = (double)(acc.x * gscale), y = (double)(acc.y * gscale), z = (double)(acc.z * gscale), w = (double)(acc.w * gscale);
      ss += x * x + y * y + z * z + w * w;
    }
    for (long long i = n4 * 4 + gtid; i < p.n; i += gthreads) {  // scalar tail
      float acc = 0.f;
      for (int q = 0; q < p.world; ++q) acc += __ldcv(p.arena[q] + i);
      p.out[i] = acc;
      p.zero_arena[i] = 0.f;
      const double x = (double)(acc * gscale);
      ss += x * x;
    }
    ss = warp_sum(ss);
    if (lane == 0) red_s[warp] = ss;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red_s[w];
      atomicAdd(&sc->sumsq, t);
      __threadfence();
      grid_barrier(false);
      publish_coefficients(*reinterpret_cast<volatile double*>(&sc->sumsq));
      __threadfence();
      if (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1) {  // the last CTA to have read the sum resets it
        sc->ticket = 0u;
        sc->sumsq = 0.0;
      }
    }
  }
  __syncthreads();
  const float clip = coef_s[0], step_size = coef_s[1], bc2_sqrt = coef_s[2];

  // ---- phase C: Adam + operand copies from the private sum
  const int n_tiles = r.tile_prefix[6], n_items = n_tiles + (r.n_rest + 255) / 256;
  for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
    if (w < n_tiles) {
      int seg = 0;
      while (w >= r.tile_prefix[seg + 1]) ++seg;
      __syncthreads();  // the previous tile's transpose reads are done
      adam_cast_tile<PREC>(r.c, seg, w - r.tile_prefix[seg], lane, warp, clip, step_size, bc2_sqrt, tile);
    } else {
      adam_rest(r.c, (w - n_tiles) * 256 + tid, clip, step_size, bc2_sqrt);
    }
  }
}

// advances the device-side Philox offset after a sampling launch (keeps the launch graph-capturable)
__global__ void rng_bump_kernel(unsigned long long* rng_state, unsigned long long n) { rng_state[1] += n; }

// ---- host-side layout helpers ---------------------------------------------------------------------------
struct Dims {
  int in[3], out[3], in_pad[3];
};

static bool dims_ok(const catb200_mlp_dims_t* d) {
  if (!d) return false;
  if (d->obs_dim <= 0 || d->obs_pad < d->obs_dim || d->obs_pad % 64 != 0 || d->obs_pad > 256) return false;
  if (d->act_dim <= 0 || d->act_dim > kMaxAct) return false;
  if (d->h1 % 128 || d->h2 % 128 || d->h1 <= 0 || d->h2 <= 0) return false;
  if (d->h3 != 128) return false;  // the head kernel maps one lane to 4 of 128 features
  if (d->prec != kPrecBf16 && d->prec != kPrecTf32) return false;
  return true;
}

static Dims make_dims(const catb200_mlp_dims_t* d) {
  Dims x;
  x.in[0] = d->obs_dim; x.in_pad[0] = d->obs_pad; x.out[0] = d->h1;
  x.in[1] = d->h1; x.in_pad[1] = d->h1; x.out[1] = d->h2;
  x.in[2] = d->h2; x.in_pad[2] = d->h2; x.out[2] = d->h3;
  return x;
}

static size_t esize(const catb200_mlp_dims_t* d) { return d->prec == kPrecTf32 ? 4 : 2; }
static int ch_elems(const catb200_mlp_dims_t* d) { return d->prec == kPrecTf32 ? 32 : 64; }  // elements per 128-byte row

struct ActLayout {  // byte offsets into the activation workspace
  size_t X, H[2][3], dZ[2][3], mb, gacc[2][3], head_part, scal_mb, act_mb, total;
  int splits[3], m_range[3], head_rows;
};

static ActLayout act_layout(const catb200_mlp_dims_t* d, int rows, bool training) {
  ActLayout L = {};
  Dims x = make_dims(d);
  const size_t es = esize(d);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  L.mb = take(sizeof(MbStats));
  if (training) {
    // zero-initialised by the caller, left zero by fold_grads_kernel: at fixed offsets so that workspaces sized for
    // different row counts share them
    for (int l = 0; l < 3; ++l)
      for (int z = 0; z < 2; ++z) L.gacc[z][l] = take((size_t)x.out[l] * x.in_pad[l] * 4);
  }
  L.X = take((size_t)rows * d->obs_pad * es);
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) L.H[z][l] = take((size_t)rows * x.out[l] * es);
  if (training) {
    L.head_rows = min((rows + 7) / 8, kNumSMs);
    L.head_part = take((size_t)L.head_rows * kHeadValues * 32 * 4);
    L.scal_mb = take((size_t)rows * 16);
    L.act_mb = take((size_t)rows * d->act_dim * 4);
    for (int z = 0; z < 2; ++z)
      for (int l = 0; l < 3; ++l) L.dZ[z][l] = take((size_t)rows * x.out[l] * es);
    for (int l = 0; l < 3; ++l) {
      const int bn = tc_wgrad_bn(d->prec, x.in_pad[l]);
      const int tiles = (x.out[l] / 128) * (x.in_pad[l] / bn) * 2;
      const int want = max(1, kNumSMs / tiles);  // ~one CTA per SM per layer: fewer, fatter splits = fewer reductions
      int m_range = ((rows + want - 1) / want + 63) / 64 * 64;
      m_range = max(m_range, 64);
      L.m_range[l] = m_range;
      L.splits[l] = (rows + m_range - 1) / m_range;
    }
  }
  L.total = off;
  return L;
}

static int fill_layout(const catb200_mlp_dims_t* d, catb200_mlp_layout_t* L) {
  Dims x = make_dims(d);
  int64_t off = 0;
  for (int z = 0; z < 2; ++z) {
    for (int l = 0; l < 3; ++l) {
      L->w[z][l] = off; off += (int64_t)x.out[l] * x.in[l];
      L->b[z][l] = off; off += x.out[l];
    }
    const int head = z == 0 ? 1 : d->act_dim;
    L->w[z][3] = off; off += (int64_t)head * d->h3;
    L->b[z][3] = off; off += head;
  }
  L->logstd = off; off += d->act_dim;
  L->n_params = off;
  int64_t oc = 0;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      L->wc[z][l] = oc; oc += (int64_t)x.out[l] * x.in_pad[l];
      if (l > 0) { L->wtc[z][l] = oc; oc += (int64_t)x.in[l] * x.out[l]; }
      else L->wtc[z][l] = -1;
    }
  L->n_wc = oc;
  return CATB200_OK;
}

// Launches whose activations outgrow the L2 alternate the direction in which they walk the rows (tc_gemm.cu, "row
// order"): layer 0 upwards, layer 1 downwards, ... so that every launch starts on the rows the previous one finished with.
// Opt-in (CATB200_ZIGZAG=1): the L2 did not reward it.
static bool zigzag_rows(const catb200_mlp_dims_t* d, int rows) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("CATB200_ZIGZAG");
    enabled = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured neutral on the B200 (210.5 vs 211.4 us per minibatch)
  }
  // both nets' activations and their gradients: 2 * 2 * (h1 + h2 + h3) elements per row
  const size_t bytes = (size_t)rows * 4 * (d->h1 + d->h2 + d->h3) * (d->prec == kPrecTf32 ? 4 : 2);
  return enabled == 1 && bytes > (size_t)96 << 20;
}

// L2 eviction-priority hints on the TMA traffic of the minibatch-sized launches (tc_ptx.cuh): what the next launch re-reads
// is kept (evict_last), what is dead after this read goes first.  CATB200_L2_HINTS=0 disables them.
static bool l2_hints(int rows) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_L2_HINTS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && rows >= 8192;  // rollout-sized launches fit the L2 anyway
}

static int launch_forward(const catb200_mlp_dims_t* d, const catb200_mlp_layout_t& P, const ActLayout& L, const void* X,
                          int rows, const float* params, const void* wc, char* ws, cudaStream_t st, bool zigzag = false) {
  Dims x = make_dims(d);
  const size_t es = esize(d);
  const int ch = ch_elems(d), prec = d->prec;
  const char* wcb = static_cast<const char*>(wc);
  for (int l = 0; l < 3; ++l) {
    TcGemmArgs t = {};
    for (int z = 0; z < 2; ++z) {
      const void* A = l == 0 ? X : static_cast<const void*>(ws + L.H[z][l - 1]);
      int rc = make_tmap(&t.mapA[z], prec, A, x.in_pad[l], rows, x.in_pad[l], ch, 128);
      if (rc == CATB200_OK) rc = make_tmap(&t.mapB[z], prec, wcb + P.wc[z][l] * es, x.in_pad[l], x.out[l], x.in_pad[l], ch, 128);
      if (rc == CATB200_OK) rc = make_tmap(&t.mapC[z], prec, ws + L.H[z][l], x.out[l], rows, x.out[l], ch, 32);
      if (rc != CATB200_OK) return rc;
      t.bias[z] = params + P.b[z][l];
    }
    t.M = rows; t.N = x.out[l]; t.K = x.in_pad[l];
    t.reverse = zigzag && (l & 1);
    if (l2_hints(rows)) {  // H_{l-1} is not needed again before the backward pass; H_l is the next launch's operand
      t.hintA = kL2EvictFirst; t.hintB = kL2EvictLast; t.hintC = kL2EvictLast;
    }
    int rc = tc_gemm_launch(kTcFwd, prec, t, st);
    if (rc != CATB200_OK) return rc;
  }
  return CATB200_OK;
}

int launch_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* wcv, cudaStream_t st) {
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  const size_t es = esize(dims);
  char* wc = static_cast<char*>(wcv);
  CastArgs c;
  int max_tiles = 1;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      CastSeg& s = c.seg[z * 3 + l];
      s.src = params + P.w[z][l];
      s.dst = wc + P.wc[z][l] * es;
      s.dst_t = l > 0 ? wc + P.wtc[z][l] * es : nullptr;
      s.rows = x.out[l]; s.cols = x.in[l]; s.cols_pad = x.in_pad[l];
      max_tiles = max(max_tiles, (s.rows / 32) * ((s.cols_pad + 31) / 32));
    }
  if (dims->prec == kPrecTf32) CATB200_CUDA_TRY(launch_pdl(cast_weights_kernel<kPrecTf32>, dim3(max_tiles, 6), dim3(32, 8), 0, st, c));
  else CATB200_CUDA_TRY(launch_pdl(cast_weights_kernel<kPrecBf16>, dim3(max_tiles, 6), dim3(32, 8), 0, st, c));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// segments / ranges of the flat vector for the Adam kernels; returns the grid width adam_cast_kernel needs
static int fill_adam_cast_args(const catb200_mlp_dims_t* dims, const AdamState& a, void* wcv, AdamCastArgs* out) {
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  const size_t es = esize(dims);
  char* wc = static_cast<char*>(wcv);
  AdamCastArgs& c = *out;
  c = AdamCastArgs{};
  int max_tiles = 1;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      CastSeg& s = c.seg[z * 3 + l];
      s.src = a.params + P.w[z][l];
      s.dst = wc + P.wc[z][l] * es;
      s.dst_t = l > 0 ? wc + P.wtc[z][l] * es : nullptr;
      s.rows = x.out[l]; s.cols = x.in[l]; s.cols_pad = x.in_pad[l];
      c.seg_off[z * 3 + l] = P.w[z][l];
      max_tiles = max(max_tiles, (s.rows / 32) * ((s.cols_pad + 31) / 32));
    }
  // the rest of the flat vector: the gaps between consecutive hidden matrices (b0 | b1 | b2, head weight, head bias) and
  // what follows the last one (actor b2, head, log-std)
  int n_rest = 0;
  long long cursor = 0;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      if (P.w[z][l] > cursor) {
        c.rest_begin[c.n_rest_ranges] = cursor;
        c.rest_len[c.n_rest_ranges] = (int)(P.w[z][l] - cursor);
        n_rest += c.rest_len[c.n_rest_ranges++];
      }
      cursor = P.w[z][l] + (long long)x.out[l] * x.in[l];
    }
  if (P.n_params > cursor) {
    c.rest_begin[c.n_rest_ranges] = cursor;
    c.rest_len[c.n_rest_ranges] = (int)(P.n_params - cursor);
    n_rest += c.rest_len[c.n_rest_ranges++];
  }
  c.a = a;
  return max(max_tiles, (n_rest + 255) / 256);
}

int launch_adam_cast(const catb200_mlp_dims_t* dims, const AdamState& a, void* wcv, cudaStream_t st) {
  AdamCastArgs c;
  const int max_tiles = fill_adam_cast_args(dims, a, wcv, &c);
  if (dims->prec == kPrecTf32) CATB200_CUDA_TRY(launch_pdl(adam_cast_kernel<kPrecTf32>, dim3(max_tiles, 7), dim3(32, 8), 0, st, c));
  else CATB200_CUDA_TRY(launch_pdl(adam_cast_kernel<kPrecBf16>, dim3(max_tiles, 7), dim3(32, 8), 0, st, c));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// training heads: warp-level tensor-core kernel (default) or the warp-per-sample kernel (CATB200_HEAD=warp)
static bool head_use_mma(const catb200_mlp_dims_t* d) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_HEAD");
    v = (e && e[0] == 'w') ? 0 : 1;
  }
  return v == 1 && d->h3 == 128 && d->act_dim <= 16;
}

static int launch_head_mma(int prec, const HeadArgs& a, int grid, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(head_mma_kernel<kPrecTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHmSmem));
    CATB200_CUDA_TRY(cudaFuncSetAttribute(head_mma_kernel<kPrecBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHmSmem));
    attr = true;
  }
  if (prec == kPrecTf32) CATB200_CUDA_TRY(launch_pdl(head_mma_kernel<kPrecTf32>, dim3(grid), dim3(kHmWarps * 32), (size_t)kHmSmem, st, a));
  else CATB200_CUDA_TRY(launch_pdl(head_mma_kernel<kPrecBf16>, dim3(grid), dim3(kHmWarps * 32), (size_t)kHmSmem, st, a));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

template <bool TRAIN>
static int launch_head(int prec, const HeadArgs& a, int grid, size_t smem, cudaStream_t st) {
  static bool attr = false;
  if (TRAIN && !attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(head_kernel<true, kPrecTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
    CATB200_CUDA_TRY(cudaFuncSetAttribute(head_kernel<true, kPrecBf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
    attr = true;
  }
  if (prec == kPrecTf32) CATB200_CUDA_TRY(launch_pdl(head_kernel<TRAIN, kPrecTf32>, dim3(grid), dim3(kHeadThreads), smem, st, a));
  else CATB200_CUDA_TRY(launch_pdl(head_kernel<TRAIN, kPrecBf16>, dim3(grid), dim3(kHeadThreads), smem, st, a));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // namespace catb200

using namespace catb200;

extern "C" {

int catb200_mlp_layout(const catb200_mlp_dims_t* dims, catb200_mlp_layout_t* layout) {
  if (!dims || !layout) return CATB200_ERR_INVALID_ARGUMENT;
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  return fill_layout(dims, layout);
}

int catb200_mlp_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* wc, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!params || !wc) return CATB200_ERR_INVALID_ARGUMENT;
  return launch_cast_weights(dims, params, wc, as_stream(stream));
}

int catb200_obs_to_operand(const catb200_mlp_dims_t* dims, const float* obs, int64_t rows, void* obs_op, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!obs || !obs_op || rows <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  const long long total = rows * dims->obs_pad;
  const int grid = (int)min((total + 255) / 256, (long long)kNumSMs * 16);
  if (dims->prec == kPrecTf32)
    obs_to_operand_kernel<kPrecTf32><<<grid, 256, 0, as_stream(stream)>>>(obs, rows, dims->obs_dim, dims->obs_pad, static_cast<float*>(obs_op));
  else
    obs_to_operand_kernel<kPrecBf16><<<grid, 256, 0, as_stream(stream)>>>(obs, rows, dims->obs_dim, dims->obs_pad, static_cast<bf16*>(obs_op));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

size_t catb200_mlp_workspace_bytes(const catb200_mlp_dims_t* dims, int32_t rows, int32_t training) {
  if (!dims_ok(dims) || rows <= 0) return 0;
  return act_layout(dims, rows, training != 0).total;
}

int catb200_mlp_act(const catb200_mlp_dims_t* dims, const void* obs_op, int32_t rows, const float* params, const void* wc,
                    const float* noise, uint64_t* rng_state, const float* action_in, float* action, float* logprob,
                    float* value, float* mean_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!obs_op || rows <= 0 || !params || !wc || !workspace) return CATB200_ERR_INVALID_ARGUMENT;
  const ActLayout L = act_layout(dims, rows, false);
  if (workspace_bytes < L.total) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  int rc = launch_forward(dims, P, L, obs_op, rows, params, wc, ws, st);
  if (rc != CATB200_OK) return rc;
  HeadArgs a = {};
  for (int z = 0; z < 2; ++z) a.H3[z] = ws + L.H[z][2];
  a.W4c = params + P.w[0][3]; a.b4c = params + P.b[0][3];
  a.W4a = params + P.w[1][3]; a.b4a = params + P.b[1][3];
  a.logstd = params + P.logstd;
  a.M = rows; a.h3 = dims->h3; a.A = dims->act_dim;
  a.noise = noise; a.action_in = action_in; a.action = action; a.logprob = logprob; a.value = value; a.mean_out = mean_out;
  const bool sample = rng_state != nullptr && action_in == nullptr && noise == nullptr;
  a.rng_state = sample ? reinterpret_cast<const unsigned long long*>(rng_state) : nullptr;
  const int grid = min((rows + 7) / 8, kNumSMs * 4);
  rc = launch_head<false>(dims->prec, a, grid, 0, st);
  if (rc != CATB200_OK) return rc;
  if (sample) {  // one Philox counter block of 4 per element pair is ample: advance by the element count
    rng_bump_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(rng_state), (unsigned long long)rows * dims->act_dim);
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

}  // extern "C"

// forward + backward of one minibatch; the fold of the padded accumulators / head rows into `grads` is launched here, or
// -- defer_fold != NULL -- described there for the caller's fused optimizer-step kernel
static int minibatch_backward(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                              const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                              const float* logprobs_all, const float* advantages_all, const float* returns_all,
                              const float* values_all, const float* norm_stats, const float* params, const void* wcv,
                              float* grads, float* loss_acc, void* workspace, size_t workspace_bytes, void* stream,
                              FoldArgs* defer_fold) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!hp || mb_rows <= 0 || !mb_inds || !obs_op_all || !actions_all || !logprobs_all || !advantages_all || !returns_all ||
      !values_all || !norm_stats || !params || !wcv || !grads || !loss_acc || !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  const int M = mb_rows;
  const ActLayout L = act_layout(dims, M, true);
  if (workspace_bytes < L.total) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  const size_t es = esize(dims);
  const int ch = ch_elems(dims), prec = dims->prec;
  const int bk = 64;  // reduction rows per TMA box of the weight-gradient kernel (tc_gemm.cu: kWgRows)
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  const char* wc = static_cast<const char*>(wcv);
  void* X = ws + L.X;
  MbStats* mb = reinterpret_cast<MbStats*>(ws + L.mb);

  // 1. gather + advantage statistics
  const int row_bytes = (int)(dims->obs_pad * es);
  CATB200_CUDA_TRY(launch_pdl(gather_kernel, dim3(min((M * (row_bytes / 16) + 255) / 256, kNumSMs * 8)), dim3(256), 0, st,
                              mb_inds, M, static_cast<const uint4*>(obs_op_all), row_bytes, advantages_all, logprobs_all,
                              returns_all, values_all, actions_all, (int)dims->act_dim, static_cast<uint4*>(X),
                              reinterpret_cast<float4*>(ws + L.scal_mb), reinterpret_cast<float*>(ws + L.act_mb), mb));
  CATB200_LAUNCH_CHECK();
  // 2. forward through the three hidden layers of both nets
  const bool zigzag = zigzag_rows(dims, M);  // up: layers 0 and 2, the weight gradients; down: layer 1, heads, dgrads
  int rc = launch_forward(dims, P, L, X, M, params, wc, ws, st, zigzag);
  if (rc != CATB200_OK) return rc;
  // 3. heads + loss + dZ3
  {
    HeadArgs a = {};
    for (int z = 0; z < 2; ++z) {
      a.H3[z] = ws + L.H[z][2];
      a.dZ3[z] = ws + L.dZ[z][2];
    }
    a.W4c = params + P.w[0][3]; a.b4c = params + P.b[0][3];
    a.W4a = params + P.w[1][3]; a.b4a = params + P.b[1][3];
    a.logstd = params + P.logstd;
    a.M = M; a.h3 = dims->h3; a.A = dims->act_dim;
    a.scal_mb = reinterpret_cast<const float4*>(ws + L.scal_mb);
    a.act_mb = reinterpret_cast<const float*>(ws + L.act_mb);
    a.norm_stats = norm_stats; a.mb = mb; a.hp = *hp;
    a.head_part = reinterpret_cast<float*>(ws + L.head_part);
    a.reverse = zigzag;
    rc = head_use_mma(dims) ? launch_head_mma(prec, a, L.head_rows, st) : launch_head<true>(prec, a, L.head_rows, kHeadSmem, st);
    if (rc != CATB200_OK) return rc;
  }
  // 4. backward through the hidden layers
  FoldArgs fold = {};
  SideStream* side = side_stream();
  for (int l = 2; l >= 0; --l) {
    // dZ_l is complete on `st` here (head kernel or the previous dgrad): the weight gradient may start beside dgrad
    cudaStream_t wst = st;
    if (side && l > 0) {
      CATB200_CUDA_TRY(cudaEventRecord(side->dz_ready[l], st));
      CATB200_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->dz_ready[l], 0));
      wst = side->stream;
    }
    {  // dW_l += dZ_l^T H_{l-1}, db_l += dZ_l^T 1
      TcWgradArgs t = {};
      for (int z = 0; z < 2; ++z) {
        const void* Hin = l == 0 ? X : static_cast<const void*>(ws + L.H[z][l - 1]);
        const int swz32 = prec == kPrecTf32 ? 1 : 0;  // MN-major 32-bit operands: 32-byte-granular swizzle
        rc = make_tmap(&t.mapA[z], prec, ws + L.dZ[z][l], x.out[l], M, x.out[l], ch, bk, swz32);
        if (rc == CATB200_OK) rc = make_tmap(&t.mapB[z], prec, Hin, x.in_pad[l], M, x.in_pad[l], ch, bk, swz32);
        if (rc != CATB200_OK) return rc;
        t.gw[z] = reinterpret_cast<float*>(ws + L.gacc[z][l]);
        t.gb[z] = grads + P.b[z][l];
        const int seg = fold.n_segments++;
        fold.acc[seg] = t.gw[z];
        fold.grad[seg] = grads + P.w[z][l];
        fold.N[seg] = x.out[l]; fold.Kpad[seg] = x.in_pad[l]; fold.Ktrue[seg] = x.in[l];
      }
      t.outs = x.out[l]; t.ins_pad = x.in_pad[l]; t.rows = M; t.m_range = L.m_range[l];
      t.reverse = 0;
      if (l2_hints(M)) {
        // dgrad_l re-reads both operands next (dZ_l as A, H_{l-1} for ELU'); layer 0 has no dgrad.  CATB200_L2_HINTS=2: keep
        // only dZ_l (H_{l-1} of the 512-wide layer alone is half the L2)
        static int keep_h = -1;
        if (keep_h < 0) {
          const char* e = std::getenv("CATB200_L2_HINTS");
          keep_h = (e && e[0] == '2') ? 0 : 1;
        }
        t.hintA = l > 0 ? kL2EvictLast : kL2EvictFirst;
        t.hintB = l > 0 ? (keep_h ? kL2EvictLast : kL2EvictNormal) : kL2EvictFirst;
      }
      rc = tc_wgrad_launch(prec, t, L.splits[l], wst);
      if (rc != CATB200_OK) return rc;
    }
    if (l > 0) {  // dZ_{l-1} = (dZ_l W_l) * ELU'(H_{l-1}); A = dZ_l [M, out_l], B = W_l^T [in_l, out_l]
      TcGemmArgs t = {};
      for (int z = 0; z < 2; ++z) {
        rc = make_tmap(&t.mapA[z], prec, ws + L.dZ[z][l], x.out[l], M, x.out[l], ch, 128);
        if (rc == CATB200_OK) rc = make_tmap(&t.mapB[z], prec, wc + P.wtc[z][l] * es, x.out[l], x.in[l], x.out[l], ch, 128);
        if (rc == CATB200_OK) rc = make_tmap(&t.mapC[z], prec, ws + L.dZ[z][l - 1], x.in[l], M, x.in[l], ch, 32);
        if (rc == CATB200_OK) rc = make_tmap(&t.mapH[z], prec, ws + L.H[z][l - 1], x.in[l], M, x.in[l], ch, 32);
        if (rc != CATB200_OK) return rc;
      }
      t.M = M; t.N = x.in[l]; t.K = x.out[l];
      t.reverse = zigzag;
      if (l2_hints(M)) {  // last reads of dZ_l and H_{l-1}; dZ_{l-1} feeds the next two launches
        t.hintA = kL2EvictFirst; t.hintH = kL2EvictFirst; t.hintB = kL2EvictLast; t.hintC = kL2EvictLast;
      }
      rc = tc_gemm_launch(kTcDgrad, prec, t, st);
      if (rc != CATB200_OK) return rc;
    }
  }
  if (side) {  // join: the accumulators written on the side stream are folded next
    CATB200_CUDA_TRY(cudaEventRecord(side->done, side->stream));
    CATB200_CUDA_TRY(cudaStreamWaitEvent(st, side->done, 0));
  }
  // 5. fold the padded weight-gradient accumulators and the head kernel's per-CTA rows into the flat gradient
  fold.head_part = reinterpret_cast<const float*>(ws + L.head_part);
  fold.head_rows = L.head_rows;
  fold.gW4c = grads + P.w[0][3]; fold.gb4c = grads + P.b[0][3];
  fold.gW4a = grads + P.w[1][3]; fold.gb4a = grads + P.b[1][3];
  fold.glogstd = grads + P.logstd;
  fold.logstd = params + P.logstd; fold.loss_acc = loss_acc;
  fold.A = dims->act_dim; fold.h3 = dims->h3; fold.M = M;
  fold.ent_coef = hp->ent_coef; fold.vf_coef = hp->vf_coef;
  if (defer_fold) {
    *defer_fold = fold;
    return CATB200_OK;
  }
  // x: 76 CTAs cover the head segment (one per 32 values); the weight segments walk their quads with a grid stride
  CATB200_CUDA_TRY(launch_pdl(fold_grads_kernel, dim3(kHeadValues, fold.n_segments + 1), dim3(256), 0, st, fold));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// the optimizer-step arguments shared by the single-GPU and the peer kernel: fold description `o.f` already filled
static void fill_opt_step(const catb200_mlp_dims_t* dims, const AdamState& a, void* wcv, float max_grad_norm, int32_t* step_dev,
                          float* grad_norm_out, void* opt_ws, OptStepArgs& o) {
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  fill_adam_cast_args(dims, a, wcv, &o.c);
  for (int s = 0; s < o.f.n_segments; ++s) o.quad_prefix[s + 1] = o.quad_prefix[s] + o.f.N[s] * o.f.Kpad[s] / 4;
  for (int s = 0; s < 6; ++s) o.tile_prefix[s + 1] = o.tile_prefix[s] + (o.c.seg[s].rows / 32) * ((o.c.seg[s].cols_pad + 31) / 32);
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      o.bias_begin[z * 3 + l] = P.b[z][l];
      o.bias_len[z * 3 + l] = x.out[l];
    }
  for (int i = 0; i < o.c.n_rest_ranges; ++i) o.n_rest += o.c.rest_len[i];
  o.max_norm = max_grad_norm;
  o.step = step_dev;
  o.grad_norm_out = grad_norm_out;
  o.sc = static_cast<OptScratch*>(opt_ws);
}

// co-resident CTAs per SM for the grid-barrier kernels, at most 4: they are a few latency chains over 1.5 MB -- one CTA per
// SM took 31 us, more than the three launches it replaces
template <typename K>
static int coresident_per_sm(K k32, K k16) {
  int a32 = 0, a16 = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a32, k32, 256, 0) != cudaSuccess) return 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a16, k16, 256, 0) != cudaSuccess) return 1;
  return max(1, min(4, min(a32, a16)));
}

extern "C" {

int catb200_ppo_minibatch_grad(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                               const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                               const float* logprobs_all, const float* advantages_all, const float* returns_all,
                               const float* values_all, const float* norm_stats, const float* params, const void* wcv,
                               float* grads, float* loss_acc, void* workspace, size_t workspace_bytes, void* stream) {
  return minibatch_backward(dims, hp, mb_rows, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all,
                            values_all, norm_stats, params, wcv, grads, loss_acc, workspace, workspace_bytes, stream, nullptr);
}

int catb200_ppo_minibatch_update(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                                 const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                                 const float* logprobs_all, const float* advantages_all, const float* returns_all,
                                 const float* values_all, const float* norm_stats, float* params, void* wcv, float* grads,
                                 float* loss_acc, void* workspace, size_t workspace_bytes, float* exp_avg, float* exp_avg_sq,
                                 const float* lr_dev, int32_t* step_dev, float max_grad_norm, float beta1, float beta2,
                                 float eps, float grad_scale, float* grad_norm_out, void* opt_ws, void* stream) {
  if (!exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !opt_ws) return CATB200_ERR_INVALID_ARGUMENT;
  OptStepArgs o = {};
  int rc = minibatch_backward(dims, hp, mb_rows, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all,
                              values_all, norm_stats, params, wcv, grads, loss_acc, workspace, workspace_bytes, stream, &o.f);
  if (rc != CATB200_OK) return rc;
  AdamState a = {params, grads, exp_avg, exp_avg_sq, lr_dev, static_cast<OptScratch*>(opt_ws), beta1, beta2, eps, grad_scale};
  fill_opt_step(dims, a, wcv, max_grad_norm, step_dev, grad_norm_out, opt_ws, o);
  cudaStream_t st = as_stream(stream);
  // a plain (fully stream-ordered) launch of as many CTAs as are co-resident (the grid barrier needs that)
  static int per_sm = 0;
  if (per_sm == 0) per_sm = coresident_per_sm(opt_step_kernel<kPrecTf32>, opt_step_kernel<kPrecBf16>);
  if (dims->prec == kPrecTf32) opt_step_kernel<kPrecTf32><<<kNumSMs * per_sm, 256, 0, st>>>(o);
  else opt_step_kernel<kPrecBf16><<<kNumSMs * per_sm, 256, 0, st>>>(o);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_ppo_minibatch_update_peer(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                                      const int64_t* mb_inds, const void* obs_op_all, const float* actions_all,
                                      const float* logprobs_all, const float* advantages_all, const float* returns_all,
                                      const float* values_all, const float* norm_stats, float* params, void* wcv,
                                      float* loss_acc, void* workspace, size_t workspace_bytes, float* exp_avg,
                                      float* exp_avg_sq, const float* lr_dev, int32_t* step_dev, float max_grad_norm,
                                      float beta1, float beta2, float eps, float* grad_norm_out, void* opt_ws,
                                      void* const* peer_bases, int32_t rank, int32_t world, int32_t parity, float* grad_sum,
                                      uint32_t* epoch_dev, int32_t* err_dev, void* stream) {
  if (!exp_avg || !exp_avg_sq || !lr_dev || !step_dev || !opt_ws || !peer_bases || world < 1 || world > kPeerMax || rank < 0 ||
      rank >= world || !grad_sum || !epoch_dev || !err_dev || (parity != 0 && parity != 1) || !dims_ok(dims))
    return CATB200_ERR_INVALID_ARGUMENT;
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  const size_t n_pad = peer_n_pad(P.n_params);
  OptStepPeerArgs pa = {};
  for (int q = 0; q < world; ++q) {
    if (!peer_bases[q]) return CATB200_ERR_INVALID_ARGUMENT;
    char* base = static_cast<char*>(peer_bases[q]);
    pa.flags[q] = reinterpret_cast<uint32_t*>(base);
    pa.arena[q] = reinterpret_cast<const float*>(base + kFlagWords * 4) + (size_t)parity * n_pad;
  }
  float* own = reinterpret_cast<float*>(static_cast<char*>(peer_bases[rank]) + kFlagWords * 4);
  float* arena = own + (size_t)parity * n_pad;  // this minibatch's gradient accumulates here
  int rc = minibatch_backward(dims, hp, mb_rows, mb_inds, obs_op_all, actions_all, logprobs_all, advantages_all, returns_all,
                              values_all, norm_stats, params, wcv, arena, loss_acc, workspace, workspace_bytes, stream, &pa.o.f);
  if (rc != CATB200_OK) return rc;
  // exchange pattern: pull whole arenas (one handshake; default) or reduce-scatter / all-gather (two handshakes, 1 / world of
  // the traffic per peer; CATB200_PEER_RS=1).  Measured per optimizer step on B200s (profiles/README.md, round 2): 4 ranks
  // 38.8 us pull vs 43.6 us reduce-scatter, 8 ranks 65.7 vs 68.2 us -- at 1.5 MB the second handshake costs more than the
  // saved NVLink traffic.
  static int rs_mode = -1;
  if (rs_mode < 0) {
    const char* e = std::getenv("CATB200_PEER_RS");
    rs_mode = (e && e[0] == '1') ? 1 : 0;
  }
  pa.rs = rs_mode == 1;
  for (int q = 0; q < world; ++q)
    pa.gsum[q] = reinterpret_cast<float*>(static_cast<char*>(peer_bases[q]) + kFlagWords * 4) + 2 * n_pad;
  float* summed = pa.rs ? pa.gsum[rank] : grad_sum;  // what Adam consumes
  AdamState a = {params, summed, exp_avg, exp_avg_sq, lr_dev, static_cast<OptScratch*>(opt_ws), beta1, beta2, eps, 1.0f / (float)world};
  fill_opt_step(dims, a, wcv, max_grad_norm, step_dev, grad_norm_out, opt_ws, pa.o);
  pa.zero_arena = own + (size_t)(parity ^ 1) * n_pad;
  pa.out = grad_sum; pa.n = P.n_params; pa.rank = rank; pa.world = world;
  pa.epoch = epoch_dev; pa.err = err_dev; pa.parity = (unsigned int)parity;
  cudaStream_t st = as_stream(stream);
  static int per_sm = 0;
  if (per_sm == 0) per_sm = coresident_per_sm(opt_step_peer_kernel<kPrecTf32>, opt_step_peer_kernel<kPrecBf16>);
  if (dims->prec == kPrecTf32) opt_step_peer_kernel<kPrecTf32><<<kNumSMs * per_sm, 256, 0, st>>>(pa);
  else opt_step_peer_kernel<kPrecBf16><<<kNumSMs * per_sm, 256, 0, st>>>(pa);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

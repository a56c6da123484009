// Actor-critic MLP forward / backward and the PPO-clip loss on sm_100a.
//
// Reference: Agent (U/cleanrl/ppo.py:71-123) and the minibatch body of PPO() (ppo.py:298-352).
//
// Structure of one minibatch (M rows, both nets batched over blockIdx.z: 0 = critic, 1 = actor):
//   gather_kernel      : X[M, obs_pad] <- obs16_all[mb_inds] (128-byte rows) + advantage mean / unbiased std
//   gemm_nt (x3)       : H_l = ELU(H_{l-1} W_l^T + b_l), bf16 activations kept for the backward pass
//   head_loss_kernel   : fp32 heads (h3 -> act_dim / 1), Normal log-prob, PPO-clip + clipped value loss +
//                        entropy, their gradients w.r.t. the head weights / biases / log-std (atomics into
//                        the flat gradient) and dZ3 = dH3 * ELU'(H3) for both nets
//   wgrad (x3)         : dW_l = dZ_l^T H_{l-1}, split over M into per-CTA partial sums (plain stores)
//   gemm_nt dgrad (x2) : dZ_{l-1} = (dZ_l W_l) * ELU'(H_{l-1}), with the bias gradient (column sums of the
//                        fp32 result) folded into the epilogue
//   reduce_partials    : sums the split partials into the flat gradient (deterministic order)
// Tensor-core work is bf16 x bf16 -> fp32 (mma.sync m16n8k16 fed by cp.async + ldmatrix from XOR-swizzled
// shared memory, 3-stage pipeline); everything the reference does elementwise is fused into epilogues.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mma.cuh"
#include "tc_gemm.cuh"

namespace catb200 {

// GEMM backend: tcgen05/TMEM/TMA (default) or the mma.sync kernels below (CATB200_GEMM=mma), kept as the
// on-device cross-check of the tensor-core path.
static bool use_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_GEMM");
    v = (e && std::strcmp(e, "mma") == 0) ? 0 : 1;
  }
  return v == 1;
}

// fused three-layer forward kernel (tc_fwd3.cu) instead of three tc_gemm launches: CATB200_FUSED_FWD=1
static bool use_fused_fwd(int rows) {
  static int v = -1;  // row limit: 0 = never, INT_MAX = always ("1"), "r<N>" = only for at most N rows (rollout-sized launches)
  if (v < 0) {
    const char* e = std::getenv("CATB200_FUSED_FWD");
    v = 0;
    if (e && e[0] == '1') v = 0x7fffffff;
    if (e && e[0] == 'r') v = std::atoi(e + 1);
  }
  return rows <= v;
}

// Backward-pass overlap: the weight-gradient GEMM of layer l and the data-gradient GEMM that produces dZ_{l-1} both
// only read dZ_l, so they run concurrently -- wgrad on a library-owned side stream forked from / joined back into
// the caller's stream with events (inside a CUDA graph capture these become plain graph edges).  All work is still
// complete when the caller's stream reaches the end of the call.  Opt-in (CATB200_BWD_OVERLAP=1): measured +0.5 % at
// 4096 envs and +1.1 % at 16384 -- every GEMM of the chain already fills the machine for at least one wave, so only
// the tails overlap -- which does not pay for a hidden stream behind the C ABI.
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

constexpr int kGemmThreads = 256;
constexpr int kBM = 128, kBN = 128, kBK = 64;
constexpr int kStages = 3;
constexpr int kStageBytesNT = (kBM + kBN) * kBK * 2;  // 32 KiB
constexpr int kSmemNT = kStages * kStageBytesNT;       // 96 KiB

enum Epilogue { kEpiBiasElu = 0, kEpiMulDelu = 1 };

constexpr int kHeadThreads = 256;
constexpr int kMaxAct = 16;
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;
// layout of one head-kernel CTA's partial row: [value][lane]; values 0..63 = gW4a[j][f] (j = v/4, f = v%4)
constexpr int kHvW4c = 64, kHvB3c = 68, kHvB3a = 72, kHvB4a = 76, kHvLogstd = 77, kHvScalars = 78;
constexpr int kHeadValues = 84;
constexpr int kHeadSmem = (kHeadThreads / 32) * kHeadValues * 32 * 4;  // 84 KiB

struct GemmNTArgs {
  const bf16* A[2];  // [M, K] row-major, lda
  const bf16* B[2];  // [N, K] row-major, ldb
  bf16* C[2];        // [M, N] row-major, ldc
  const float* bias[2];  // kEpiBiasElu: [N]
  const bf16* H[2];      // kEpiMulDelu: forward activation [M, N] (ldc) whose ELU' scales the result
  float* dbias[2];       // kEpiMulDelu: += column sums of the scaled result
  int lda, ldb, ldc;
  int M, N, K;
};

// C = epi(A * B^T): CTA tile 128x128, 8 warps as 2 (m) x 4 (n), warp tile 64x32.
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads)
gemm_nt_kernel(const __grid_constant__ GemmNTArgs g) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int z = blockIdx.z;
  const int m_base = blockIdx.x * kBM, n_base = blockIdx.y * kBN;
  const bf16* __restrict__ A = g.A[z];
  const bf16* __restrict__ B = g.B[z];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const uint32_t smem0 = smem_u32(smem_raw);

  auto load_stage = [&](int stage, int k0) {
    const uint32_t sa = smem0 + stage * kStageBytesNT, sb = sa + kBM * kBK * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * kGemmThreads;
      const int row = idx >> 3, chunk = idx & 7;
      const int gm = m_base + row;
      cp_async16(sa + swz(row, chunk, 128), A + (size_t)min(gm, g.M - 1) * g.lda + k0 + chunk * 8, gm < g.M);
      cp_async16(sb + swz(row, chunk, 128), B + (size_t)(n_base + row) * g.ldb + k0 + chunk * 8, true);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

  const int kt_total = g.K / kBK;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < kt_total) load_stage(s, s * kBK);
    cp_async_commit();
  }
  for (int kt = 0; kt < kt_total; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    {  // prefetch tile kt + stages - 1 into the slot freed in the previous iteration
      const int nk = kt + kStages - 1;
      if (nk < kt_total) load_stage(nk % kStages, nk * kBK);
      cp_async_commit();
    }
    const uint32_t sa = smem0 + (kt % kStages) * kStageBytesNT, sb = sa + kBM * kBK * 2;
#pragma unroll
    for (int kk = 0; kk < kBK / 16; ++kk) {
      uint32_t af[4][4], bfr[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = wm * 64 + i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(af[i], sa + swz(row, kk * 2 + (lane >> 4), 128));
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int row = wn * 32 + jj * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4(bfr[jj], sb + swz(row, kk * 2 + ((lane >> 3) & 1), 128));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_16816(acc[i][j], af[i], bfr[j >> 1][(j & 1) * 2], bfr[j >> 1][(j & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue -----------------------------------------------------------------------------------
  const int gq = lane >> 2, tq = lane & 3;
  bf16* __restrict__ C = g.C[z];
  if (EPI == kEpiBiasElu) {
    const float* __restrict__ bias = g.bias[z];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n_base + wn * 32 + j * 8 + tq * 2;
      const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r0 = m_base + wm * 64 + i * 16 + gq;
        if (r0 < g.M)
          *reinterpret_cast<uint32_t*>(C + (size_t)r0 * g.ldc + col) = pack_bf16x2(elu(acc[i][j][0] + b0), elu(acc[i][j][1] + b1));
        if (r0 + 8 < g.M)
          *reinterpret_cast<uint32_t*>(C + (size_t)(r0 + 8) * g.ldc + col) = pack_bf16x2(elu(acc[i][j][2] + b0), elu(acc[i][j][3] + b1));
      }
    }
  } else {
    const bf16* __restrict__ H = g.H[z];
    float* __restrict__ dbias = g.dbias[z];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n_base + wn * 32 + j * 8 + tq * 2;
      float cs0 = 0.0f, cs1 = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r0 = m_base + wm * 64 + i * 16 + gq;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = r0 + half * 8;
          if (r < g.M) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(H + (size_t)r * g.ldc + col);
            const float v0 = acc[i][j][half * 2] * elu_grad_from_output(__low2float(h));
            const float v1 = acc[i][j][half * 2 + 1] * elu_grad_from_output(__high2float(h));
            *reinterpret_cast<uint32_t*>(C + (size_t)r * g.ldc + col) = pack_bf16x2(v0, v1);
            cs0 += v0;
            cs1 += v1;
          }
        }
      }
      // column sums over the warp's 64 rows: reduce across the 8 row groups (lane bits 2..4)
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        cs0 += __shfl_xor_sync(0xffffffffu, cs0, o);
        cs1 += __shfl_xor_sync(0xffffffffu, cs1, o);
      }
      if (gq == 0) {
        atomicAdd(dbias + col, cs0);
        atomicAdd(dbias + col + 1, cs1);
      }
    }
  }
}

// ---- weight gradient: dW[N, K] = dZ[M, N]^T * Hin[M, K], split over M ------------------------------
struct WgradArgs {
  const bf16* dZ[2];   // [M, N] row-major, ld = N
  const bf16* Hin[2];  // [M, Kpad] row-major, ld = Kpad
  float* part[2];      // [splits, N, Kpad] fp32 partial sums
  int M, N, Kpad, m_range;  // rows of M handled per split (multiple of 64)
};

constexpr int kWgBM = 64;  // reduction rows per pipeline stage

template <int KT>
__global__ void __launch_bounds__(kGemmThreads)
wgrad_kernel(const __grid_constant__ WgradArgs g) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  constexpr int kRowBytesZ = 128 * 2, kRowBytesH = KT * 2;
  constexpr int kStageBytes = kWgBM * (kRowBytesZ + kRowBytesH);
  constexpr int NJ = KT / 32;  // n8 tiles per warp along the input-feature axis
  const int z = blockIdx.z;
  const int k_tiles = g.Kpad / KT;
  const int n_base = (blockIdx.x / k_tiles) * 128, k_base = (blockIdx.x % k_tiles) * KT;
  const int m_begin = blockIdx.y * g.m_range, m_end = min(g.M, m_begin + g.m_range);
  const bf16* __restrict__ dZ = g.dZ[z];
  const bf16* __restrict__ Hin = g.Hin[z];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const uint32_t smem0 = smem_u32(smem_raw);

  auto load_stage = [&](int stage, int m0) {
    const uint32_t sz = smem0 + stage * kStageBytes, sh = sz + kWgBM * kRowBytesZ;
#pragma unroll
    for (int i = 0; i < (kWgBM * 16) / kGemmThreads; ++i) {  // dZ tile: 64 rows x 16 chunks
      const int idx = tid + i * kGemmThreads;
      const int row = idx >> 4, chunk = idx & 15;
      const int gm = m0 + row;
      cp_async16(sz + swz(row, chunk, kRowBytesZ), dZ + (size_t)min(gm, g.M - 1) * g.N + n_base + chunk * 8, gm < m_end);
    }
    constexpr int kChunksH = KT / 8;
#pragma unroll
    for (int i = 0; i < (kWgBM * kChunksH) / kGemmThreads; ++i) {
      const int idx = tid + i * kGemmThreads;
      const int row = idx / kChunksH, chunk = idx % kChunksH;
      const int gm = m0 + row;
      cp_async16(sh + swz(row, chunk, kRowBytesH), Hin + (size_t)min(gm, g.M - 1) * g.Kpad + k_base + chunk * 8, gm < m_end);
    }
  };

  float acc[4][NJ][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

  const int mt_total = (max(m_end - m_begin, 0) + kWgBM - 1) / kWgBM;
#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < mt_total) load_stage(s, m_begin + s * kWgBM);
    cp_async_commit();
  }
  for (int mt = 0; mt < mt_total; ++mt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    {
      const int nm = mt + kStages - 1;
      if (nm < mt_total) load_stage(nm % kStages, m_begin + nm * kWgBM);
      cp_async_commit();
    }
    const uint32_t sz = smem0 + (mt % kStages) * kStageBytes, sh = sz + kWgBM * kRowBytesZ;
#pragma unroll
    for (int kk = 0; kk < kWgBM / 16; ++kk) {
      uint32_t af[4][4], bfr[NJ / 2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // A = dZ^T: stored [m][n], transposed on load
        const int row = kk * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = (wm * 64 + i * 16) / 8 + ((lane >> 3) & 1);
        ldmatrix_x4_trans(af[i], sz + swz(row, chunk, kRowBytesZ));
      }
#pragma unroll
      for (int jj = 0; jj < NJ / 2; ++jj) {  // B = Hin: stored [m][k], transposed on load
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = (wn * (KT / 4) + jj * 16) / 8 + (lane >> 4);
        ldmatrix_x4_trans(bfr[jj], sh + swz(row, chunk, kRowBytesH));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma_bf16_16816(acc[i][j], af[i], bfr[j >> 1][(j & 1) * 2], bfr[j >> 1][(j & 1) * 2 + 1]);
    }
  }
  cp_async_wait<0>();

  const int gq = lane >> 2, tq = lane & 3;
  float* __restrict__ part = g.part[z] + (size_t)blockIdx.y * g.N * g.Kpad;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n0 = n_base + wm * 64 + i * 16 + gq;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k0 = k_base + wn * (KT / 4) + j * 8 + tq * 2;
      *reinterpret_cast<float2*>(part + (size_t)n0 * g.Kpad + k0) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(part + (size_t)(n0 + 8) * g.Kpad + k0) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
  }
}

// grads[n, k] += sum_s part[s, n, k] for k < Ktrue (the padded input columns of layer 0 are dropped).
// blockIdx.y selects the segment; the last segment is the head kernel's per-CTA rows.
struct ReduceArgs {
  const float* part[6];
  float* grad[6];
  int N[6], Kpad[6], Ktrue[6], splits[6];
  int n_segments;
  // head segment
  const float* head_part; int head_rows;
  float* gW4c; float* gb4c; float* gW4a; float* gb4a; float* glogstd; float* gb3[2];
  const float* logstd; float* loss_acc;
  int A, h3, M;
  float ent_coef, vf_coef;
};

__global__ void __launch_bounds__(256) reduce_partials_kernel(const __grid_constant__ ReduceArgs r) {
  pdl_launch_dependents();
  pdl_wait();
  const int seg = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg < r.n_segments) {
    const int N = r.N[seg], Kpad = r.Kpad[seg], Kt = r.Ktrue[seg], S = r.splits[seg];
    const int total = N * Kpad;
    if (e >= total) return;
    const int n = e / Kpad, k = e - n * Kpad;
    if (k >= Kt) return;
    const float* __restrict__ p = r.part[seg] + e;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    int q = 0;
    for (; q + 4 <= S; q += 4) {
      s0 += __ldcs(p + (size_t)q * total);
      s1 += __ldcs(p + (size_t)(q + 1) * total);
      s2 += __ldcs(p + (size_t)(q + 2) * total);
      s3 += __ldcs(p + (size_t)(q + 3) * total);
    }
    for (; q < S; ++q) s0 += __ldcs(p + (size_t)q * total);
    r.grad[seg][(size_t)n * Kt + k] += (s0 + s1) + (s2 + s3);
    return;
  }
  // ---- head segment: sum the per-CTA rows, route every value to its gradient slot / loss accumulator
  if (e >= kHeadValues * 32) return;
  float t = 0.0f;
  for (int c = 0; c < r.head_rows; ++c) t += r.head_part[(size_t)c * kHeadValues * 32 + e];
  const int v = e >> 5, lane = e & 31;
  const float inv_M = 1.0f / (float)r.M;
  if (v < kHvW4c) {
    const int j = v >> 2, f = v & 3;
    if (j < r.A) r.gW4a[j * r.h3 + lane * 4 + f] += t;
  } else if (v < kHvB3c) {
    r.gW4c[lane * 4 + (v - kHvW4c)] += t;
  } else if (v < kHvB3a) {
    r.gb3[0][lane * 4 + (v - kHvB3c)] += t;
  } else if (v < kHvB4a) {
    r.gb3[1][lane * 4 + (v - kHvB3a)] += t;
  } else if (v == kHvB4a) {  // per-action-dim scalars live in the even lane of pair (2j, 2j+1)
    if ((lane & 1) == 0 && (lane >> 1) < r.A) r.gb4a[lane >> 1] += t;
  } else if (v == kHvLogstd) {
    // policy part + d(-ent_coef * mean entropy)/d logstd_j = -ent_coef (entropy is sample independent)
    if ((lane & 1) == 0 && (lane >> 1) < r.A) r.glogstd[lane >> 1] += t - r.ent_coef;
  } else if (lane == 0) {
    const int k = v - kHvScalars;  // g_b4c, pg, v, kl, clip, old_kl
    if (k == 0) {
      r.gb4c[0] += t;
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
}

// ---- minibatch gather + advantage statistics -----------------------------------------------------------
struct MbStats {
  float adv_mean, adv_std;  // unbiased std (ppo.py:316-318)
  unsigned int ticket, pad;
  double sum, sumsq;
};

__global__ void __launch_bounds__(256)
gather_kernel(const int64_t* __restrict__ mb_inds, int M, const bf16* __restrict__ obs16_all, int obs_pad,
              const float* __restrict__ adv_all, const float* __restrict__ logp_all, const float* __restrict__ ret_all,
              const float* __restrict__ val_all, const float* __restrict__ act_all, int A, bf16* __restrict__ X,
              float4* __restrict__ scal_mb, float* __restrict__ act_mb, MbStats* __restrict__ st) {
  pdl_launch_dependents();
  pdl_wait();
  const int chunks = obs_pad / 8;  // 16-byte chunks per row
  const int total = M * chunks;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int m = e / chunks, c = e - m * chunks;
    const int64_t src = mb_inds[m];
    reinterpret_cast<uint4*>(X)[e] = __ldg(reinterpret_cast<const uint4*>(obs16_all + (size_t)src * obs_pad) + c);
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
  if (last_block_ticket(&st->ticket, gridDim.x)) {
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
  const bf16* H3[2];   // [M, h3] activations of the last hidden layer (0 critic, 1 actor)
  bf16* dZ3[2];        // [M, h3] out (training): gradient w.r.t. the pre-activation of that layer
  const float* W4c; const float* b4c;  // critic head [1, h3], [1]
  const float* W4a; const float* b4a;  // actor head  [A, h3], [A]
  const float* logstd;                 // [A]
  int M, h3, A;
  // rollout outputs / inputs
  const float* noise; const float* action_in; float* action; float* logprob; float* value; float* mean_out;
  // training inputs, packed in minibatch order by gather_kernel
  const float4* scal_mb;  // {old log-prob, advantage, return, old value}
  const float* act_mb;    // [M, A]
  const float* norm_stats; const MbStats* mb;
  catb200_ppo_hparams_t hp;
  // training outputs
  float* head_part;  // [gridDim.x][kHeadValues][32] per-CTA partial sums (training)
};

// One warp per sample; lane owns features lane*4 .. lane*4+3 of each 128-wide slice of h3.
template <bool TRAIN>
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
  float gw4a[kMaxAct][F], gw4c[F], gb3c[F], gb3a[F];
  float g_b4a = 0.0f, g_logstd = 0.0f, g_b4c = 0.0f;
  float l_pg = 0.0f, l_v = 0.0f, l_kl = 0.0f, l_clip = 0.0f, l_oldkl = 0.0f;
  if (TRAIN) {
#pragma unroll
    for (int f = 0; f < F; ++f) {
      gw4c[f] = 0.0f;
      gb3c[f] = 0.0f;
      gb3a[f] = 0.0f;
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

  // software pipeline: the inputs of the next sample are in flight while the current one is processed
  const int m_stride = gridDim.x * warps;
  int m = blockIdx.x * warps + warp;
  uint2 n_rc = make_uint2(0, 0), n_ra = make_uint2(0, 0);
  float4 n_sc = make_float4(0.f, 0.f, 0.f, 0.f);
  float n_act = 0.0f;
  auto fetch = [&](int mm) {
    n_rc = __ldg(reinterpret_cast<const uint2*>(a.H3[0] + (size_t)mm * a.h3) + lane);
    n_ra = __ldg(reinterpret_cast<const uint2*>(a.H3[1] + (size_t)mm * a.h3) + lane);
    if (TRAIN) {
      n_sc = __ldg(a.scal_mb + mm);
      if (owner) n_act = __ldg(a.act_mb + (size_t)mm * A + aj);
    }
  };
  if (m < a.M) fetch(m);
  for (; m < a.M; m += m_stride) {
    // ---- the two 128-wide activation rows (8 bytes per lane each, coalesced) + scalars of this sample
    float hc[F], ha[F];
    const float4 sc = n_sc;
    const float act_in = n_act;
    {
      const uint2 rc = n_rc, ra = n_ra;
      const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&rc);
      const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ra);
      hc[0] = __low2float(pc[0]); hc[1] = __high2float(pc[0]); hc[2] = __low2float(pc[1]); hc[3] = __high2float(pc[1]);
      ha[0] = __low2float(pa[0]); ha[1] = __high2float(pa[0]); ha[2] = __low2float(pa[1]); ha[3] = __high2float(pa[1]);
    }
    if (m + m_stride < a.M) fetch(m + m_stride);
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
          const float eps = a.noise ? __ldg(a.noise + (size_t)m * A + aj) : 0.0f;
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
      gb3a[f] += dza[f];
      gb3c[f] += dzc[f];
    }
    uint2 oa, oc;
    oa.x = pack_bf16x2(dza[0], dza[1]); oa.y = pack_bf16x2(dza[2], dza[3]);
    oc.x = pack_bf16x2(dzc[0], dzc[1]); oc.y = pack_bf16x2(dzc[2], dzc[3]);
    reinterpret_cast<uint2*>(a.dZ3[1] + (size_t)m * a.h3)[lane] = oa;
    reinterpret_cast<uint2*>(a.dZ3[0] + (size_t)m * a.h3)[lane] = oc;
  }
  if (!TRAIN) return;

  // ---- CTA-level reduction: every warp parks its accumulators in shared memory ([value][lane] rows),
  // one barrier, then the CTA sums over its warps and writes ONE partial row per CTA (plain coalesced
  // stores, no atomics).  head_reduce (inside reduce_partials_kernel) folds the rows into the gradient.
  extern __shared__ float hsm[];  // [warps][kHeadValues][32]
  float* mine = hsm + (size_t)warp * kHeadValues * 32;
#pragma unroll
  for (int j = 0; j < kMaxAct; ++j)
#pragma unroll
    for (int f = 0; f < F; ++f) mine[(j * F + f) * 32 + lane] = gw4a[j][f];
#pragma unroll
  for (int f = 0; f < F; ++f) {
    mine[(kHvW4c + f) * 32 + lane] = gw4c[f];
    mine[(kHvB3c + f) * 32 + lane] = gb3c[f];
    mine[(kHvB3a + f) * 32 + lane] = gb3a[f];
  }
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

// ---- small utility kernels ------------------------------------------------------------------------------
__global__ void obs_to_bf16_kernel(const float* __restrict__ obs, long long rows, int dim, int pad, bf16* __restrict__ out) {
  const long long total = rows * pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / pad;
    const int c = (int)(e - r * pad);
    out[e] = __float2bfloat16(c < dim ? obs[r * dim + c] : 0.0f);
  }
}

struct CastSeg {
  const float* src; bf16* dst; bf16* dst_t;
  int rows, cols, cols_pad;  // src [rows, cols] -> dst [rows, cols_pad] and dst_t [cols, rows]
};
struct CastArgs { CastSeg seg[6]; };

// 32x32 tiles: coalesced fp32 reads, coalesced bf16 writes of W, and a shared-memory transpose for W^T.
// grid = (max tiles over segments, 6 segments), block = (32, 8).
__global__ void __launch_bounds__(256) cast_weights_kernel(const __grid_constant__ CastArgs c) {
  pdl_launch_dependents();
  pdl_wait();
  const CastSeg& s = c.seg[blockIdx.y];
  const int tiles_c = (s.cols_pad + 31) / 32, tiles_r = s.rows / 32;
  if ((int)blockIdx.x >= tiles_c * tiles_r) return;
  const int tr = blockIdx.x / tiles_c, tc = blockIdx.x % tiles_c;
  __shared__ float tile[32][33];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tr * 32 + threadIdx.y + i * 8, k = tc * 32 + threadIdx.x;
    const float v = k < s.cols ? s.src[(size_t)r * s.cols + k] : 0.0f;
    tile[threadIdx.y + i * 8][threadIdx.x] = v;
    if (k < s.cols_pad) s.dst[(size_t)r * s.cols_pad + k] = __float2bfloat16(v);
  }
  if (s.dst_t == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = tc * 32 + threadIdx.y + i * 8, r = tr * 32 + threadIdx.x;
    if (k < s.cols) s.dst_t[(size_t)k * s.rows + r] = __float2bfloat16(tile[threadIdx.x][threadIdx.y + i * 8]);
  }
}

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
  return true;
}

static Dims make_dims(const catb200_mlp_dims_t* d) {
  Dims x;
  x.in[0] = d->obs_dim; x.in_pad[0] = d->obs_pad; x.out[0] = d->h1;
  x.in[1] = d->h1; x.in_pad[1] = d->h1; x.out[1] = d->h2;
  x.in[2] = d->h2; x.in_pad[2] = d->h2; x.out[2] = d->h3;
  return x;
}

struct ActLayout {  // byte offsets into the activation workspace
  size_t X, H[2][3], dZ[2][3], mb, part[2][3], head_part, scal_mb, act_mb, total;
  int splits[3], m_range[3], head_rows;
};

static ActLayout act_layout(const catb200_mlp_dims_t* d, int rows, bool training) {
  ActLayout L = {};
  Dims x = make_dims(d);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
  L.mb = take(sizeof(MbStats));
  L.X = take((size_t)rows * d->obs_pad * 2);
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) L.H[z][l] = take((size_t)rows * x.out[l] * 2);
  if (training) {
    L.head_rows = min((rows + 7) / 8, kNumSMs);
    L.head_part = take((size_t)L.head_rows * kHeadValues * 32 * 4);
    L.scal_mb = take((size_t)rows * 16);
    L.act_mb = take((size_t)rows * d->act_dim * 4);
    for (int z = 0; z < 2; ++z)
      for (int l = 0; l < 3; ++l) L.dZ[z][l] = take((size_t)rows * x.out[l] * 2);
    for (int l = 0; l < 3; ++l) {
      const int kt = x.in_pad[l] >= 128 ? 128 : 64;
      const int tiles = (x.out[l] / 128) * (x.in_pad[l] / kt) * 2;
      int want = max(1, kNumSMs / tiles);  // ~one CTA per SM per layer: fewer, fatter splits = fewer partial bytes
      int m_range = ((rows + want - 1) / want + kWgBM - 1) / kWgBM * kWgBM;
      m_range = max(m_range, kWgBM);
      L.m_range[l] = m_range;
      L.splits[l] = (rows + m_range - 1) / m_range;
      for (int z = 0; z < 2; ++z) L.part[z][l] = take((size_t)L.splits[l] * x.out[l] * x.in_pad[l] * 4);
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
  int64_t o16 = 0;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      L->w16[z][l] = o16; o16 += (int64_t)x.out[l] * x.in_pad[l];
      if (l > 0) { L->wt16[z][l] = o16; o16 += (int64_t)x.in[l] * x.out[l]; }
      else L->wt16[z][l] = -1;
    }
  L->n_w16 = o16;
  return CATB200_OK;
}

static int launch_forward(const catb200_mlp_dims_t* d, const catb200_mlp_layout_t& P, const ActLayout& L, const bf16* X,
                          int rows, const float* params, const bf16* w16, char* ws, cudaStream_t st) {
  Dims x = make_dims(d);
  static bool attr_set = false;
  if (!attr_set) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<kEpiBiasElu>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNT));
    CATB200_CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel<kEpiMulDelu>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNT));
    attr_set = true;
  }
  if (use_tc() && use_fused_fwd(rows) && d->obs_pad == 64 && d->h1 <= 512 && d->h2 <= 256 && d->h3 == 128) {
    Fwd3Args f = {};
    for (int z = 0; z < 2; ++z) {
      int rc = make_tmap_bf16(&f.mapX[z], X, 64, rows, 64, 64, 128);
      for (int l = 0; l < 3 && rc == CATB200_OK; ++l) {
        rc = make_tmap_bf16(&f.mapW[z][l], w16 + P.w16[z][l], x.in_pad[l], x.out[l], x.in_pad[l], 64, 128);
        if (rc == CATB200_OK) rc = make_tmap_bf16(&f.mapH[z][l], ws + L.H[z][l], x.out[l], rows, x.out[l], 64, 32);
        f.bias[z][l] = params + P.b[z][l];
      }
      if (rc != CATB200_OK) return rc;
    }
    f.M = rows; f.h1 = d->h1; f.h2 = d->h2; f.h3 = d->h3;
    return fwd3_launch(f, st);
  }
  for (int l = 0; l < 3 && use_tc(); ++l) {
    TcGemmArgs t = {};
    for (int z = 0; z < 2; ++z) {
      const bf16* A = l == 0 ? X : reinterpret_cast<const bf16*>(ws + L.H[z][l - 1]);
      int rc = make_tmap_bf16(&t.mapA[z], A, x.in_pad[l], rows, x.in_pad[l], 64, 128);
      if (rc == CATB200_OK) rc = make_tmap_bf16(&t.mapB[z], w16 + P.w16[z][l], x.in_pad[l], x.out[l], x.in_pad[l], 64, 128);
      if (rc != CATB200_OK) return rc;
      t.C[z] = reinterpret_cast<bf16*>(ws + L.H[z][l]);
      if (rc == CATB200_OK) rc = make_tmap_bf16(&t.mapC[z], t.C[z], x.out[l], rows, x.out[l], 64, 32);
      if (rc != CATB200_OK) return rc;
      t.bias[z] = params + P.b[z][l];
    }
    t.ldc = x.out[l]; t.M = rows; t.N = x.out[l]; t.K = x.in_pad[l];
    int rc = tc_gemm_launch(kTcFwd, t, 1, st);
    if (rc != CATB200_OK) return rc;
  }
  for (int l = 0; l < 3 && !use_tc(); ++l) {
    GemmNTArgs g = {};
    for (int z = 0; z < 2; ++z) {
      g.A[z] = l == 0 ? X : reinterpret_cast<const bf16*>(ws + L.H[z][l - 1]);
      g.B[z] = w16 + P.w16[z][l];
      g.C[z] = reinterpret_cast<bf16*>(ws + L.H[z][l]);
      g.bias[z] = params + P.b[z][l];
    }
    g.lda = x.in_pad[l]; g.ldb = x.in_pad[l]; g.ldc = x.out[l];
    g.M = rows; g.N = x.out[l]; g.K = x.in_pad[l];
    dim3 grid((rows + kBM - 1) / kBM, x.out[l] / kBN, 2);
    CATB200_CUDA_TRY(launch_pdl(gemm_nt_kernel<kEpiBiasElu>, grid, dim3(kGemmThreads), kSmemNT, st, g));
    CATB200_LAUNCH_CHECK();
  }
  return CATB200_OK;
}

int launch_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* w16v, cudaStream_t st) {
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  bf16* w16 = static_cast<bf16*>(w16v);
  CastArgs c;
  int max_tiles = 1;
  for (int z = 0; z < 2; ++z)
    for (int l = 0; l < 3; ++l) {
      CastSeg& s = c.seg[z * 3 + l];
      s.src = params + P.w[z][l];
      s.dst = w16 + P.w16[z][l];
      s.dst_t = l > 0 ? w16 + P.wt16[z][l] : nullptr;
      s.rows = x.out[l]; s.cols = x.in[l]; s.cols_pad = x.in_pad[l];
      max_tiles = max(max_tiles, (s.rows / 32) * ((s.cols_pad + 31) / 32));
    }
  CATB200_CUDA_TRY(launch_pdl(cast_weights_kernel, dim3(max_tiles, 6), dim3(32, 8), 0, st, c));
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

int catb200_mlp_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* w16v, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!params || !w16v) return CATB200_ERR_INVALID_ARGUMENT;
  return launch_cast_weights(dims, params, w16v, as_stream(stream));
}

int catb200_obs_to_bf16(const float* obs, int64_t rows, int32_t obs_dim, int32_t obs_pad, void* obs16, void* stream) {
  if (!obs || !obs16 || rows <= 0 || obs_dim <= 0 || obs_pad < obs_dim) return CATB200_ERR_INVALID_ARGUMENT;
  const long long total = rows * obs_pad;
  const int grid = (int)min((total + 255) / 256, (long long)kNumSMs * 16);
  obs_to_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(obs, rows, obs_dim, obs_pad, static_cast<bf16*>(obs16));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

size_t catb200_mlp_workspace_bytes(const catb200_mlp_dims_t* dims, int32_t rows, int32_t training) {
  if (!dims_ok(dims) || rows <= 0) return 0;
  return act_layout(dims, rows, training != 0).total;
}

int catb200_mlp_act(const catb200_mlp_dims_t* dims, const void* obs16, int32_t rows, const float* params, const void* w16,
                    const float* noise, const float* action_in, float* action, float* logprob, float* value,
                    float* mean_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!obs16 || rows <= 0 || !params || !w16 || !workspace) return CATB200_ERR_INVALID_ARGUMENT;
  const ActLayout L = act_layout(dims, rows, false);
  if (workspace_bytes < L.total) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  int rc = launch_forward(dims, P, L, static_cast<const bf16*>(obs16), rows, params, static_cast<const bf16*>(w16), ws, st);
  if (rc != CATB200_OK) return rc;
  HeadArgs a = {};
  for (int z = 0; z < 2; ++z) a.H3[z] = reinterpret_cast<const bf16*>(ws + L.H[z][2]);
  a.W4c = params + P.w[0][3]; a.b4c = params + P.b[0][3];
  a.W4a = params + P.w[1][3]; a.b4a = params + P.b[1][3];
  a.logstd = params + P.logstd;
  a.M = rows; a.h3 = dims->h3; a.A = dims->act_dim;
  a.noise = noise; a.action_in = action_in; a.action = action; a.logprob = logprob; a.value = value; a.mean_out = mean_out;
  const int grid = min((rows + 7) / 8, kNumSMs * 4);
  CATB200_CUDA_TRY(launch_pdl(head_kernel<false>, dim3(grid), dim3(kHeadThreads), 0, st, a));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_ppo_minibatch_grad(const catb200_mlp_dims_t* dims, const catb200_ppo_hparams_t* hp, int32_t mb_rows,
                               const int64_t* mb_inds, const void* obs16_all, const float* actions_all,
                               const float* logprobs_all, const float* advantages_all, const float* returns_all,
                               const float* values_all, const float* norm_stats, const float* params, const void* w16v,
                               float* grads, float* loss_acc, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dims_ok(dims)) return CATB200_ERR_UNSUPPORTED;
  if (!hp || mb_rows <= 0 || !mb_inds || !obs16_all || !actions_all || !logprobs_all || !advantages_all || !returns_all ||
      !values_all || !norm_stats || !params || !w16v || !grads || !loss_acc || !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  const int M = mb_rows;
  const ActLayout L = act_layout(dims, M, true);
  if (workspace_bytes < L.total) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  catb200_mlp_layout_t P;
  fill_layout(dims, &P);
  Dims x = make_dims(dims);
  cudaStream_t st = as_stream(stream);
  char* ws = static_cast<char*>(workspace);
  const bf16* w16 = static_cast<const bf16*>(w16v);
  bf16* X = reinterpret_cast<bf16*>(ws + L.X);
  MbStats* mb = reinterpret_cast<MbStats*>(ws + L.mb);

  // 1. gather + advantage statistics
  CATB200_CUDA_TRY(launch_pdl(gather_kernel, dim3(min((M * (dims->obs_pad / 8) + 255) / 256, kNumSMs * 8)), dim3(256), 0, st,
                              mb_inds, M, static_cast<const bf16*>(obs16_all), (int)dims->obs_pad, advantages_all, logprobs_all,
                              returns_all, values_all, actions_all, (int)dims->act_dim, X, reinterpret_cast<float4*>(ws + L.scal_mb),
                              reinterpret_cast<float*>(ws + L.act_mb), mb));
  CATB200_LAUNCH_CHECK();
  // 2. forward through the three hidden layers of both nets
  int rc = launch_forward(dims, P, L, X, M, params, w16, ws, st);
  if (rc != CATB200_OK) return rc;
  // 3. heads + loss + dZ3
  {
    HeadArgs a = {};
    for (int z = 0; z < 2; ++z) {
      a.H3[z] = reinterpret_cast<const bf16*>(ws + L.H[z][2]);
      a.dZ3[z] = reinterpret_cast<bf16*>(ws + L.dZ[z][2]);
    }
    a.W4c = params + P.w[0][3]; a.b4c = params + P.b[0][3];
    a.W4a = params + P.w[1][3]; a.b4a = params + P.b[1][3];
    a.logstd = params + P.logstd;
    a.M = M; a.h3 = dims->h3; a.A = dims->act_dim;
    a.scal_mb = reinterpret_cast<const float4*>(ws + L.scal_mb);
    a.act_mb = reinterpret_cast<const float*>(ws + L.act_mb);
    a.norm_stats = norm_stats; a.mb = mb; a.hp = *hp;
    a.head_part = reinterpret_cast<float*>(ws + L.head_part);
    static bool head_attr = false;
    if (!head_attr) {
      CATB200_CUDA_TRY(cudaFuncSetAttribute(head_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHeadSmem));
      head_attr = true;
    }
    CATB200_CUDA_TRY(launch_pdl(head_kernel<true>, dim3(L.head_rows), dim3(kHeadThreads), kHeadSmem, st, a));
    CATB200_LAUNCH_CHECK();
  }
  // 4. backward through the hidden layers
  static bool wg_attr = false;
  if (!wg_attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kWgBM * (256 + 256)));
    CATB200_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kWgBM * (256 + 128)));
    wg_attr = true;
  }
  ReduceArgs red = {};
  SideStream* side = use_tc() ? side_stream() : nullptr;
  for (int l = 2; l >= 0; --l) {
    // dZ_l is complete on `st` here (head kernel or the previous dgrad): the weight gradient may start beside dgrad
    cudaStream_t wst = st;
    if (side && l > 0) {
      CATB200_CUDA_TRY(cudaEventRecord(side->dz_ready[l], st));
      CATB200_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->dz_ready[l], 0));
      wst = side->stream;
    }
    WgradArgs wgt = {};
    for (int z = 0; z < 2; ++z) {
      wgt.dZ[z] = reinterpret_cast<const bf16*>(ws + L.dZ[z][l]);
      wgt.Hin[z] = l == 0 ? X : reinterpret_cast<const bf16*>(ws + L.H[z][l - 1]);
      wgt.part[z] = reinterpret_cast<float*>(ws + L.part[z][l]);
      const int seg = red.n_segments++;
      red.part[seg] = wgt.part[z];
      red.grad[seg] = grads + P.w[z][l];
      red.N[seg] = x.out[l]; red.Kpad[seg] = x.in_pad[l]; red.Ktrue[seg] = x.in[l]; red.splits[seg] = L.splits[l];
    }
    wgt.M = M; wgt.N = x.out[l]; wgt.Kpad = x.in_pad[l]; wgt.m_range = L.m_range[l];
    if (use_tc()) {
      TcGemmArgs t = {};
      for (int z = 0; z < 2; ++z) {
        int rc = make_tmap_bf16(&t.mapA[z], wgt.dZ[z], x.out[l], M, x.out[l], 64, 64);
        if (rc == CATB200_OK) rc = make_tmap_bf16(&t.mapB[z], wgt.Hin[z], x.in_pad[l], M, x.in_pad[l], 64, 64);
        if (rc != CATB200_OK) return rc;
        t.part[z] = wgt.part[z];
      }
      t.M = x.out[l]; t.N = x.in_pad[l]; t.K = M; t.m_range = L.m_range[l];
      int rc = tc_gemm_launch(kTcWgrad, t, L.splits[l], wst);
      if (rc != CATB200_OK) return rc;
    } else if (x.in_pad[l] >= 128) {
      dim3 grid((x.out[l] / 128) * (x.in_pad[l] / 128), L.splits[l], 2);
      CATB200_CUDA_TRY(launch_pdl(wgrad_kernel<128>, grid, dim3(kGemmThreads), kStages * kWgBM * (256 + 256), st, wgt));
      CATB200_LAUNCH_CHECK();
    } else {
      dim3 grid((x.out[l] / 128) * (x.in_pad[l] / 64), L.splits[l], 2);
      CATB200_CUDA_TRY(launch_pdl(wgrad_kernel<64>, grid, dim3(kGemmThreads), kStages * kWgBM * (256 + 128), st, wgt));
      CATB200_LAUNCH_CHECK();
    }
    if (l > 0 && use_tc()) {  // dZ_{l-1} = (dZ_l W_l) * ELU'(H_{l-1}); A = dZ_l [M, out_l], B = W_l^T [in_l, out_l]
      TcGemmArgs t = {};
      for (int z = 0; z < 2; ++z) {
        int rc = make_tmap_bf16(&t.mapA[z], reinterpret_cast<const bf16*>(ws + L.dZ[z][l]), x.out[l], M, x.out[l], 64, 128);
        if (rc == CATB200_OK) rc = make_tmap_bf16(&t.mapB[z], w16 + P.wt16[z][l], x.out[l], x.in[l], x.out[l], 64, 128);
        if (rc != CATB200_OK) return rc;
        t.C[z] = reinterpret_cast<bf16*>(ws + L.dZ[z][l - 1]);
        if (rc == CATB200_OK) rc = make_tmap_bf16(&t.mapC[z], t.C[z], x.in[l], M, x.in[l], 64, 32);
        if (rc != CATB200_OK) return rc;
        t.H[z] = reinterpret_cast<const bf16*>(ws + L.H[z][l - 1]);
        t.dbias[z] = grads + P.b[z][l - 1];
      }
      t.ldc = x.in[l]; t.M = M; t.N = x.in[l]; t.K = x.out[l];
      int rc = tc_gemm_launch(kTcDgrad, t, 1, st);
      if (rc != CATB200_OK) return rc;
    } else if (l > 0) {
      GemmNTArgs g = {};
      for (int z = 0; z < 2; ++z) {
        g.A[z] = reinterpret_cast<const bf16*>(ws + L.dZ[z][l]);
        g.B[z] = w16 + P.wt16[z][l];
        g.C[z] = reinterpret_cast<bf16*>(ws + L.dZ[z][l - 1]);
        g.H[z] = reinterpret_cast<const bf16*>(ws + L.H[z][l - 1]);
        g.dbias[z] = grads + P.b[z][l - 1];
      }
      g.lda = x.out[l]; g.ldb = x.out[l]; g.ldc = x.in[l];
      g.M = M; g.N = x.in[l]; g.K = x.out[l];
      dim3 grid((M + kBM - 1) / kBM, x.in[l] / kBN, 2);
      CATB200_CUDA_TRY(launch_pdl(gemm_nt_kernel<kEpiMulDelu>, grid, dim3(kGemmThreads), kSmemNT, st, g));
      CATB200_LAUNCH_CHECK();
    }
  }
  if (side) {  // join: the partial sums written on the side stream are reduced next
    CATB200_CUDA_TRY(cudaEventRecord(side->done, side->stream));
    CATB200_CUDA_TRY(cudaStreamWaitEvent(st, side->done, 0));
  }
  // 5. fold the split partial sums (hidden layers) and the head kernel's per-CTA rows into the flat gradient
  red.head_part = reinterpret_cast<const float*>(ws + L.head_part);
  red.head_rows = L.head_rows;
  red.gW4c = grads + P.w[0][3]; red.gb4c = grads + P.b[0][3];
  red.gW4a = grads + P.w[1][3]; red.gb4a = grads + P.b[1][3];
  red.glogstd = grads + P.logstd;
  for (int z = 0; z < 2; ++z) red.gb3[z] = grads + P.b[z][2];
  red.logstd = params + P.logstd; red.loss_acc = loss_acc;
  red.A = dims->act_dim; red.h3 = dims->h3; red.M = M;
  red.ent_coef = hp->ent_coef; red.vf_coef = hp->vf_coef;
  int max_elems = kHeadValues * 32;
  for (int sgm = 0; sgm < red.n_segments; ++sgm) max_elems = max(max_elems, red.N[sgm] * red.Kpad[sgm]);
  CATB200_CUDA_TRY(launch_pdl(reduce_partials_kernel, dim3((max_elems + 255) / 256, red.n_segments + 1), dim3(256), 0, st, red));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

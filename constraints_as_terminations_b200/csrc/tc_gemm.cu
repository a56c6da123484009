// 5th-generation tensor-core GEMMs for the actor-critic MLP (sm_100a: tcgen05.mma + TMEM + TMA).
//
// One warp-specialised kernel, three modes (both nets batched over blockIdx.z):
//   kFwd   : C[M,N]  = ELU(A[M,K] B[N,K]^T + bias)            A, B K-major            (hidden layers, forward)
//   kDgrad : C[M,N]  = (A[M,K] B[N,K]^T) * ELU'(H[M,N]), db += colsum      K-major    (dZ_{l-1} from dZ_l, W_l^T)
//   kWgrad : P[s,N,K] = sum_{m in split s} dZ[m,N]^T Hin[m,K]  A, B MN-major          (weight-gradient partials)
//
// CTA = one 128 x BN accumulator tile living in TMEM (128 lanes x BN fp32 columns).
//   warp 0      : TMA producer  - cp.async.bulk.tensor 2D boxes (64 elements = 128 B inner, SWIZZLE_128B)
//                                 into a 4-stage shared-memory ring, completion on mbarriers
//   warp 1      : TMEM allocator + MMA issuer - one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                                 (M = 128, N = BN, K = 16) from shared-memory matrix descriptors;
//                                 tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2..5  : epilogue      - tcgen05.ld (32 lanes x 32 columns per instruction) -> registers ->
//                                 bias/ELU or ELU'-scale (+ recursive-halving column sums) or fp32 partials
// Several CTAs are resident per SM (<= 96 KiB of shared memory, BN <= 128 TMEM columns each), so one CTA's
// epilogue overlaps another CTA's main loop without a persistent scheduler.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace catb200 {

constexpr int kTcThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 warps per TMEM lane quarter)
constexpr int kTcStages = 3;     // ring slots at most (fewer when the reduction is short)
constexpr int kTcBK = 64;  // reduction elements per stage (= 4 UMMA K-steps of 16)

template <int MODE, int BN>
struct TcSmem {
  static constexpr int kABytes = 128 * kTcBK * 2;  // 16 KiB: 128 (M) x 64 (K) bf16, or 2 boxes of 64 x 64 (MN-major)
  static constexpr int kBBytes = BN * kTcBK * 2;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kAux = 256 /*barriers*/ + 4 * BN * 4 /*bias tile (fwd) or 4 x BN column-sum scratch (dgrad)*/;
  static constexpr int total(int stages) { return stages * kStage + 1024 /*alignment slack*/ + kAux; }
};

template <int MODE, int BN>
__global__ void __launch_bounds__(kTcThreads, 2)
tc_gemm_kernel(const __grid_constant__ TcGemmArgs g) {
  using S = TcSmem<MODE, BN>;
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();  // let the next kernel of the chain get resident while this one runs
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const int stages = g.stages;
  const uint32_t bars = tiles + stages * S::kStage;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kTcStages, tmem_full_bar = bars + 16 * kTcStages;
  const uint32_t tmem_slot = bars + 16 * kTcStages + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_sm = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const CUtensorMap* mapA = &g.mapA[z];
  const CUtensorMap* mapB = &g.mapB[z];

  // tile coordinates and reduction range
  int row_base, col_base, k_begin, k_blocks;
  if (MODE == kTcWgrad) {
    const int k_tiles = g.N / BN;  // output columns (input features) per BN tile
    row_base = (blockIdx.x / k_tiles) * 128;  // dW rows (output features of the layer)
    col_base = (blockIdx.x % k_tiles) * BN;
    k_begin = blockIdx.y * g.m_range;
    const int k_end = min(g.K, k_begin + g.m_range);
    k_blocks = max(0, (k_end - k_begin + kTcBK - 1) / kTcBK);
  } else {
    row_base = blockIdx.x * 128;
    col_base = blockIdx.y * BN;
    k_begin = 0;
    k_blocks = g.K / kTcBK;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapA));
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(mapB));
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  // everything above (barriers, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on global data produced by it is read
  pdl_wait();
  if (MODE == kTcFwd && warp >= 2) {
    for (int c = threadIdx.x - 64; c < BN; c += kTcThreads - 64) bias_sm[c] = __ldg(g.bias[z] + col_base + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % stages;
        mbar_wait(empty_bar + 8 * s, ((kb / stages) & 1) ^ 1);
        const uint32_t sa = tiles + s * S::kStage, sb = sa + S::kABytes;
        mbar_expect_tx(full_bar + 8 * s, S::kStage);
        const int k0 = k_begin + kb * kTcBK;
        if (MODE == kTcWgrad) {
          // MN-major operands: boxes of 64 (contiguous features) x 64 (reduction rows), 8 KiB each
          for (int h = 0; h < 2; ++h) tma_load_2d(sa + h * 8192, mapA, full_bar + 8 * s, row_base + h * 64, k0);
          for (int h = 0; h < BN / 64; ++h) tma_load_2d(sb + h * 8192, mapB, full_bar + 8 * s, col_base + h * 64, k0);
        } else {
          tma_load_2d(sa, mapA, full_bar + 8 * s, k0, row_base);  // 64 (K) x 128 rows
          tma_load_2d(sb, mapB, full_bar + 8 * s, k0, col_base);  // 64 (K) x BN rows
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, BN, MODE == kTcWgrad, MODE == kTcWgrad);
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % stages;
      mbar_wait(full_bar + 8 * s, (kb / stages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = tiles + s * S::kStage, sb = sa + S::kABytes;
#pragma unroll
        for (int k = 0; k < kTcBK / 16; ++k) {
          uint64_t da, db;
          if (MODE == kTcWgrad) {
            // MN-major SW128: 64-feature chunks LBO = 8 KiB apart, 8-row reduction groups SBO = 1 KiB apart;
            // one UMMA K-step (16 reduction rows) = 2 KiB further
            da = make_smem_desc(sa + k * 2048, 8192, 1024);
            db = make_smem_desc(sb + k * 2048, 8192, 1024);
          } else {
            // K-major SW128: rows are 128 B, 8-row groups SBO = 1 KiB apart; one K-step = 32 B further
            da = make_smem_desc(sa + k * 32, 16, 1024);
            db = make_smem_desc(sb + k * 32, 16, 1024);
          }
          umma_bf16(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
      if (elect_one()) {
        umma_commit(empty_bar + 8 * s);                         // frees the ring slot once these MMAs retire
        if (kb == k_blocks - 1) umma_commit(tmem_full_bar);    // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // TMEM lanes 32*quarter .. +31 are the ones a warp may read (quarter = warp % 4); the two warps of a
    // quarter split the BN columns in halves.  Both 32-column chunks of a half are fetched before one wait.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = BN / 2;       // columns per warp
    constexpr int NCH = HC / 32;     // 32-column chunks per warp (1 or 2)
    const int row = row_base + quarter * 32 + lane;
    const int c_first = half * HC;
    // dgrad: the forward activations whose ELU' scales the result are fetched while the MMAs still run
    uint4 hv[NCH][4];
    if (MODE == kTcDgrad) {
      const bf16* __restrict__ hrow = g.H[z] + (size_t)row * g.ldc + col_base + c_first;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          hv[i][q] = row < g.M ? __ldg(reinterpret_cast<const uint4*>(hrow + i * 32 + q * 8)) : make_uint4(0, 0, 0, 0);
    }
    if (k_blocks > 0) {
      if (lane == 0) mbar_wait(tmem_full_bar, 0);  // one sleeping lane per warp instead of 256 pollers
      __syncwarp();
      mbar_wait(tmem_full_bar, 0);                 // already complete: a single acquire per thread
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + c_first;
    uint32_t v[NCH][32];
    if (k_blocks > 0) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) tmem_ld32(taddr + i * 32, v[i]);
#pragma unroll
      for (int i = 0; i < NCH; ++i) tmem_ld_wait(v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int e = 0; e < 32; ++e) v[i][e] = 0u;
    }
    // Output path (fwd / dgrad): the warp's 32 rows x 64 columns (128 B per row) are written into the ring's
    // first stage -- free by now, all MMAs have retired -- in the SWIZZLE_128B layout and leave with ONE TMA
    // tensor store per warp: full 128-byte lines instead of 32 scattered 16-byte stores per instruction.
    // box `half` = columns [half*64, half*64+64) of the tile: [128 rows][128 B], 16 KiB, rows of this warp at +quarter*4 KiB
    const int trow = quarter * 32 + lane;  // row inside the 128-row tile
    const uint32_t cbox = tiles + half * 16384;
    if (MODE == kTcFwd) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = q * 8 + e * 2;
            const float x0 = __uint_as_float(v[i][cc]) + bias_sm[c_first + i * 32 + cc];
            const float x1 = __uint_as_float(v[i][cc + 1]) + bias_sm[c_first + i * 32 + cc + 1];
            op[e] = pack_bf16x2(elu_fast(x0), elu_fast(x1));
          }
          st_shared_v4(cbox + trow * 128 + (((i * 4 + q) ^ (trow & 7)) << 4), o);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&g.mapC[z], cbox + quarter * 4096, col_base + c_first, row_base + quarter * 32);
        tma_store_commit_and_wait();
      }
      __syncwarp();
    } else if (MODE == kTcDgrad) {
      const bool live = row < g.M;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        float f[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&hv[i][q]);
          uint4 o;
          uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = q * 8 + e * 2;
            const float x0 = live ? __uint_as_float(v[i][cc]) * elu_grad_from_output(__low2float(hp[e])) : 0.0f;
            const float x1 = live ? __uint_as_float(v[i][cc + 1]) * elu_grad_from_output(__high2float(hp[e])) : 0.0f;
            f[cc] = x0;
            f[cc + 1] = x1;
            op[e] = pack_bf16x2(x0, x1);
          }
          st_shared_v4(cbox + trow * 128 + (((i * 4 + q) ^ (trow & 7)) << 4), o);
        }
        // column sums over this warp's 32 rows by recursive halving: after the 5 rounds lane l holds the
        // sum of column (i*32 + l)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const bool upper = (lane & o) != 0;
#pragma unroll
          for (int j = 0; j < o; ++j) {
            const float send = upper ? f[j] : f[j + o];
            const float keep = upper ? f[j + o] : f[j];
            f[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
          }
        }
        bias_sm[quarter * BN + c_first + i * 32 + lane] = f[0];  // per-quarter column sums -> shared scratch
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&g.mapC[z], cbox + quarter * 4096, col_base + c_first, row_base + quarter * 32);
        tma_store_commit_and_wait();
      }
      // the four row quarters cover the same columns: combine them in shared memory and issue ONE atomic per
      // column per CTA (same-address atomics from many CTAs serialise in L2, tens of ns each)
      asm volatile("bar.sync 1, 256;\n" ::: "memory");  // the 8 epilogue warps only
      const int et = threadIdx.x - 64;
      if (et < BN)
        atomicAdd(g.dbias[z] + col_base + et, (bias_sm[et] + bias_sm[BN + et]) + (bias_sm[2 * BN + et] + bias_sm[3 * BN + et]));
    } else {
      // weight-gradient partial: fp32 [128 rows (layer outputs) x BN (layer inputs)]
      float* __restrict__ prow = g.part[z] + ((size_t)blockIdx.y * g.M + row) * g.N + col_base + c_first;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(prow + i * 32 + q * 4) = make_uint4(v[i][q * 4], v[i][q * 4 + 1], v[i][q * 4 + 2], v[i][q * 4 + 3]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// ---- persistent variant (forward / dgrad) ----------------------------------------------------------------------
// ncu on the one-tile-per-CTA kernel above: a 128 x 128 x 256 tile needs 0.5 us of tensor-pipe time but its CTA lives
// ~15 us (prologue: barrier init, TMEM allocation, descriptor fetch; then load -> MMA -> TMEM read -> epilogue -> store
// strictly one after the other), so the tensor pipe is 5-10 % busy even with two CTAs per SM.  Here ONE CTA per SM
// walks the tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so CTAs that run side by side share the A
// tile in L2) with the three roles decoupled across tiles:
//   warp 0   TMA producer: keeps the 4-stage ring full across tile boundaries (running k-block counter)
//   warp 1   MMA issuer  : accumulates tile j into TMEM stage j & 1 (2 x BN columns allocated once)
//   warps 2-9 epilogue   : drain stage j & 1 (tcgen05.ld), hand it back (tmem_empty barrier, one arrival per warp)
//                          and do the bias/ELU or ELU'-scale math + TMA store while the MMAs of tile j + 1 run.
// The output staging area is separate from the ring (the ring is never idle any more).
constexpr int kPStages = 4;

template <int MODE, int BN>
struct PSmem {
  static constexpr int kABytes = 128 * kTcBK * 2, kBBytes = BN * kTcBK * 2, kStage = kABytes + kBBytes;
  static constexpr int kOut = 2 * 2 * 16384;                    // 2 tiles x two [128 rows][64 cols] bf16 boxes, SWIZZLE_128B
  static constexpr int kAux = 256 + 2 * 4 * BN * 4;             // barriers + double-buffered 4 x BN column-sum scratch
  static constexpr int kTotal = kPStages * kStage + kOut + 1024 /*alignment slack*/ + kAux;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

template <int MODE, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_persist_kernel(const __grid_constant__ TcGemmArgs g) {
  using S = PSmem<MODE, BN>;
  static_assert(MODE == kTcFwd || MODE == kTcDgrad, "persistent kernel: forward / dgrad only");
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t out_sm = tiles + kPStages * S::kStage;
  const uint32_t bars = out_sm + S::kOut;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kPStages;
  const uint32_t tfull_bar = bars + 16 * kPStages, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* colsum_sm = reinterpret_cast<float*>(smem_raw + (bars + 256 - raw));  // [2][4][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = g.N / BN, tiles_m = (g.M + 127) / 128;
  const int per_net = tiles_m * tiles_n, total = 2 * per_net;
  const int k_blocks = g.K / kTcBK;

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < 2; ++z) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapA[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapB[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapC[z]));
    }
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, kTcThreads / 32 - 2);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int z = t / per_net, r = t - z * per_net;
        const int row_base = (r / tiles_n) * 128, col_base = (r % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % kPStages;
          mbar_wait(empty_bar + 8 * s, ((it / kPStages) & 1) ^ 1);
          const uint32_t sa = tiles + s * S::kStage, sb = sa + S::kABytes;
          mbar_expect_tx(full_bar + 8 * s, S::kStage);
          tma_load_2d(sa, &g.mapA[z], full_bar + 8 * s, kb * kTcBK, row_base);  // 64 (K) x 128 rows
          tma_load_2d(sb, &g.mapB[z], full_bar + 8 * s, kb * kTcBK, col_base);  // 64 (K) x BN rows
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, BN, false, false);
    uint32_t it = 0, j = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      const uint32_t a = j & 1;
      mbar_wait(tempty_bar + 8 * a, ((j >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator stage
      tc_fence_after();
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const uint32_t s = it % kPStages;
        mbar_wait(full_bar + 8 * s, (it / kPStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = tiles + s * S::kStage, sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k) {
            // K-major SW128: rows are 128 B, 8-row groups SBO = 1 KiB apart; one K-step = 32 B further
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(tmem_base + a * BN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
        if (elect_one()) {
          umma_commit(empty_bar + 8 * s);                            // frees the ring slot once these MMAs retire
          if (kb == k_blocks - 1) umma_commit(tfull_bar + 8 * a);   // accumulator complete -> epilogue
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;      // TMEM lanes 32 * quarter .. + 31 are the ones this warp may read
    const int half = (warp - 2) >> 2;  // the two warps of a quarter split the BN columns in halves
    constexpr int HC = BN / 2, NCH = HC / 32;
    const int c_first = half * HC;
    const int trow = quarter * 32 + lane;  // row inside the 128-row tile
    uint32_t j = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      const int z = t / per_net, r = t - z * per_net;
      const int row_base = (r / tiles_n) * 128, col_base = (r % tiles_n) * BN;
      const int row = row_base + trow;
      const uint32_t a = j & 1;
      const uint32_t cbox = out_sm + a * 32768 + half * 16384;  // output staging alternates with the tile parity
      // dgrad: the forward activations whose ELU' scales the result are fetched while the MMAs still run
      uint4 hv[NCH][4];
      if (MODE == kTcDgrad) {
        const bf16* __restrict__ hrow = g.H[z] + (size_t)row * g.ldc + col_base + c_first;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            hv[i][q] = row < g.M ? __ldg(reinterpret_cast<const uint4*>(hrow + i * 32 + q * 8)) : make_uint4(0, 0, 0, 0);
      }
      // forward: this warp's 64 bias values, fetched coalesced (2 loads per lane) into the warp's own scratch row while
      // the MMAs still run; the epilogue then reads them back as 16-byte broadcasts.  (Per-element warp-uniform global
      // loads cost a descriptor setup each: ncu counted 17 thread instructions per output element with them.)
      float* bsm = colsum_sm + (((j & 1) * 8 + (warp - 2)) * 64);
      if (MODE == kTcFwd) {
        const float* __restrict__ bp = g.bias[z] + col_base + c_first;
        bsm[lane] = __ldg(bp + lane);
        bsm[32 + lane] = __ldg(bp + 32 + lane);
      }
      if (lane == 0) mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);  // one sleeping lane per warp
      __syncwarp();
      mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);                 // already complete: a single acquire per thread
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + c_first;
      uint32_t v[NCH][32];
#pragma unroll
      for (int i = 0; i < NCH; ++i) tmem_ld32(taddr + i * 32, v[i]);
#pragma unroll
      for (int i = 0; i < NCH; ++i) tmem_ld_wait(v[i]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * a);  // the MMA warp may overwrite this stage (tile j + 2)

      if (MODE == kTcFwd) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b0 = *reinterpret_cast<const float4*>(bsm + i * 32 + q * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(bsm + i * 32 + q * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = q * 8 + e * 2;
              const float x0 = __uint_as_float(v[i][cc]) + bb[e * 2];
              const float x1 = __uint_as_float(v[i][cc + 1]) + bb[e * 2 + 1];
              op[e] = pack_bf16x2(elu_fast(x0), elu_fast(x1));
            }
            st_shared_v4(cbox + trow * 128 + (((i * 4 + q) ^ (trow & 7)) << 4), o);
          }
        }
      } else {
        // rows beyond M need no masking: TMA zero-fills them in the A tile, so their accumulators are exactly 0 (and
        // their ELU' factor is 1: hv was set to 0), i.e. they add nothing to the column sums; the TMA store clips them
        float* colsum = colsum_sm + (j & 1) * 4 * BN;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          float f[32];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&hv[i][q]);
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = q * 8 + e * 2;
              const float x0 = __uint_as_float(v[i][cc]) * elu_grad_from_output(__low2float(hp[e]));
              const float x1 = __uint_as_float(v[i][cc + 1]) * elu_grad_from_output(__high2float(hp[e]));
              f[cc] = x0;
              f[cc + 1] = x1;
              op[e] = pack_bf16x2(x0, x1);
            }
            st_shared_v4(cbox + trow * 128 + (((i * 4 + q) ^ (trow & 7)) << 4), o);
          }
          // column sums over this warp's 32 rows by recursive halving: lane l ends with the sum of column i*32 + l
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const bool upper = (lane & o) != 0;
#pragma unroll
            for (int jj = 0; jj < o; ++jj) {
              const float send = upper ? f[jj] : f[jj + o];
              const float keep = upper ? f[jj + o] : f[jj];
              f[jj] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          colsum[quarter * BN + c_first + i * 32 + lane] = f[0];
        }
      }
      // this warp's 32 rows x 64 columns leave with ONE TMA tensor store.  The slab is this warp's own and there are
      // two of them: only the store issued one tile ago (which read the slab the next tile will overwrite) has to be
      // done before going on, the one just issued drains behind the next tile's math; no CTA-wide barrier involved
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&g.mapC[z], cbox + quarter * 4096, col_base + c_first, row_base + quarter * 32);
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      }
      __syncwarp();
      if (MODE == kTcDgrad) {
        // the four row quarters cover the same columns: combine them in shared memory and issue ONE atomic per
        // column per tile.  The scratch alternates between two buffers: a warp that races ahead writes tile j + 1's
        // sums into the other one, and reaches tile j + 2 only through tile j + 1's barrier.
        asm volatile("bar.sync 1, 256;\n" ::: "memory");  // the 8 epilogue warps only
        const float* colsum = colsum_sm + (j & 1) * 4 * BN;
        const int et = threadIdx.x - 64;
        if (et < BN)
          atomicAdd(g.dbias[z] + col_base + et, (colsum[et] + colsum[BN + et]) + (colsum[2 * BN + et] + colsum[3 * BN + et]));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // staging read out before exit
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                     uint32_t box_outer);

static EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t inner, outer, ld;
  uint32_t bi, bo;
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
};
static TmapEntry g_tmap_cache[128];
static int g_tmap_count = 0;

// bf16 row-major [outer, inner] with leading dimension ld (elements); box = [box_outer, box_inner].
// Encodings are memoised: the trainer reuses a handful of (pointer, shape) combinations every step.
int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer) {
  for (int i = 0; i < g_tmap_count; ++i) {
    const TmapKey& k = g_tmap_cache[i].key;
    if (k.ptr == ptr && k.inner == inner && k.outer == outer && k.ld == ld && k.bi == box_inner && k.bo == box_outer) {
      *map = g_tmap_cache[i].map;
      return CATB200_OK;
    }
  }
  int rc = encode_tmap_bf16(map, ptr, inner, outer, ld, box_inner, box_outer);
  if (rc == CATB200_OK) {
    const int slot = g_tmap_count < 128 ? g_tmap_count++ : 127;
    g_tmap_cache[slot].key = TmapKey{ptr, inner, outer, ld, box_inner, box_outer};
    g_tmap_cache[slot].map = *map;
  }
  return rc;
}

int encode_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                     uint32_t box_outer) {
  EncodeTiledFn fn = encoder();
  if (!fn) return CATB200_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CATB200_OK : CATB200_ERR_CUDA;
}

template <int MODE, int BN>
static int launch_one(TcGemmArgs g, dim3 grid, cudaStream_t st) {
  using S = TcSmem<MODE, BN>;
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total(kTcStages)));
    attr = true;
  }
  // ring depth: no more slots than reduction blocks (a K = 64 layer needs one), which keeps several CTAs per SM
  const int red = MODE == kTcWgrad ? g.m_range : g.K;
  g.stages = max(1, min(kTcStages, (red + kTcBK - 1) / kTcBK));
  CATB200_CUDA_TRY(launch_pdl(tc_gemm_kernel<MODE, BN>, grid, dim3(kTcThreads), (size_t)S::total(g.stages), st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// persistent kernel for launches whose tiles do not all fit on the machine at once (two one-tile CTAs per SM);
// CATB200_TC_PERSIST=0: always one tile per CTA
static bool use_persistent(int total_tiles) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_TC_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && total_tiles > 2 * kNumSMs;
}

template <int MODE, int BN>
static int launch_persist(TcGemmArgs g, int total_tiles, cudaStream_t st) {
  using S = PSmem<MODE, BN>;
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_persist_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr = true;
  }
  g.stages = kPStages;
  CATB200_CUDA_TRY(launch_pdl(tc_gemm_persist_kernel<MODE, BN>, dim3(min(total_tiles, kNumSMs)), dim3(kTcThreads), (size_t)S::kTotal, st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int tc_gemm_launch(int mode, const TcGemmArgs& g, int splits, cudaStream_t st) {
  if (mode == kTcFwd) {
    if (g.N % 128) return CATB200_ERR_UNSUPPORTED;
    const int total = 2 * ((g.M + 127) / 128) * (g.N / 128);
    if (use_persistent(total)) return launch_persist<kTcFwd, 128>(g, total, st);
    return launch_one<kTcFwd, 128>(g, dim3((g.M + 127) / 128, g.N / 128, 2), st);
  }
  if (mode == kTcDgrad) {
    if (g.N % 128) return CATB200_ERR_UNSUPPORTED;
    const int total = 2 * ((g.M + 127) / 128) * (g.N / 128);
    if (use_persistent(total)) return launch_persist<kTcDgrad, 128>(g, total, st);
    return launch_one<kTcDgrad, 128>(g, dim3((g.M + 127) / 128, g.N / 128, 2), st);
  }
  // wgrad: g.M = layer outputs (dW rows), g.N = padded layer inputs (dW cols), g.K = minibatch rows
  if (g.M % 128) return CATB200_ERR_UNSUPPORTED;
  if (g.N % 128 == 0) return launch_one<kTcWgrad, 128>(g, dim3((g.M / 128) * (g.N / 128), splits, 2), st);
  if (g.N % 64 == 0) return launch_one<kTcWgrad, 64>(g, dim3((g.M / 128) * (g.N / 64), splits, 2), st);
  return CATB200_ERR_UNSUPPORTED;
}

}  // namespace catb200

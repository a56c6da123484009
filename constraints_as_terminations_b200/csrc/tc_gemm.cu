// 5th-generation tensor-core GEMMs for the actor-critic MLP (sm_100a: tcgen05.mma + TMEM + TMA), templated on the
// operand precision (PrecT in tc_ptx.cuh): tf32 (fp32 storage, kind::tf32 -- the reference's GPU numerics,
// scripts/clean_rl/train.py:86-87) or bf16 (kind::f16).  Both nets (0 critic, 1 actor) are batched into every launch.
//
//   mlp_gemm_kernel<kTcFwd>   : C[M,N] = ELU(A[M,K] B[N,K]^T + bias)           A, B K-major   (hidden layers, forward)
//   mlp_gemm_kernel<kTcDgrad> : C[M,N] = (A[M,K] B[N,K]^T) * ELU'(H[M,N])      A, B K-major   (dZ_{l-1} from dZ_l, W_l^T)
//   mlp_wgrad_kernel          : dW[outs,ins] += dZ^T Hin, db[outs] += dZ^T 1   A, B MN-major  (no transposed copies)
//
// mlp_gemm_kernel is persistent: one CTA per SM walks the 128 x 128 output tiles with three decoupled roles
//   warp 0     TMA producer : cp.async.bulk.tensor boxes (128-byte rows, SWIZZLE_128B) into a 4-stage ring that runs
//                             across tile boundaries
//   warp 1     MMA issuer   : one elected lane issues tcgen05.mma (M = 128, N = 128, 32 bytes of K per instruction)
//                             into one of two TMEM accumulator stages
//   warps 2-9  epilogue     : tcgen05.ld -> registers -> bias/ELU or ELU'-scale -> per-warp swizzled slab in shared
//                             memory -> ONE TMA tensor store per 32 x 128-byte slab.  In dgrad the slab first receives
//                             the H values by TMA, so ELU' needs no strided global loads.
// mlp_wgrad_kernel is one (tile, row range) per CTA: split over the minibatch rows so the launch fills the machine, the
// fp32 tile leaves TMEM straight into the gradient accumulators with red.global.add.v4.f32 (no partial-sum buffers,
// no reduction kernel), and the bias gradient is the same A operand multiplied by a tile of ones (N = 16), so the
// dgrad epilogue carries no column sums.
//
// Row order.  With fp32 activations a 16384-row minibatch moves ~720 MB through HBM (the activations of both nets are
// 234 MB: twice the L2) and warm-cache ncu counters showed EVERY consumer re-reading from DRAM what the previous launch
// had just written: producer and consumer both walked the rows upwards, so the rows a consumer wanted first were the ones
// the L2 had dropped first.  All kernels therefore process rows in one global time order (tile index = row tile first,
// then net, then column tile; the weight-gradient CTAs sweep interleaved 64-row blocks instead of owning contiguous row
// ranges) and the launches of a minibatch alternate its direction (`reverse`), so that each one starts on the rows the
// previous one touched last.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace catb200 {

constexpr int kTcThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (2 warps per TMEM lane quarter)
constexpr int kEpiWarps = kTcThreads / 32 - 2;
constexpr int kRowBytes = 128;                 // one SWIZZLE_128B row
constexpr int kATileBytes = 128 * kRowBytes;   // 16 KiB: 128 rows (M or N) x 128 bytes of K
constexpr int kSlabBytes = 32 * kRowBytes;     // 4 KiB: one epilogue warp's 32 rows x 128 bytes of output

// Epilogue geometry.  Measured on the B200 (profiles/README.md, round 2), tf32: 16 epilogue warps (4 per TMEM lane quarter,
// 32 output columns = one 128-byte slab row each, one slab buffer per warp) make the forward launches 3 % faster than 8
// -- the 8-warp fp32 epilogue issues ~12 instructions per output element from 2 warps per scheduler at IPC 0.35 -- but the
// dgrad launches 10 % slower (one buffer per warp: the H slab of the next tile cannot be requested before the store of
// this one has been read).  So: forward tf32 16 warps, everything else 8 warps with two slab buffers (tf32: the two
// 32-column halves of a warp's 64 columns; bf16: one 64-column slab, alternating per tile).
template <int PREC, int MODE, int STAGES>
struct PCfg {
  static constexpr int kPStages = STAGES;
  static constexpr int kEpi = (PREC == kPrecTf32 && MODE == kTcFwd) ? 16 : 8;
  static constexpr int kThreads = 64 + 32 * kEpi;
  static constexpr int kWCols = 128 / (kEpi / 4);                  // output columns per epilogue warp: 32 / 64
  static constexpr int kBufs = kEpi == 16 ? 1 : 2;                 // slab buffers per warp
  static constexpr int kStage = 2 * kATileBytes;                   // A tile + B tile (BN = 128)
  static constexpr int kSlabs = kEpi * kBufs * kSlabBytes;
  static constexpr int kBars = 8 * (2 * kPStages + 4 + 2 * kEpi);
  static constexpr int kBias = kEpi * kWCols * 4;                  // one private row of bias values per warp
  // no alignment slack: the kernel has no static shared memory, so the dynamic window starts 1024-byte aligned (checked)
  static constexpr int kTotal = kPStages * kStage + kSlabs + ((kBars + 15) & ~15) + 16 + kBias;
  static_assert(kTotal <= 232448, "shared memory budget");
};

// cp.async.bulk.wait_group.read with a pending count known after unrolling
__device__ __forceinline__ void bulk_wait_read(int pending) {
  if (pending <= 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
}

template <int MODE, int PREC, int STAGES>
__global__ void __launch_bounds__(PCfg<PREC, MODE, STAGES>::kThreads, 1)
mlp_gemm_kernel(const __grid_constant__ TcGemmArgs g) {
  using P = PrecT<PREC>;
  using S = PCfg<PREC, MODE, STAGES>;
  constexpr int BN = 128;
  constexpr int kPStages = STAGES;
  constexpr int kEpi = S::kEpi, kWCols = S::kWCols, kBufs = S::kBufs;
  constexpr int CH = kRowBytes / (int)sizeof(typename P::T);  // output columns per slab row: 64 bf16 / 32 fp32
  constexpr int NCHUNK = kWCols / CH;                         // slabs per warp and tile: 1, or 2 (tf32 with 8 warps)
  constexpr int LD = kWCols / 32;                             // 32-column TMEM loads per warp and tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  pdl_launch_dependents();
  const unsigned long long hA = g.hintA ? g.hintA : kL2EvictNormal, hB = g.hintB ? g.hintB : kL2EvictNormal;
  const unsigned long long hC = g.hintC ? g.hintC : kL2EvictNormal, hH = g.hintH ? g.hintH : kL2EvictNormal;
  (void)hH;
  const uint32_t raw = smem_u32(smem_raw);
  if (raw & 1023u) __trap();  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t tiles = raw;
  const uint32_t slabs = tiles + kPStages * S::kStage;
  const uint32_t bars = slabs + S::kSlabs;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kPStages;
  const uint32_t tfull_bar = bars + 16 * kPStages, tempty_bar = tfull_bar + 16;
  const uint32_t h_bar = tempty_bar + 16;  // [kEpi][2]
  const uint32_t tmem_slot = bars + ((S::kBars + 15) & ~15);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_sm = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - raw));  // [kEpi][kWCols]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = g.N / BN, tiles_m = (g.M + 127) / 128;
  const int total = 2 * tiles_m * tiles_n;
  const int k_blocks = g.K / P::kBK;
  // tile t of the launch: row tile first, then net, then column tile -- in time order, or backwards
  auto tile_coords = [&](int t, int& z, int& row_base, int& col_base) {
    const int tt = g.reverse ? total - 1 - t : t;
    const int m = tt / (2 * tiles_n), r = tt - m * 2 * tiles_n;
    z = r / tiles_n;
    row_base = m * 128;
    col_base = (r - z * tiles_n) * BN;
  };

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < 2; ++z) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapA[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapB[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapC[z]));
      if (MODE == kTcDgrad) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapH[z]));
    }
    for (int s = 0; s < kPStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, kEpi);  // one arrival per epilogue warp
    }
    for (int i = 0; i < 2 * kEpi; ++i) mbar_init(h_bar + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  // everything above (barriers, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail;
  // from here on global data produced by it is read
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
        int z, row_base, col_base;
        tile_coords(t, z, row_base, col_base);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % kPStages;
          mbar_wait(empty_bar + 8 * s, ((it / kPStages) & 1) ^ 1);
          const uint32_t sa = tiles + s * S::kStage, sb = sa + kATileBytes;
          mbar_expect_tx(full_bar + 8 * s, S::kStage);
          tma_load_2d(sa, &g.mapA[z], full_bar + 8 * s, kb * P::kBK, row_base, hA);  // 128 bytes of K x 128 rows
          tma_load_2d(sb, &g.mapB[z], full_bar + 8 * s, kb * P::kBK, col_base, hB);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(P::kFmt, 128, BN, false, false);
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
          const uint32_t sa = tiles + s * S::kStage, sb = sa + kATileBytes;
#pragma unroll
          for (int k = 0; k < P::kBK / P::kUmmaK; ++k) {
            // K-major SW128: rows are 128 B, 8-row groups SBO = 1 KiB apart; one K-step = 32 B further
            const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
            umma<PREC>(tmem_base + a * BN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
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
    // ===================== epilogue (warps 2 .. kEpi + 1) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lanes 32 * quarter .. + 31 are the ones this warp may read
    const int c_first = (ew >> 2) * kWCols;  // the warps of a quarter split the BN columns
    const uint32_t my_slabs = slabs + ew * kBufs * kSlabBytes;
    const uint32_t my_hbar = h_bar + ew * 16;
    const uint32_t lane_row = lane * kRowBytes, lane_x = lane & 7;
    uint32_t j = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      int z, row_base, col_base;
      tile_coords(t, z, row_base, col_base);
      const uint32_t a = j & 1;
      // Slab n = j * NCHUNK + c of this warp lives in buffer n % kBufs.  Before a buffer is refilled, the TMA store that
      // last read it (slab n - kBufs) must have finished reading; bulk groups complete in order, so it is enough to bound
      // the number of stores still pending.
      float* bsm = bias_sm + ew * kWCols;  // private to this warp: rewritten only after the previous tile's math
      if (MODE == kTcDgrad) {
        // the H values of this warp's slabs arrive by TMA while the MMAs still run
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) {
            const uint32_t n = j * NCHUNK + c, buf = n % kBufs;
            bulk_wait_read(kBufs - 1 - c);  // stores issued so far: up to slab j * NCHUNK - 1
            mbar_expect_tx(my_hbar + 8 * buf, kSlabBytes);
            tma_load_2d(my_slabs + buf * kSlabBytes, &g.mapH[z], my_hbar + 8 * buf, col_base + c_first + c * CH, row_base + quarter * 32, hH);
          }
        }
      } else {
        // this warp's bias values, fetched coalesced into its own scratch row; the math reads them as broadcasts
        const float* __restrict__ bp = g.bias[z] + col_base + c_first;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < kWCols; c += 32) bsm[c + lane] = __ldg(bp + c + lane);
      }
      if (lane == 0) mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);  // one sleeping lane per warp
      __syncwarp();
      mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);                 // already complete: a single acquire per thread
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + c_first;
      uint32_t v[LD][32];
#pragma unroll
      for (int l = 0; l < LD; ++l) tmem_ld32(taddr + l * 32, v[l]);
#pragma unroll
      for (int l = 0; l < LD; ++l) tmem_ld_wait(v[l]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * a);  // the MMA warp may overwrite this stage (tile j + 2)

#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const uint32_t n = j * NCHUNK + c, buf = n % kBufs;
        const uint32_t slab = my_slabs + buf * kSlabBytes;
        if (MODE == kTcDgrad) {
          mbar_wait(my_hbar + 8 * buf, (n / kBufs) & 1);
        } else {
          if (lane == 0) bulk_wait_read(kBufs - 1);  // stores issued so far: up to slab n - 1
          __syncwarp();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {  // 16-byte chunk q of this thread's 128-byte row
          const uint32_t addr = slab + lane_row + ((q ^ lane_x) << 4);
          uint4 o;
          if (PREC == kPrecTf32) {
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __uint_as_float(v[c][q * 4 + e]);
            if (MODE == kTcFwd) {
              const float4 b = *reinterpret_cast<const float4*>(bsm + c * 32 + q * 4);
              x[0] = elu_fast(x[0] + b.x); x[1] = elu_fast(x[1] + b.y); x[2] = elu_fast(x[2] + b.z); x[3] = elu_fast(x[3] + b.w);
            } else {
              const uint4 h = ld_shared_v4(addr);
              x[0] *= elu_grad_from_output(__uint_as_float(h.x)); x[1] *= elu_grad_from_output(__uint_as_float(h.y));
              x[2] *= elu_grad_from_output(__uint_as_float(h.z)); x[3] *= elu_grad_from_output(__uint_as_float(h.w));
            }
            o = make_uint4(__float_as_uint(round_tf32(x[0])), __float_as_uint(round_tf32(x[1])),
                           __float_as_uint(round_tf32(x[2])), __float_as_uint(round_tf32(x[3])));
          } else {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[(q >> 2) % LD][(q & 3) * 8 + e]);
            if (MODE == kTcFwd) {
              const float4 b0 = *reinterpret_cast<const float4*>(bsm + q * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bsm + q * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = elu_fast(x[e] + bb[e]);
            } else {
              const uint4 h = ld_shared_v4(addr);
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                x[2 * e] *= elu_grad_from_output(__low2float(hp[e]));
                x[2 * e + 1] *= elu_grad_from_output(__high2float(hp[e]));
              }
            }
            o = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
          }
          st_shared_v4(addr, o);
        }
        // rows beyond M need no masking: TMA zero-fills them in the A tile (accumulators exactly 0) and clips the store
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&g.mapC[z], slab, col_base + c_first + c * CH, row_base + quarter * 32, hC);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // stores complete before exit
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---- 256-row tiles (minibatch-sized launches) -----------------------------------------------------------------------
// With fp32 operands a 128 x 128 tile moves 32 KiB of operands per 32 reduction elements and the launch is bound by
// L2 -> SM operand traffic (~10 TB/s chip-wide), not by the tensor pipe: ncu shows 36 % tensor activity on the
// 512 -> 256 layer while 300 MB cross the L2 for 8.6 GFLOP.  This variant gives one CTA a 256 x BN tile as two 128-row
// accumulators that share every B stage (BN = 256: 64 KiB per k-block for four times the MACs of the 128 x 128 tile,
// i.e. half the operand bytes per flop; BN = 128 for the 128-wide layer: 48 KiB for twice the MACs).  BN = 256 fills all
// 512 TMEM columns with one accumulator stage (the epilogue of a tile does not overlap the next tile's MMAs -- launches
// of this size have 1-2 tiles per CTA); BN = 128 keeps two stages.  Roles, slabs and the dgrad H path are those of
// mlp_gemm_kernel; each epilogue warp walks its 32 rows x BN/2 columns of both accumulators slab by slab.
template <int BN>
struct P2Smem {
  static constexpr int kStages = BN == 256 ? 2 : 3;
  static constexpr int kStage = 2 * kATileBytes + BN * kRowBytes;   // two A tiles + one B tile
  static constexpr int kSlabs = kEpiWarps * 2 * kSlabBytes;
  static constexpr int kBars = 8 * (2 * kStages + 4 + 2 * kEpiWarps);
  static constexpr int kBias = 2 * kEpiWarps * (BN / 2) * 4;        // double-buffered BN/2 bias values per warp
  static constexpr int kTotal = kStages * kStage + kSlabs + ((kBars + 15) & ~15) + 16 + kBias + 1024 /*alignment slack*/;
};

template <int MODE, int PREC, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_gemm256_kernel(const __grid_constant__ TcGemmArgs g) {
  using P = PrecT<PREC>;
  using S = P2Smem<BN>;
  constexpr int kStages = S::kStages;
  constexpr int kAccStages = 512 / (2 * BN);                  // TMEM accumulator stages: 1 (BN = 256) or 2
  constexpr int CH = kRowBytes / (int)sizeof(typename P::T);  // output columns per slab row: 64 bf16 / 32 fp32
  constexpr int LD = CH / 32;                                 // 32-column TMEM loads per slab
  constexpr int NSLAB = 2 * (BN / 2) / CH;                    // slabs per warp and tile (both accumulators)
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const unsigned long long hA = g.hintA ? g.hintA : kL2EvictNormal, hB = g.hintB ? g.hintB : kL2EvictNormal;
  const unsigned long long hC = g.hintC ? g.hintC : kL2EvictNormal, hH = g.hintH ? g.hintH : kL2EvictNormal;
  (void)hH;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t slabs = tiles + kStages * S::kStage;
  const uint32_t bars = slabs + S::kSlabs;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kStages;
  const uint32_t tfull_bar = bars + 16 * kStages, tempty_bar = tfull_bar + 16;
  const uint32_t h_bar = tempty_bar + 16;  // [kEpiWarps][2]
  const uint32_t tmem_slot = bars + ((S::kBars + 15) & ~15);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_sm = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - raw));  // [2][kEpiWarps][BN / 2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = g.N / BN, tiles_m = (g.M + 255) / 256;
  const int per_net = tiles_m * tiles_n, total = 2 * per_net;
  const int k_blocks = g.K / P::kBK;

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < 2; ++z) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapA[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapB[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapC[z]));
      if (MODE == kTcDgrad) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapH[z]));
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, kEpiWarps);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(h_bar + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
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
        const int row_base = (r / tiles_n) * 256, col_base = (r % tiles_n) * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % kStages;
          mbar_wait(empty_bar + 8 * s, ((it / kStages) & 1) ^ 1);
          const uint32_t sa = tiles + s * S::kStage, sb = sa + 2 * kATileBytes;
          mbar_expect_tx(full_bar + 8 * s, S::kStage);
          tma_load_2d(sa, &g.mapA[z], full_bar + 8 * s, kb * P::kBK, row_base, hA);                     // rows   0..127 of the tile
          tma_load_2d(sa + kATileBytes, &g.mapA[z], full_bar + 8 * s, kb * P::kBK, row_base + 128, hA);  // rows 128..255 (zero-filled beyond M)
#pragma unroll
          for (int h = 0; h < BN / 128; ++h) tma_load_2d(sb + h * kATileBytes, &g.mapB[z], full_bar + 8 * s, kb * P::kBK, col_base + h * 128, hB);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(P::kFmt, 128, BN, false, false);
    uint32_t it = 0, j = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      const uint32_t a = j % kAccStages;
      mbar_wait(tempty_bar + 8 * a, ((j / kAccStages) & 1) ^ 1);  // the epilogue has drained this accumulator stage
      tc_fence_after();
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const uint32_t s = it % kStages;
        mbar_wait(full_bar + 8 * s, (it / kStages) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = tiles + s * S::kStage, sb = sa + 2 * kATileBytes;
#pragma unroll
          for (int k = 0; k < P::kBK / P::kUmmaK; ++k) {
            const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);  // BN rows of 128 B, 8-row groups 1 KiB apart
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint64_t da = make_smem_desc(sa + sub * kATileBytes + k * 32, 16, 1024);
              umma<PREC>(tmem_base + a * 2 * BN + sub * BN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
        }
        __syncwarp();
        if (elect_one()) {
          umma_commit(empty_bar + 8 * s);
          if (kb == k_blocks - 1) umma_commit(tfull_bar + 8 * a);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int c_first = half * (BN / 2);
    const uint32_t my_slabs = slabs + ew * 2 * kSlabBytes;
    const uint32_t my_hbar = h_bar + ew * 16;
    const uint32_t lane_row = lane * kRowBytes, lane_x = lane & 7;
    uint32_t j = 0, n_slab = 0;  // n_slab: running slab counter of this warp (buffer = n_slab & 1, H-barrier phase = n_slab >> 1)
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      const int z = t / per_net, r = t - z * per_net;
      const int row_base = (r / tiles_n) * 256, col_base = (r % tiles_n) * BN;
      const uint32_t a = j % kAccStages;
      float* bsm = bias_sm + ((j & 1) * kEpiWarps + ew) * (BN / 2);
      // slab i of the tile: accumulator sub = i / (NSLAB / 2), columns c_first + (i % (NSLAB / 2)) * CH
      auto slab_col = [&](int i) { return col_base + c_first + (i % (NSLAB / 2)) * CH; };
      auto slab_row = [&](int i) { return row_base + (i / (NSLAB / 2)) * 128 + quarter * 32; };
      if (MODE == kTcDgrad) {
        if (lane == 0) {  // H of the first slab, in flight while the MMAs still run
          asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
          const uint32_t buf = n_slab & 1;
          mbar_expect_tx(my_hbar + 8 * buf, kSlabBytes);
          tma_load_2d(my_slabs + buf * kSlabBytes, &g.mapH[z], my_hbar + 8 * buf, slab_col(0), slab_row(0), hH);
        }
      } else {
        const float* __restrict__ bp = g.bias[z] + col_base + c_first;
#pragma unroll
        for (int c = lane; c < BN / 2; c += 32) bsm[c] = __ldg(bp + c);
      }
      if (lane == 0) mbar_wait(tfull_bar + 8 * a, (j / kAccStages) & 1);
      __syncwarp();
      mbar_wait(tfull_bar + 8 * a, (j / kAccStages) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * 2 * BN + c_first;
#pragma unroll 1
      for (int i = 0; i < NSLAB; ++i, ++n_slab) {
        const uint32_t buf = n_slab & 1;
        const uint32_t slab = my_slabs + buf * kSlabBytes;
        const int sub = i / (NSLAB / 2), ci = i % (NSLAB / 2);
        uint32_t v[LD][32];
#pragma unroll
        for (int l = 0; l < LD; ++l) tmem_ld32(tbase + sub * BN + ci * CH + l * 32, v[l]);
        if (MODE == kTcDgrad) {
          if (lane == 0 && i + 1 < NSLAB) {  // H of the next slab: its buffer was last read by the store of slab i - 1
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
            const uint32_t nb = buf ^ 1;
            mbar_expect_tx(my_hbar + 8 * nb, kSlabBytes);
            tma_load_2d(my_slabs + nb * kSlabBytes, &g.mapH[z], my_hbar + 8 * nb, slab_col(i + 1), slab_row(i + 1), hH);
          }
          mbar_wait(my_hbar + 8 * buf, (n_slab >> 1) & 1);
        } else {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");  // store of slab i - 2 (same buffer)
          __syncwarp();
        }
#pragma unroll
        for (int l = 0; l < LD; ++l) tmem_ld_wait(v[l]);
        if (i == NSLAB - 1) {  // last TMEM read of this tile: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + 8 * a);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t addr = slab + lane_row + ((q ^ lane_x) << 4);
          uint4 o;
          if (PREC == kPrecTf32) {
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __uint_as_float(v[0][q * 4 + e]);
            if (MODE == kTcFwd) {
              const float4 b = *reinterpret_cast<const float4*>(bsm + ci * CH + q * 4);
              x[0] = elu_fast(x[0] + b.x); x[1] = elu_fast(x[1] + b.y); x[2] = elu_fast(x[2] + b.z); x[3] = elu_fast(x[3] + b.w);
            } else {
              const uint4 h = ld_shared_v4(addr);
              x[0] *= elu_grad_from_output(__uint_as_float(h.x)); x[1] *= elu_grad_from_output(__uint_as_float(h.y));
              x[2] *= elu_grad_from_output(__uint_as_float(h.z)); x[3] *= elu_grad_from_output(__uint_as_float(h.w));
            }
            o = make_uint4(__float_as_uint(round_tf32(x[0])), __float_as_uint(round_tf32(x[1])),
                           __float_as_uint(round_tf32(x[2])), __float_as_uint(round_tf32(x[3])));
          } else {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[(q >> 2) % LD][(q & 3) * 8 + e]);
            if (MODE == kTcFwd) {
              const float4 b0 = *reinterpret_cast<const float4*>(bsm + ci * CH + q * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bsm + ci * CH + q * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = elu_fast(x[e] + bb[e]);
            } else {
              const uint4 h = ld_shared_v4(addr);
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                x[2 * e] *= elu_grad_from_output(__low2float(hp[e]));
                x[2 * e + 1] *= elu_grad_from_output(__high2float(hp[e]));
              }
            }
            o = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
          }
          st_shared_v4(addr, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&g.mapC[z], slab, slab_col(i), slab_row(i), hC);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- CTA pairs: 256 x 256 tiles on two SMs (cta_group::2) -----------------------------------------------------------
// The 128 x 128 kernel is bound by L2 -> SM operand traffic (fp32 operands: 32 KiB per 128 x 128 x 32 MACs, ~13 TB/s
// chip-wide at 16384 rows).  A CTA pair (a 2-CTA cluster = the two SMs of a TPC) computes one 256 x 256 tile with
// tcgen05.mma.cta_group::2: each CTA stages ITS 128 rows of A and ITS 128 of the 256 B rows -- the same 32 KiB per
// k-block and CTA as before, for twice the MACs: half the operand bytes per flop, and unlike the single-CTA 256-row
// variant above each CTA's accumulator is 256 columns, so two TMEM stages fit and the epilogue keeps overlapping the
// next tile's MMAs.
//   both CTAs : warp 0 = TMA producer for its own A rows / B half (the loads complete on the LEADER's full barrier,
//               which expects both CTAs' bytes); warps 2-9 = epilogue of its own 128 rows (slabs / dgrad H path as in
//               mlp_gemm_kernel); arrivals on the leader's tmem_empty barrier come from both CTAs' epilogue warps
//   leader    : warp 1 issues the pair's MMAs; tcgen05.commit multicasts "ring slot free" and "accumulator complete" to
//               the barriers of both CTAs
// Epilogue: ncu / timing of the first version (8 epilogue warps) showed these kernels bound by the EPILOGUE, not by
// operands -- a dgrad tile cost ~9 us per CTA however few bytes the main loop moved, because every 32 x 32 slab waited a
// full L2 round trip for its H values and only 8 warps were there to overlap anything.  tf32 therefore runs 16 epilogue
// warps (4 per TMEM lane quarter, 64 columns = two slabs each) and requests the H slabs of the NEXT tile while the
// current one is still being computed; bf16 (half the bytes per element, 64-column slabs) keeps 8.
template <int PREC>
struct P2cCfg {
  static constexpr int kEpi = PREC == kPrecTf32 ? 16 : 8;      // epilogue warps
  static constexpr int kThreads = 64 + 32 * kEpi;
  static constexpr int kStages = PREC == kPrecTf32 ? 2 : 3;    // 16 epilogue warps' slabs (128 KiB) leave room for two 32 KiB stages
  static constexpr int kStage = 2 * kATileBytes;                // this CTA's 128 A rows + its 128 B rows, 128 B of K each
  static constexpr int kSlabs = kEpi * 2 * kSlabBytes;          // two slabs per epilogue warp
  static constexpr int kBars = 8 * (2 * kStages + 4 + 2 * kEpi);
  static constexpr int kBias = kEpi * 64 * 4;                   // 64 bias values per warp (bf16: per half of its 128 columns)
  // no alignment slack: the kernel has no static shared memory, so the dynamic window starts 1024-byte aligned (checked)
  static constexpr int kTotal = kStages * kStage + kSlabs + ((kBars + 15) & ~15) + 16 + kBias;
  static_assert(kTotal <= 232448, "shared memory budget");
};

template <int MODE, int PREC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P2cCfg<PREC>::kThreads, 1)
mlp_gemm2cta_kernel(const __grid_constant__ TcGemmArgs g) {
  using P = PrecT<PREC>;
  using S = P2cCfg<PREC>;
  constexpr int BN = 256;
  constexpr int kEpi = S::kEpi, k2Stages = S::kStages;
  constexpr int CH = kRowBytes / (int)sizeof(typename P::T);
  constexpr int LD = CH / 32;
  constexpr int WCOLS = BN / (kEpi / 4);  // columns per epilogue warp: 64 (fp32) / 128 (bf16)
  constexpr int NSLAB = WCOLS / CH;       // = 2 slabs per warp and tile
  static_assert(NSLAB == 2, "the H prefetch below assumes two slabs per warp and tile");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = raw;
  const uint32_t slabs = tiles + k2Stages * S::kStage;
  const uint32_t bars = slabs + S::kSlabs;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * k2Stages;
  const uint32_t tfull_bar = bars + 16 * k2Stages, tempty_bar = tfull_bar + 16;
  const uint32_t h_bar = tempty_bar + 16;  // [kEpi][2]
  const uint32_t tmem_slot = bars + ((S::kBars + 15) & ~15);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  float* bias_sm = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - raw));  // [kEpi][64], private to each warp
  if (raw & 1023u) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();  // 0 = leader
  const int n_pairs = gridDim.x / 2, pair = blockIdx.x / 2;
  const int tiles_n = g.N / BN, tiles_m = (g.M + 255) / 256;
  const int per_net = tiles_m * tiles_n, total = 2 * per_net;
  const int k_blocks = g.K / P::kBK;

  if (warp == 0 && lane == 0) {
    for (int z = 0; z < 2; ++z) {
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapA[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapB[z]));
      asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapC[z]));
      if (MODE == kTcDgrad) asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapH[z]));
    }
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);   // the leader's producer arrives (with both CTAs' byte count); used on the leader only
      mbar_init(empty_bar + 8 * s, 1);  // one multicast commit per use
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 2 * kEpi);  // the epilogue warps of BOTH CTAs; used on the leader only
    }
    for (int i = 0; i < 2 * kEpi; ++i) mbar_init(h_bar + 8 * i, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised and visible cluster-wide, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = pair; t < total; t += n_pairs) {
        const int z = t / per_net, r = t - z * per_net;
        const int row_base = (r / tiles_n) * 256 + (int)cta * 128, col_base = (r % tiles_n) * BN + (int)cta * 128;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % k2Stages;
          mbar_wait(empty_bar + 8 * s, ((it / k2Stages) & 1) ^ 1);
          const uint32_t sa = tiles + s * S::kStage, sb = sa + kATileBytes;
          const uint32_t leader_full = map_to_cta(full_bar + 8 * s, 0);
          if (cta == 0) mbar_expect_tx(full_bar + 8 * s, 2 * S::kStage);  // both CTAs' A rows and B halves
          tma_load_2d_2cta(sa, &g.mapA[z], leader_full, kb * P::kBK, row_base);  // my 128 rows of the 256-row tile
          tma_load_2d_2cta(sb, &g.mapB[z], leader_full, kb * P::kBK, col_base);  // my 128 of the 256 B rows
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (cta == 0) {
      constexpr uint32_t idesc = make_idesc(P::kFmt, 256, BN, false, false);
      uint32_t it = 0, j = 0;
      for (int t = pair; t < total; t += n_pairs, ++j) {
        const uint32_t a = j & 1;
        mbar_wait(tempty_bar + 8 * a, ((j >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const uint32_t s = it % k2Stages;
          mbar_wait(full_bar + 8 * s, (it / k2Stages) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = tiles + s * S::kStage, sb = sa + kATileBytes;
#pragma unroll
            for (int k = 0; k < P::kBK / P::kUmmaK; ++k) {
              const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024);
              umma_2cta<PREC>(tmem_base + a * BN, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          __syncwarp();
          if (elect_one()) {
            umma_commit_2cta(empty_bar + 8 * s, 3);                            // ring slot free in both CTAs
            if (kb == k_blocks - 1) umma_commit_2cta(tfull_bar + 8 * a, 3);   // accumulator complete in both CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs: own 128 rows x 256 columns, two slabs per warp) =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int c_first = (ew >> 2) * WCOLS;
    const uint32_t my_slabs = slabs + ew * 2 * kSlabBytes;
    const uint32_t my_hbar = h_bar + ew * 16;
    const uint32_t lane_row = lane * kRowBytes, lane_x = lane & 7;
    auto tile_coords = [&](int t, int& z, int& srow, int& col0) {
      z = t / per_net;
      const int r = t - z * per_net;
      srow = (r / tiles_n) * 256 + (int)cta * 128 + quarter * 32;
      col0 = (r % tiles_n) * BN + c_first;
    };
    auto load_h = [&](int z, int srow, int col, uint32_t buf) {  // lane 0 only
      mbar_expect_tx(my_hbar + 8 * buf, kSlabBytes);
      tma_load_2d(my_slabs + buf * kSlabBytes, &g.mapH[z], my_hbar + 8 * buf, col, srow);
    };
    uint32_t j = 0;
    bool h0_issued = false;  // slab 0's H of the tile about to start was already requested at the end of the previous one
    for (int t = pair; t < total; t += n_pairs, ++j) {
      int z, srow, col0;
      tile_coords(t, z, srow, col0);
      const uint32_t a = j & 1;
      float* bsm = bias_sm + ew * 64;  // private: no other warp reads or writes it
      if (MODE == kTcDgrad) {
        if (lane == 0) {
          if (!h0_issued) {
            asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
            load_h(z, srow, col0, 0);
          }
          asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // buffer 1: the previous tile's second store
          load_h(z, srow, col0 + CH, 1);
        }
      } else if (WCOLS == 64) {  // fp32: the warp's 64 bias values, once per tile (the previous tile's math is behind us)
        const float* __restrict__ bp = g.bias[z] + col0;
        __syncwarp();
        bsm[lane] = __ldg(bp + lane);
        bsm[lane + 32] = __ldg(bp + lane + 32);
        __syncwarp();
      }
      if (lane == 0) mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);
      __syncwarp();
      mbar_wait(tfull_bar + 8 * a, (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + c_first;
#pragma unroll
      for (int i = 0; i < NSLAB; ++i) {
        const uint32_t slab = my_slabs + i * kSlabBytes;  // slab i of every tile lives in buffer i
        uint32_t v[LD][32];
#pragma unroll
        for (int l = 0; l < LD; ++l) tmem_ld32(tbase + i * CH + l * 32, v[l]);
        if (MODE == kTcDgrad) {
          mbar_wait(my_hbar + 8 * i, j & 1);
        } else {
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");  // the previous tile's store of slab i
          if (WCOLS == 128) {  // bf16: 64 bias values per slab
            const float* __restrict__ bp = g.bias[z] + col0 + i * CH;
            __syncwarp();
            bsm[lane] = __ldg(bp + lane);
            bsm[lane + 32] = __ldg(bp + lane + 32);
          }
          __syncwarp();
        }
#pragma unroll
        for (int l = 0; l < LD; ++l) tmem_ld_wait(v[l]);
        if (i == NSLAB - 1) {  // last TMEM read of this tile: hand the stage back to the leader's MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar + 8 * a, 0));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t addr = slab + lane_row + ((q ^ lane_x) << 4);
          uint4 o;
          if (PREC == kPrecTf32) {
            float x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) x[e] = __uint_as_float(v[0][q * 4 + e]);
            if (MODE == kTcFwd) {
              const float4 b = *reinterpret_cast<const float4*>(bsm + i * CH + q * 4);
              x[0] = elu_fast(x[0] + b.x); x[1] = elu_fast(x[1] + b.y); x[2] = elu_fast(x[2] + b.z); x[3] = elu_fast(x[3] + b.w);
            } else {
              const uint4 h = ld_shared_v4(addr);
              x[0] *= elu_grad_from_output(__uint_as_float(h.x)); x[1] *= elu_grad_from_output(__uint_as_float(h.y));
              x[2] *= elu_grad_from_output(__uint_as_float(h.z)); x[3] *= elu_grad_from_output(__uint_as_float(h.w));
            }
            o = make_uint4(__float_as_uint(round_tf32(x[0])), __float_as_uint(round_tf32(x[1])),
                           __float_as_uint(round_tf32(x[2])), __float_as_uint(round_tf32(x[3])));
          } else {
            float x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[(q >> 2) % LD][(q & 3) * 8 + e]);
            if (MODE == kTcFwd) {
              const float4 b0 = *reinterpret_cast<const float4*>(bsm + q * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(bsm + q * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = elu_fast(x[e] + bb[e]);
            } else {
              const uint4 h = ld_shared_v4(addr);
              const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                x[2 * e] *= elu_grad_from_output(__low2float(hp[e]));
                x[2 * e + 1] *= elu_grad_from_output(__high2float(hp[e]));
              }
            }
            o = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
          }
          st_shared_v4(addr, o);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&g.mapC[z], slab, col0 + i * CH, srow);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
      // dgrad: request slab 0's H of the NEXT tile now (its buffer is free as soon as this tile's first store has been
      // read; the second one may still be in flight): it travels while this warp waits for the next accumulator
      h0_issued = false;
      if (MODE == kTcDgrad && t + n_pairs < total) {
        if (lane == 0) {
          int nz, nrow, ncol;
          tile_coords(t + n_pairs, nz, nrow, ncol);
          asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
          load_h(nz, nrow, ncol, 0);
        }
        h0_issued = true;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    tc_fence_before();
  }
  // nobody frees TMEM or leaves (its shared memory is a multicast target) before both CTAs are completely done
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ---- weight gradient ---------------------------------------------------------------------------------------------
// Weight-gradient pipeline: a stage holds kWgRows = 64 reduction rows of the 128 dZ features and the BN input features
// (TMA boxes of one 128-byte feature chunk x 64 rows: 8 KiB per box in tf32 -- with 32-row boxes the 4 KiB requests, six
// per stage, bounded the 45 -> 512 layer).  The launches have at most one CTA per SM (tiles x splits <= 148), so the bytes
// in flight per SM are stages x stage size: the ring is as deep as 192 KiB of shared memory allow (three 32 KiB stages
// streamed only 64 GB/s per SM).
constexpr int kWgRows = 64;

template <int PREC, int BN>
struct WgSmem {
  static constexpr int kStage = (128 + BN) * kWgRows * (int)sizeof(typename PrecT<PREC>::T);
  static constexpr int kStages = 196608 / kStage;        // 3 (tf32, BN 128), 4 (tf32, BN 64), 6 / 8 (bf16)
  static constexpr int kOnes = 2048;                     // 16 rows x 128 B of ones (K-major B operand of the bias MMA)
  static constexpr int kTotal = kStages * kStage + kOnes + 256 /*barriers + tmem slot*/ + 1024 /*alignment slack*/;
};

template <int PREC, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_wgrad_kernel(const __grid_constant__ TcWgradArgs g) {
  using P = PrecT<PREC>;
  using T = typename P::T;
  using S = WgSmem<PREC, BN>;
  constexpr int kWgStages = S::kStages;
  constexpr int CH = kRowBytes / (int)sizeof(T);   // features per 128-byte row: 64 / 32
  constexpr int kBoxBytes = kWgRows * kRowBytes;   // one TMA box: CH features x 64 rows (8 KiB)
  constexpr int kABytes = (128 / CH) * kBoxBytes;  // the 128 dZ features of a stage
  constexpr int kTmemCols = BN == 256 ? 512 : BN == 128 ? 256 : 128;  // BN accumulator columns + 16 for the bias MMA, power of two
  extern __shared__ uint8_t smem_raw[];
  pdl_launch_dependents();
  const unsigned long long hA = g.hintA ? g.hintA : kL2EvictNormal, hB = g.hintB ? g.hintB : kL2EvictNormal;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t tiles = (raw + 1023u) & ~1023u;
  const uint32_t ones = tiles + kWgStages * S::kStage;
  const uint32_t bars = ones + S::kOnes;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kWgStages, tmem_full_bar = bars + 16 * kWgStages;
  const uint32_t tmem_slot = tmem_full_bar + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int n_tiles = g.ins_pad / BN;
  const int row_base = (blockIdx.x / n_tiles) * 128;  // dW rows = output features of the layer
  const int col_base = (blockIdx.x % n_tiles) * BN;   // dW columns = input features
  const bool with_bias = col_base == 0;                // one column tile per row tile also produces db
  // the splits sweep the minibatch together: split y takes the 64-row blocks y, y + splits, y + 2 splits, ... (from the
  // far end when `reverse`), so the launch as a whole walks the rows in one direction
  const int blocks_total = (g.rows + kWgRows - 1) / kWgRows;
  const int k_blocks = max(0, (blocks_total - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapA[z]));
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&g.mapB[z]));
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  if (warp >= 2) {  // the tile of ones, written through the generic proxy and published to the async proxy
    T* op = reinterpret_cast<T*>(smem_raw + (ones - raw));
    const T one = PREC == kPrecTf32 ? T(1.0f) : T(__float2bfloat16(1.0f));
    for (int i = threadIdx.x - 64; i < S::kOnes / (int)sizeof(T); i += kTcThreads - 64) op[i] = one;
    fence_proxy_async_smem();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kWgStages;
        mbar_wait(empty_bar + 8 * s, ((kb / kWgStages) & 1) ^ 1);
        const uint32_t sa = tiles + s * S::kStage, sb = sa + kABytes;
        mbar_expect_tx(full_bar + 8 * s, S::kStage);
        const int rb = kb * (int)gridDim.y + (int)blockIdx.y;
        const int k0 = (g.reverse ? blocks_total - 1 - rb : rb) * kWgRows;
        // MN-major operands: boxes of CH contiguous features x 64 reduction rows; rows beyond the matrix are zero-filled
        for (int h = 0; h < 128 / CH; ++h) tma_load_2d(sa + h * kBoxBytes, &g.mapA[z], full_bar + 8 * s, row_base + h * CH, k0, hA);
        for (int h = 0; h < BN / CH; ++h) tma_load_2d(sb + h * kBoxBytes, &g.mapB[z], full_bar + 8 * s, col_base + h * CH, k0, hB);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(P::kFmt, 128, BN, true, true);
    constexpr uint32_t idesc_ones = make_idesc(P::kFmt, 128, 16, true, false);
    for (int kb = 0; kb < k_blocks; ++kb) {
      const int s = kb % kWgStages;
      mbar_wait(full_bar + 8 * s, (kb / kWgStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = tiles + s * S::kStage, sb = sa + kABytes;
        const uint64_t d1 = make_smem_desc(ones, 16, 1024);  // K-major, 16 rows x 128 B, every element 1
#pragma unroll
        for (int k = 0; k < kWgRows / P::kUmmaK; ++k) {
          // MN-major: CH-feature chunks LBO = one box apart; one K-step (kUmmaK reduction rows) = kUmmaK * 128 B
          // further.  bf16: SWIZZLE_128B, 8-row reduction groups SBO = 1 KiB apart.  tf32: the tensor core takes
          // MN-major 32-bit operands only in the 32-byte-granular swizzle (4-row groups, SBO = 512 B).
          constexpr uint32_t lt = PREC == kPrecTf32 ? kLayoutSw128Base32 : kLayoutSw128;
          constexpr uint32_t sbo = PREC == kPrecTf32 ? 512 : 1024;
          const uint64_t da = make_smem_desc(sa + k * P::kUmmaK * kRowBytes, kBoxBytes, sbo, lt);
          const uint64_t db = make_smem_desc(sb + k * P::kUmmaK * kRowBytes, kBoxBytes, sbo, lt);
          umma<PREC>(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          if (with_bias) umma<PREC>(tmem_base + BN, da, d1, idesc_ones, (kb | k) != 0 ? 1u : 0u);
        }
      }
      __syncwarp();
      if (elect_one()) {
        umma_commit(empty_bar + 8 * s);
        if (kb == k_blocks - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else if (k_blocks > 0) {
    // ===================== epilogue (warps 2..9): TMEM -> red.global.add =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int HC = BN / 2, NCH = HC / 32;
    const int row = row_base + quarter * 32 + lane;
    const int c_first = half * HC;
    if (lane == 0) mbar_wait(tmem_full_bar, 0);
    __syncwarp();
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float* __restrict__ grow = g.gw[z] + (size_t)row * g.ins_pad + col_base + c_first;
    constexpr int G = NCH >= 2 ? 2 : 1;  // 32-column chunks read from TMEM at a time
#pragma unroll
    for (int i0 = 0; i0 < NCH; i0 += G) {
      uint32_t v[G][32];
#pragma unroll
      for (int i = 0; i < G; ++i) tmem_ld32(taddr + c_first + (i0 + i) * 32, v[i]);
#pragma unroll
      for (int i = 0; i < G; ++i) tmem_ld_wait(v[i]);
#pragma unroll
      for (int i = 0; i < G; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          red_add_v4(grow + (i0 + i) * 32 + q * 4, __uint_as_float(v[i][q * 4]), __uint_as_float(v[i][q * 4 + 1]),
                     __uint_as_float(v[i][q * 4 + 2]), __uint_as_float(v[i][q * 4 + 3]));
    }
    if (with_bias && half == 0) {
      uint32_t b[8];
      tmem_ld8(taddr + BN, b);
      red_add(g.gb[z] + row, __uint_as_float(b[0]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
  int prec, swz32;
};
struct TmapEntry {
  TmapKey key;
  CUtensorMap map;
};
static TmapEntry g_tmap_cache[256];
static int g_tmap_count = 0;

static int encode_tmap(CUtensorMap* map, int prec, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                       uint32_t box_inner, uint32_t box_outer, int swz32) {
  EncodeTiledFn fn = encoder();
  if (!fn) return CATB200_ERR_UNSUPPORTED;
  const uint64_t esz = prec == kPrecTf32 ? 4 : 2;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, prec == kPrecTf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? CATB200_OK : CATB200_ERR_CUDA;
}

// Encodings are memoised: the trainer reuses a handful of (pointer, shape) combinations every step.
int make_tmap(CUtensorMap* map, int prec, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer, int swz32) {
  for (int i = 0; i < g_tmap_count; ++i) {
    const TmapKey& k = g_tmap_cache[i].key;
    if (k.ptr == ptr && k.inner == inner && k.outer == outer && k.ld == ld && k.bi == box_inner && k.bo == box_outer &&
        k.prec == prec && k.swz32 == swz32) {
      *map = g_tmap_cache[i].map;
      return CATB200_OK;
    }
  }
  int rc = encode_tmap(map, prec, ptr, inner, outer, ld, box_inner, box_outer, swz32);
  if (rc == CATB200_OK) {
    const int slot = g_tmap_count < 256 ? g_tmap_count++ : 255;
    g_tmap_cache[slot].key = TmapKey{ptr, inner, outer, ld, box_inner, box_outer, prec, swz32};
    g_tmap_cache[slot].map = *map;
  }
  return rc;
}

template <int MODE, int PREC, int STAGES>
static int launch_gemm_st(const TcGemmArgs& g, cudaStream_t st) {
  using S = PCfg<PREC, MODE, STAGES>;
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm_kernel<MODE, PREC, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr = true;
  }
  const int total = 2 * ((g.M + 127) / 128) * (g.N / 128);
  CATB200_CUDA_TRY(launch_pdl(mlp_gemm_kernel<MODE, PREC, STAGES>, dim3(min(total, kNumSMs)), dim3(S::kThreads), (size_t)S::kTotal, st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// ring depth: CATB200_GEMM_STAGES = 3, 4 (default) or 5 -- 32 KiB of operands per stage
template <int MODE, int PREC>
static int launch_gemm(const TcGemmArgs& g, cudaStream_t st) {
  static int stages = 0;
  if (stages == 0) {
    const char* e = std::getenv("CATB200_GEMM_STAGES");
    stages = e ? atoi(e) : 4;
    if (stages < 3 || stages > 5) stages = 4;
  }
  if (stages == 3) return launch_gemm_st<MODE, PREC, 3>(g, st);
  if (stages == 5) return launch_gemm_st<MODE, PREC, 5>(g, st);
  return launch_gemm_st<MODE, PREC, 4>(g, st);
}

template <int MODE, int PREC, int BN>
static int launch_gemm256(const TcGemmArgs& g, cudaStream_t st) {
  using S = P2Smem<BN>;
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm256_kernel<MODE, PREC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr = true;
  }
  const int total = 2 * ((g.M + 255) / 256) * (g.N / BN);
  CATB200_CUDA_TRY(launch_pdl(mlp_gemm256_kernel<MODE, PREC, BN>, dim3(min(total, kNumSMs)), dim3(kTcThreads), (size_t)S::kTotal, st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// 256-row tiles (half the operand bytes per flop).  Measured per launch inside the step graph at 16384 rows (B200,
// profiles/README.md, round 2; 128 x 128 tiles -> 256-row tiles): forward 512 -> 256: 30.3 -> 25.0 us, forward 256 -> 128:
// 12.4 -> 11.4 us, forward 45 -> 512: 15.2 -> 19.4 us, dgrads 14.7 -> 15.8 and 34.1 -> 35.9 us.  The launches with a long
// reduction are bound by L2 -> SM operand traffic (~10 TB/s) and gain; where the epilogue dominates (K = 64, the dgrads'
// H loads) the single accumulator stage of BN = 256 loses.  Default: forward launches with K >= 256 that fill most of
// the machine.  CATB200_TILE256=1 uses them everywhere, =0 nowhere.
static bool use_tile256(int mode, int prec, int M, int N, int K) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_TILE256");
    v = (e && e[0] == '1') ? 1 : (e && e[0] == '0') ? 0 : 2;
  }
  const int bn = N % 256 == 0 ? 256 : 128;
  if (v == 0 || 2 * ((M + 255) / 256) * (N / bn) < 96) return false;
  return v == 1 || (mode == kTcFwd && prec == kPrecTf32 && K >= 256);  // (not measured with bf16 operands)
}

// CTA pairs (256 x 256 tiles, cta_group::2) for launches with N a multiple of 256 that give every pair work: opt-in
// (CATB200_PAIRS=1).  Measured on the B200 (profiles/README.md, round 2): correct, half the operand bytes per flop, but
// no faster than the 128 x 128 tiles (forward 23 vs 19.5 us, dgrad 25.4 vs 22.7 us per launch at 16384 rows).
static bool use_pairs(int M, int N) {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_PAIRS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1 && N % 256 == 0 && 2 * ((M + 255) / 256) * (N / 256) >= 64;
}

template <int MODE, int PREC>
static int launch_gemm_pairs(const TcGemmArgs& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(mlp_gemm2cta_kernel<MODE, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2cCfg<PREC>::kTotal));
    attr = true;
  }
  const int total = 2 * ((g.M + 255) / 256) * (g.N / 256);
  const int pairs = min(total, kNumSMs / 2);
  // the cluster shape is compiled into the kernel (__cluster_dims__): a plain launch with an even grid
  mlp_gemm2cta_kernel<MODE, PREC><<<dim3(2 * pairs), dim3(P2cCfg<PREC>::kThreads), (size_t)P2cCfg<PREC>::kTotal, st>>>(g);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

template <int MODE, int PREC>
static int launch_gemm_any(const TcGemmArgs& g, cudaStream_t st) {
  if (use_pairs(g.M, g.N)) return launch_gemm_pairs<MODE, PREC>(g, st);
  if (use_tile256(MODE, PREC, g.M, g.N, g.K)) {
    if (g.N % 256 == 0) return launch_gemm256<MODE, PREC, 256>(g, st);
    return launch_gemm256<MODE, PREC, 128>(g, st);
  }
  return launch_gemm<MODE, PREC>(g, st);
}

int tc_gemm_launch(int mode, int prec, const TcGemmArgs& g, cudaStream_t st) {
  const int bk = prec == kPrecTf32 ? 32 : 64;
  if (g.N % 128 || g.K % bk || g.M <= 0) return CATB200_ERR_UNSUPPORTED;
  if (mode == kTcFwd) return prec == kPrecTf32 ? launch_gemm_any<kTcFwd, kPrecTf32>(g, st) : launch_gemm_any<kTcFwd, kPrecBf16>(g, st);
  return prec == kPrecTf32 ? launch_gemm_any<kTcDgrad, kPrecTf32>(g, st) : launch_gemm_any<kTcDgrad, kPrecBf16>(g, st);
}

template <int PREC, int BN>
static int launch_wgrad(const TcWgradArgs& g, int splits, cudaStream_t st) {
  using S = WgSmem<PREC, BN>;
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(mlp_wgrad_kernel<PREC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr = true;
  }
  const dim3 grid((g.outs / 128) * (g.ins_pad / BN), splits, 2);
  CATB200_CUDA_TRY(launch_pdl(mlp_wgrad_kernel<PREC, BN>, grid, dim3(kTcThreads), (size_t)S::kTotal, st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// Column tile of the weight gradient.  256 input features per CTA (a CTA streams its 128 dZ features once per 256 instead
// of once per 128 input features) is opt-in, CATB200_WGRAD256=1: measured slower on the B200 (profiles/README.md, round 2:
// 17.0 vs 12.5 us and 30.7 vs 29.4 us per launch) -- its 96 KiB stages leave a two-deep ring.
int tc_wgrad_bn(int prec, int ins_pad) {
  static int wide = -1;
  if (wide < 0) {
    const char* e = std::getenv("CATB200_WGRAD256");
    wide = (e && e[0] == '1') ? 1 : 0;
  }
  if (wide && prec == kPrecTf32 && ins_pad % 256 == 0) return 256;
  return ins_pad % 128 == 0 ? 128 : 64;
}

int tc_wgrad_launch(int prec, const TcWgradArgs& g, int splits, cudaStream_t st) {
  if (g.outs % 128 || splits <= 0) return CATB200_ERR_UNSUPPORTED;
  const int bn = tc_wgrad_bn(prec, g.ins_pad);
  if (bn == 256) return launch_wgrad<kPrecTf32, 256>(g, splits, st);
  if (g.ins_pad % 128 == 0)
    return prec == kPrecTf32 ? launch_wgrad<kPrecTf32, 128>(g, splits, st) : launch_wgrad<kPrecBf16, 128>(g, splits, st);
  if (g.ins_pad % 64 == 0)
    return prec == kPrecTf32 ? launch_wgrad<kPrecTf32, 64>(g, splits, st) : launch_wgrad<kPrecBf16, 64>(g, splits, st);
  return CATB200_ERR_UNSUPPORTED;
}

}  // namespace catb200

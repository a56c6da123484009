// Fused three-layer forward of one MLP (obs_pad -> h1 -> h2 -> h3, ELU) on tcgen05: activations never leave the SM.
//
// One CTA = 128 rows of one net (blockIdx.z: 0 critic, 1 actor).  The layer-l activation tile lives in shared
// memory as the K-major SWIZZLE_128B operand of layer l+1 (8 / 4 boxes of [128 rows x 64 features]); the very same
// boxes leave for global memory through TMA tensor stores (the backward pass and the head kernel need H1..H3), so
// the epilogue writes every value exactly once.  Weights stream through a 4-slot ring of 16 KiB tiles
// ([128 output rows x 64 inputs]); accumulators use all 512 TMEM columns for layer 1, 256 / 128 afterwards.
//   warp 0     : TMA producer (X tile, then 4 + 16 + 4 weight tiles)
//   warp 1     : TMEM allocator + MMA issuer (one elected lane, tcgen05.mma 128x128x16)
//   warps 2..9 : epilogues: TMEM -> bias + ELU -> swizzled smem -> TMA store, then hand the tile to the MMA warp
// Replaces three tc_gemm launches of the forward pass (opt-in: CATB200_FUSED_FWD=1).
#include <cuda.h>

#include "common.cuh"
#include "mma.cuh"
#include "tc_gemm.cuh"
#include "tc_ptx.cuh"

namespace catb200 {

constexpr int kF3Threads = 320;
constexpr int kF3Slots = 4;
constexpr int kF3Tile = 16384;  // one [128 x 64] bf16 box
constexpr int kF3MaxH1 = 512;

struct F3Smem {
  static constexpr int kAct = 0;                            // 8 boxes: H1 (then H2 / H3 in the first boxes)
  static constexpr int kX = kAct + 8 * kF3Tile;             // X tile
  static constexpr int kRing = kX + kF3Tile;                // weight ring
  static constexpr int kBias = kRing + kF3Slots * kF3Tile;  // b1 | b2 | b3 (fp32)
  static constexpr int kBars = kBias + (512 + 256 + 128) * 4;
  static constexpr int kTotal = kBars + 256 + 1024;         // + barriers + alignment slack
};

__global__ void __launch_bounds__(kF3Threads, 1)
fwd3_kernel(const __grid_constant__ Fwd3Args g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t act = base + F3Smem::kAct, xs = base + F3Smem::kX, ring = base + F3Smem::kRing;
  float* bias_sm = reinterpret_cast<float*>(base_ptr + F3Smem::kBias);
  const uint32_t bars = base + F3Smem::kBars;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kF3Slots, x_full = bars + 64, acc_full = bars + 72, epi_done = bars + 80;
  const uint32_t tmem_slot = bars + 96;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(base_ptr + F3Smem::kBars + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.z;
  const int row_base = blockIdx.x * 128;
  const int H1 = g.h1, H2 = g.h2, H3 = g.h3;  // 512, 256, 128 (multiples of 128, H1 <= 512)
  const int nt1 = H1 / 128, nt2 = H2 / 128, kb2 = H1 / 64, kb3 = H2 / 64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kF3Slots; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    mbar_init(x_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(epi_done, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    for (int c = et; c < H1; c += 256) bias_sm[c] = __ldg(g.bias[z][0] + c);
    for (int c = et; c < H2; c += 256) bias_sm[512 + c] = __ldg(g.bias[z][1] + c);
    for (int c = et; c < H3; c += 256) bias_sm[768 + c] = __ldg(g.bias[z][2] + c);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(x_full, kF3Tile);
      tma_load_2d(xs, &g.mapX[z], x_full, 0, row_base);
      int i = 0;
      auto push = [&](const CUtensorMap* map, int k0, int n0) {
        const int s = i % kF3Slots;
        mbar_wait(empty_bar + 8 * s, ((i / kF3Slots) & 1) ^ 1);
        mbar_expect_tx(full_bar + 8 * s, kF3Tile);
        tma_load_2d(ring + s * kF3Tile, map, full_bar + 8 * s, k0, n0);
        ++i;
      };
      for (int j = 0; j < nt1; ++j) push(&g.mapW[z][0], 0, j * 128);
      for (int n = 0; n < nt2; ++n)
        for (int kb = 0; kb < kb2; ++kb) push(&g.mapW[z][1], kb * 64, n * 128);
      for (int kb = 0; kb < kb3; ++kb) push(&g.mapW[z][2], kb * 64, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc(128, 128, false, false);
    int i = 0;
    // D[tmem_base + d_col] (+)= A(a_addr: [128 x 64] K-major box) * B(ring slot)^T, 4 K-steps of 16
    auto mma_tile = [&](uint32_t a_addr, int d_col, bool first) {
      const int s = i % kF3Slots;
      mbar_wait(full_bar + 8 * s, (i / kF3Slots) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = ring + s * kF3Tile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + d_col, make_smem_desc(a_addr + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                    (first && k == 0) ? 0u : 1u);
      }
      __syncwarp();
      if (elect_one()) umma_commit(empty_bar + 8 * s);
      __syncwarp();
      ++i;
    };
    mbar_wait(x_full, 0);
    for (int j = 0; j < nt1; ++j) mma_tile(xs, j * 128, true);  // layer 1: K = 64 (one box)
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
    mbar_wait(epi_done, 0);  // H1 is in shared memory, TMEM drained
    tc_fence_after();
    for (int n = 0; n < nt2; ++n)
      for (int kb = 0; kb < kb2; ++kb) mma_tile(act + kb * kF3Tile, n * 128, kb == 0);
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
    mbar_wait(epi_done, 1);  // H2 is in shared memory
    tc_fence_after();
    for (int kb = 0; kb < kb3; ++kb) mma_tile(act + kb * kF3Tile, 0, kb == 0);
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else {
    // ===================== epilogues (warps 2..9) =====================
    const int quarter = warp & 3;        // TMEM lanes 32*quarter .. +31
    const int part = (warp - 2) >> 2;    // which half of the layer's columns
    const int trow = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int widths[3] = {H1, H2, H3};
    const int bias_off[3] = {0, 512, 768};
#pragma unroll 1
    for (int layer = 0; layer < 3; ++layer) {
      if (lane == 0) mbar_wait(acc_full, layer & 1);
      __syncwarp();
      mbar_wait(acc_full, layer & 1);
      tc_fence_after();
      const int boxes = widths[layer] / 128;  // 64-column boxes per warp (the two parts split the layer)
      const float* bsm = bias_sm + bias_off[layer];
#pragma unroll 1
      for (int bi = 0; bi < boxes; ++bi) {
        const int kb = part * boxes + bi;  // box index = columns [kb*64, kb*64+64)
        uint32_t v[2][32];
        tmem_ld32(taddr + kb * 64, v[0]);
        tmem_ld32(taddr + kb * 64 + 32, v[1]);
        tmem_ld_wait(v[0]);
        tmem_ld_wait(v[1]);
        const uint32_t box = act + kb * kF3Tile;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int cc = q * 8 + e * 2;
              const float x0 = __uint_as_float(v[h][cc]) + bsm[kb * 64 + h * 32 + cc];
              const float x1 = __uint_as_float(v[h][cc + 1]) + bsm[kb * 64 + h * 32 + cc + 1];
              op[e] = pack_bf16x2(elu_fast(x0), elu_fast(x1));
            }
            st_shared_v4(box + trow * 128 + (((h * 4 + q) ^ (trow & 7)) << 4), o);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&g.mapH[z][layer], box + quarter * 4096, kb * 64, row_base + quarter * 32);
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
        __syncwarp();
      }
      // the stores must have finished reading shared memory before the next layer's epilogue (any warp) rewrites
      // these boxes; every warp drains its own bulk group before it signals
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
      __syncwarp();
      tc_fence_before();
      if (layer < 2 && lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(epi_done) : "memory");
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int fwd3_launch(const Fwd3Args& g, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    CATB200_CUDA_TRY(cudaFuncSetAttribute(fwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F3Smem::kTotal));
    attr = true;
  }
  if (g.h1 % 128 || g.h2 % 128 || g.h3 != 128 || g.h1 > kF3MaxH1 || g.h2 > 256) return CATB200_ERR_UNSUPPORTED;
  CATB200_CUDA_TRY(launch_pdl(fwd3_kernel, dim3((g.M + 127) / 128, 1, 2), dim3(kF3Threads), (size_t)F3Smem::kTotal, st, g));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // namespace catb200

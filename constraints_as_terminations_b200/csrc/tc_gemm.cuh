// Internal interface of the tcgen05 GEMMs (tc_gemm.cu), used by mlp.cu.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "mma.cuh"

namespace catb200 {

enum TcMode { kTcFwd = 0, kTcDgrad = 1 };

// forward / dgrad, both nets per launch (index 0 = critic, 1 = actor).  "row" below = one 128-byte shared-memory row:
// CH = 64 bf16 or 32 fp32 elements.
struct TcGemmArgs {
  CUtensorMap mapA[2];   // A [M, K]: load boxes CH (K) x 128 rows
  CUtensorMap mapB[2];   // B [N, K]: load boxes CH (K) x 128 rows
  CUtensorMap mapC[2];   // output C [M, N]: store boxes CH (cols) x 32 (rows), SWIZZLE_128B
  CUtensorMap mapH[2];   // dgrad: forward activation H [M, N] whose ELU' scales the result, same boxes as mapC
  const float* bias[2];  // fwd: [N]
  int M, N, K;           // rows, output features, reduction length
  int reverse;           // mlp_gemm_kernel: walk the row tiles from the last to the first (see "row order" in tc_gemm.cu)
  unsigned long long hintA, hintB, hintC, hintH;  // L2 eviction-priority hints of the TMA loads / stores (tc_ptx.cuh); 0 = normal
};

// weight gradient dW[outs, ins] += dZ[rows, outs]^T Hin[rows, ins] over a range of minibatch rows, and
// db[outs] += column sums of dZ (a second, 8-column MMA against a tile of ones).
struct TcWgradArgs {
  CUtensorMap mapA[2];  // dZ  [rows, outs]: load boxes CH (features) x kBK (rows); tf32: SWIZZLE_128B_ATOM_32B
  CUtensorMap mapB[2];  // Hin [rows, ins_pad]: same box shape
  float* gw[2];         // fp32 accumulators [outs, ins_pad] (16-byte aligned rows): red.global.add.v4.f32
  float* gb[2];         // bias gradient [outs]
  int outs, ins_pad, rows;
  int m_range;          // minibatch rows per split (multiple of kBK): sizes the split count only
  int reverse;          // sweep the minibatch rows from the last block to the first
  unsigned long long hintA, hintB;  // L2 eviction-priority hints of the operand loads; 0 = normal
};

// 2-D tensor map over a row-major [outer, inner] matrix of bf16 (prec 0) or fp32 (prec 1) elements with leading
// dimension ld (elements); box = [box_outer, box_inner], SWIZZLE_128B (swz32 != 0: SWIZZLE_128B_ATOM_32B, the layout
// MN-major 32-bit tensor-core operands need).  Encodings are memoised.
int make_tmap(CUtensorMap* map, int prec, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer, int swz32 = 0);
int tc_gemm_launch(int mode, int prec, const TcGemmArgs& g, cudaStream_t st);
int tc_wgrad_launch(int prec, const TcWgradArgs& g, int splits, cudaStream_t st);
int tc_wgrad_bn(int prec, int ins_pad);  // input features per weight-gradient CTA: 256, 128 or 64

}  // namespace catb200

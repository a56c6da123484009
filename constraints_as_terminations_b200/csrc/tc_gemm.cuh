// Internal interface of the tcgen05 GEMM (tc_gemm.cu), used by mlp.cu.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "mma.cuh"

namespace catb200 {

enum TcMode { kTcFwd = 0, kTcDgrad = 1, kTcWgrad = 2 };

struct TcGemmArgs {
  CUtensorMap mapA[2];  // per net; fwd/dgrad: A [M, K] boxes 64(K) x 128; wgrad: dZ [rows, outs] boxes 64 x 64
  CUtensorMap mapB[2];  // fwd/dgrad: B [N, K] boxes 64(K) x 128; wgrad: Hin [rows, ins_pad] boxes 64 x 64
  CUtensorMap mapC[2];  // fwd/dgrad output C [M, N]: store boxes 64 (cols) x 32 (rows), SWIZZLE_128B
  bf16* C[2];           // fwd/dgrad output [M, N], ldc
  const float* bias[2]; // fwd
  const bf16* H[2];     // dgrad: forward activation whose ELU' scales the result (same layout as C)
  float* dbias[2];      // dgrad: += column sums
  float* part[2];       // wgrad: fp32 partials [splits, M(outs), N(ins_pad)]
  int ldc;
  int M, N, K;          // fwd/dgrad: rows, output features, reduction; wgrad: outs, ins_pad, minibatch rows
  int m_range;          // wgrad: minibatch rows per split (multiple of 64)
  int stages;           // ring depth, filled by tc_gemm_launch
};

// fused three-layer forward (tc_fwd3.cu)
struct Fwd3Args {
  CUtensorMap mapX[2];     // X [M, obs_pad = 64]: load box 64 x 128
  CUtensorMap mapW[2][3];  // W_l [out_l, in_pad_l]: load boxes 64 (inputs) x 128 (outputs)
  CUtensorMap mapH[2][3];  // H_l [M, out_l]: store boxes 64 (features) x 32 (rows)
  const float* bias[2][3];
  int M, h1, h2, h3;
};
int fwd3_launch(const Fwd3Args& g, cudaStream_t st);

int make_tmap_bf16(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                   uint32_t box_outer);
int tc_gemm_launch(int mode, const TcGemmArgs& g, int splits, cudaStream_t st);

}  // namespace catb200

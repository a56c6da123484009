// Shared helpers for libcatb200 (sm_100a).  Internal header; the public ABI is include/catb200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "catb200.h"

namespace catb200 {

extern thread_local cudaError_t g_last_cuda_error;

inline int cuda_status(cudaError_t e) {
  if (e == cudaSuccess) return CATB200_OK;
  g_last_cuda_error = e;
  return CATB200_ERR_CUDA;
}

#define CATB200_CUDA_TRY(expr)                         \
  do {                                                 \
    cudaError_t _e = (expr);                           \
    if (_e != cudaSuccess) return ::catb200::cuda_status(_e); \
  } while (0)

// every kernel launch of the library passes through here; the count backs bench.py's `gpu_launches`
extern unsigned long long g_launch_count;
#define CATB200_LAUNCH_CHECK()                     \
  do {                                             \
    ++::catb200::g_launch_count;                   \
    CATB200_CUDA_TRY(cudaPeekAtLastError());       \
  } while (0)

// mlp.cu: refresh the compute copies (W and W^T of the hidden layers, operand precision) from the fp32 master parameters
int launch_cast_weights(const catb200_mlp_dims_t* dims, const float* params, void* wc, cudaStream_t st);

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// The update path is a chain of ~14 short dependent kernels per minibatch.  With PDL the next kernel's CTAs
// are launched (and run their prologue: barrier init, TMEM allocation, descriptor prefetch) while the
// previous kernel drains; `pdl_wait()` then blocks until the previous grid has completed and its writes are
// visible, so every kernel still only reads finished data.  Opt-in with CATB200_PDL=1 (measured neutral
// when the chain already replays from a CUDA graph).
bool pdl_enabled();

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// Monotone map float -> uint32 (unsigned compare == float compare); 0 is below every real number,
// so a zero-initialised scratch word is the identity of atomicMax.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(b);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Last-block-done ticket: returns true in exactly one block (the last to arrive), for all its threads.
// Callers must have made their global writes visible (__threadfence) before calling.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int total_blocks) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == total_blocks - 1);
    if (s_last) *counter = 0u;  // leave the workspace clean for the next launch
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// Two-level variant for grids of tens of thousands of CTAs: one shared counter would queue every CTA's
// atomic on a single L2 address (tens of ns each).  CTAs first count inside one of kTicketGroups groups;
// only the last CTA of each group touches the global counter.  `counters` = 1 + kTicketGroups words, zero
// between launches (left clean).  Returns true in exactly one block, for all its threads.
constexpr unsigned int kTicketGroups = 32;

__device__ __forceinline__ bool last_block_ticket_grouped(unsigned int* counters, unsigned int total_blocks) {
  __shared__ bool s_last_g;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int groups = min(kTicketGroups, total_blocks);
    const unsigned int g = blockIdx.x % groups;
    const unsigned int members = total_blocks / groups + (g < total_blocks % groups ? 1u : 0u);
    bool last = false;
    if (atomicAdd(&counters[1 + g], 1u) == members - 1) {
      counters[1 + g] = 0u;
      __threadfence();
      if (atomicAdd(&counters[0], 1u) == groups - 1) {
        counters[0] = 0u;
        last = true;
      }
    }
    s_last_g = last;
  }
  __syncthreads();
  if (s_last_g) __threadfence();
  return s_last_g;
}

}  // namespace catb200

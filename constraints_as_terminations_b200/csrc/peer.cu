// Multi-GPU gradient exchange over NVLink peer memory (sm_100a): a one-shot all-reduce fused with the global gradient
// norm, so that the optimizer step of every rank stays inside one CUDA graph (no NCCL call on the critical path).
//
// Every rank owns ONE peer-visible allocation (cudaMalloc + cudaIpcGetMemHandle; the peers map it with
// cudaIpcOpenMemHandle): [flags: 64 x u32][gradient arena 0][gradient arena 1].  The flat fp32 gradient of minibatch k
// is accumulated straight into arena k & 1 by the backward kernels.  grad_allreduce_norm_kernel then
//   1. announces "my arena k & 1 is complete" by storing the epoch number into its slot of every peer's flag row
//      (st.release.sys over NVLink), and waits until every peer's announcement of this epoch has arrived in its own row;
//   2. reads the W arenas (its own + W-1 remote ones, 16-byte loads over NVLink / NVSwitch), sums them in rank order --
//      every rank computes bit-identical sums -- writes the result to a private buffer for the Adam kernel and
//      accumulates the squared norm on the way; the last CTA turns it into the clip coefficient and the Adam bias
//      corrections (the job of grad_norm_kernel on a single GPU);
//   3. zeroes its own OTHER arena: having seen everybody's announcement of epoch e, nobody can still be reading the arena
//      of epoch e - 1, and the next minibatch accumulates into it.  One barrier per optimizer step is all it takes.
// Waiting is bounded: a peer that never arrives raises an error flag after ~2 s instead of hanging the GPU.
#include <cstring>

#include "common.cuh"
#include "optim.cuh"

namespace catb200 {

struct PeerArgs {
  const float* arena[kPeerMax];  // gradient arena of this epoch's parity on every rank (index = rank)
  uint32_t* flags[kPeerMax];     // flag row of every rank
  float* zero_arena;             // this rank's other arena
  float* out;                    // private summed gradient [n]
  long long n;
  int rank, world;
  unsigned int* epoch;           // device-local epoch counter (number of all-reduces done)
  int* err;                      // device-local error flag
  unsigned int parity;           // arena parity the host chose: must equal *epoch & 1
  float grad_scale, max_norm, beta1, beta2;
  int* step;
  float* grad_norm_out;
  OptScratch* sc;
};

__global__ void __launch_bounds__(256) grad_allreduce_norm_kernel(const __grid_constant__ PeerArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const unsigned int e = *a.epoch + 1;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  if (blockIdx.x == 0 && threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.rank, e);  // "rank a.rank has finished arena e & 1" into the peer's row
  }
  if ((int)threadIdx.x < a.world && (int)threadIdx.x != a.rank) {
    const uint32_t* mine = a.flags[a.rank] + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) < e) {
      if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz: give up loudly, do not hang the GPU
        s_ok = 0;
        atomicExch(a.err, 1);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (((*a.epoch) & 1u) != a.parity && threadIdx.x == 0) atomicExch(a.err, 2);
  double s = 0.0;
  const long long n4 = a.n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int p = 0; p < a.world; ++p) {  // rank order on every rank: identical sums everywhere
      const float4 v = __ldcv(reinterpret_cast<const float4*>(a.arena[p]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(a.out)[i] = acc;
    reinterpret_cast<float4*>(a.zero_arena)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const double x = (double)(acc.x * a.grad_scale), y = (double)(acc.y * a.grad_scale), z = (double)(acc.z * a.grad_scale),
                 w = (double)(acc.w * a.grad_scale);
    s += x * x + y * y + z * z + w * w;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {  // scalar tail
    float acc = 0.f;
    for (int p = 0; p < a.world; ++p) acc += __ldcv(a.arena[p] + i);
    a.out[i] = acc;
    a.zero_arena[i] = 0.f;
    const double x = (double)(acc * a.grad_scale);
    s += x * x;
  }
  s = warp_sum(s);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(&a.sc->sumsq, t);
  }
  if (last_block_ticket(&a.sc->ticket, gridDim.x)) {
    if (threadIdx.x == 0) {
      const double tot = __longlong_as_double(atomicExch((unsigned long long*)&a.sc->sumsq, 0ull));
      const float norm = (float)sqrt(tot);
      a.sc->clip_coef = fminf(a.max_norm / (norm + 1e-6f), 1.0f);  // torch.nn.utils.clip_grad_norm_
      a.sc->total_norm = norm;
      const int t = *a.step + 1;
      *a.step = t;
      const double bc1 = 1.0 - pow((double)a.beta1, (double)t), bc2 = 1.0 - pow((double)a.beta2, (double)t);
      a.sc->step_size_scale = (float)(1.0 / bc1);
      a.sc->bc2_sqrt = (float)sqrt(bc2);
      if (a.grad_norm_out) *a.grad_norm_out = norm;
      *a.epoch = e;  // every CTA has read the old value long ago (it is the last one to finish)
    }
  }
}

}  // namespace catb200

using namespace catb200;

extern "C" {

size_t catb200_peer_arena_bytes(int64_t n_params) {
  const size_t n_pad = ((size_t)n_params + 63) / 64 * 64;
  return kFlagWords * 4 + 3 * n_pad * 4;  // flags, two arenas, the summed gradient of the reduce-scatter path
}

int catb200_peer_alloc(size_t bytes, void** ptr, uint8_t* ipc_handle64) {
  if (!ptr || !ipc_handle64 || bytes == 0) return CATB200_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CATB200_CUDA_TRY(cudaMalloc(ptr, bytes));
  CATB200_CUDA_TRY(cudaMemset(*ptr, 0, bytes));
  CATB200_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  CATB200_CUDA_TRY(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(ipc_handle64, &h, 64);
  return CATB200_OK;
}

int catb200_peer_open(const uint8_t* ipc_handle64, void** ptr) {
  if (!ptr || !ipc_handle64) return CATB200_ERR_INVALID_ARGUMENT;
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, 64);
  CATB200_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CATB200_OK;
}

int catb200_peer_close(void* ptr) {
  if (!ptr) return CATB200_ERR_INVALID_ARGUMENT;
  CATB200_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return CATB200_OK;
}

int catb200_peer_free(void* ptr) {
  if (!ptr) return CATB200_ERR_INVALID_ARGUMENT;
  CATB200_CUDA_TRY(cudaFree(ptr));
  return CATB200_OK;
}

int catb200_grad_allreduce_norm(void* const* peer_bases, int32_t rank, int32_t world, int64_t n_params, int32_t parity,
                                float* grad_sum, float grad_scale, float max_grad_norm, float beta1, float beta2,
                                int32_t* step_dev, float* grad_norm_out, void* opt_ws, uint32_t* epoch_dev, int32_t* err_dev,
                                void* stream) {
  if (!peer_bases || world < 1 || world > kPeerMax || rank < 0 || rank >= world || n_params <= 0 || !grad_sum || !step_dev ||
      !opt_ws || !epoch_dev || !err_dev || (parity != 0 && parity != 1))
    return CATB200_ERR_INVALID_ARGUMENT;
  const size_t n_pad = ((size_t)n_params + 63) / 64 * 64;
  PeerArgs a = {};
  for (int p = 0; p < world; ++p) {
    if (!peer_bases[p]) return CATB200_ERR_INVALID_ARGUMENT;
    char* base = static_cast<char*>(peer_bases[p]);
    a.flags[p] = reinterpret_cast<uint32_t*>(base);
    a.arena[p] = reinterpret_cast<const float*>(base + kFlagWords * 4) + (size_t)parity * n_pad;
  }
  a.zero_arena = reinterpret_cast<float*>(static_cast<char*>(peer_bases[rank]) + kFlagWords * 4) + (size_t)(parity ^ 1) * n_pad;
  a.out = grad_sum; a.n = n_params; a.rank = rank; a.world = world;
  a.epoch = epoch_dev; a.err = err_dev; a.parity = (unsigned int)parity;
  a.grad_scale = grad_scale; a.max_norm = max_grad_norm; a.beta1 = beta1; a.beta2 = beta2;
  a.step = step_dev; a.grad_norm_out = grad_norm_out; a.sc = static_cast<OptScratch*>(opt_ws);
  CATB200_CUDA_TRY(launch_pdl(grad_allreduce_norm_kernel, dim3(kNumSMs), dim3(256), 0, as_stream(stream), a));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

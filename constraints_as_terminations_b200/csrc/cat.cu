// Per-step CaT path on sm_100a: constraint terms -> termination probabilities.
//
// Replaces, in one launch (up to ~9.5 k envs: kEvalFused -- evaluation, grid barrier, apply phase straight out of the
// shared-memory tiles) or two (beyond), the ~250-300 eager kernels + 13 host syncs of the reference's
// ConstraintManager.compute() (U/cat/constraint_manager.py:213-229 driving constraints.py:23-235 and
// CaT.add/get_probs :39-82) and the reward/dones lines of CaTEnv.step (U/cat/cat_env.py:102-121).
//
// Data flow (N envs in tiles of 32, K constraint columns, S statistics slots):
//   cat_eval_kernel : persistent CTAs walk the 32-env tiles (lane = env, up to 16 warps per CTA).  The tile's rows
//                     of every source tensor are copied verbatim into shared memory by the bulk async-copy engine
//                     (cp.async.bulk + mbarrier; every source has its own issuing lane, no per-element staging
//                     instructions), double buffered: tile i+1 lands while tile i is evaluated.  Contact-force
//                     peaks are computed once per (history tensor, body) pair into a shared table (dealt to the
//                     warps); then every TERM is evaluated by one warp -- its prologue (sources, scalars, command
//                     gate) runs once per tile, the op switch sits outside the column loop, terms are ranked by
//                     cost at plan time and dealt in snake order -- which writes the raw constraints into a shared
//                     [K][32] tile and the column maxima over the 32 envs (one CREDUX.MAX.F32 per column) into a
//                     shared row.  The tile leaves as ONE bulk async store into the workspace (C_T, tile-major
//                     [n_tiles][K][32]); at the end each CTA issues at most one atomicMax per column into one of up
//                     to 64 scratch rows and the last CTA (two-level ticket) applies the clamp + Polyak update to
//                     running_max[K] (:55-61).
//   cat_apply_kernel: one CTA per 32 envs (lane = env), the warps split the statistics slots.  Reads the tile's K
//                     constraint values back (coalesced, L2 hits), maps violations to probabilities (:64-72),
//                     takes the per-term and overall row max (:82,:225), updates the two per-term episode
//                     statistics (:226-227) and writes cstr_prob plus, optionally, the scaled reward / float dones
//                     and -- fused reset -- the episode means of the envs flagged in reset_buf (:190-211).
//
// The cross-env column max is a true global dependency (probability of env i depends on the max over
// all envs of this step), hence two phases.  HBM-bound streaming work: no tensor cores involved.
#include <cstdlib>

#include "common.cuh"

namespace catb200 {

thread_local cudaError_t g_last_cuda_error = cudaSuccess;
unsigned long long g_launch_count = 0;

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("CATB200_PDL");
    v = (e && e[0] == '1') ? 1 : 0;  // opt-in: measured neutral inside CUDA graphs on B200 (profiles/)
  }
  return v == 1;
}

constexpr int kTile = 32;          // envs per CTA in the eval and apply kernels (one per lane)
constexpr int kEvalMaxWarps = 16;  // eval kernel: 1, 2, 4, 8 or 16 warps share a tile (chosen per launch)
constexpr int kApplyMaxWarps = 16; // apply kernel: one warp per statistics slot, at most 16 (chosen per launch)
constexpr int kApplyThreads = 64;   // cat_probs_kernel (debug matrix): one thread per env

// sqrt(x^2 + y^2 + z^2) the way torch.norm reduces a short contiguous dim on CPU and CUDA:
// sequential fused multiply-adds from a zero accumulator, then a correctly rounded sqrt.
__device__ __forceinline__ float sumsq3(float x, float y, float z) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  return __fmaf_rn(z, z, acc);
}
__device__ __forceinline__ float norm3(float x, float y, float z) { return __fsqrt_rn(sumsq3(x, y, z)); }
__device__ __forceinline__ float norm2(float x, float y) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  return __fsqrt_rn(acc);
}

// Same-address atomics serialise in L2 (tens of ns each): with one scratch word per column, 32 768 CTAs
// (1 M envs) would queue 32 768 deep.  CTAs therefore fold their column maxima into one of kMaxGroups
// scratch rows (blockIdx % groups) and the last CTA reduces the rows.
constexpr int kMaxGroups = 64;

struct CatWorkspace {
  // layout inside the caller's workspace (all 256-byte aligned)
  unsigned int* ticket;   // 1 + kTicketGroups words (padded); word 48: generation flag of the single-launch kernel's grid barrier
  uint32_t* colmax;       // [kMaxGroups][CATB200_MAX_COLS] ordered-float column maxima, 0 between launches
  float* c_t;             // [n_tiles][K][32] raw constraints, tile-major
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ __device__ inline CatWorkspace carve(void* base, int num_envs) {
  CatWorkspace w;
  char* p = static_cast<char*>(base);
  w.ticket = reinterpret_cast<unsigned int*>(p);
  w.colmax = reinterpret_cast<uint32_t*>(p + 256);
  w.c_t = reinterpret_cast<float*>(p + 256 + align256(sizeof(uint32_t) * CATB200_MAX_COLS * kMaxGroups));
  (void)num_envs;
  return w;
}

enum EvalMode { kEvalStep = 0, kEvalRowMajor = 1, kEvalFused = 2 };

// what the apply phase reads and writes besides the constraint tile (cat_apply_kernel's arguments; kEvalFused carries them
// into the single-launch kernel)
struct ApplyArgs {
  float* episode_sums; float* mean_values; float* cstr_prob;
  const float* raw_reward; const uint8_t* reset_buf; float* reward_out; float* dones_out;
  const int64_t* episode_length; float* reset_out; struct ResetScratch* rsc;
};

// max over the warp of an fp32 value: one CREDUX.MAX.F32 on sm_100a
__device__ __forceinline__ float warp_max_f32(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;\n" : "=f"(m) : "f"(v));
  return m;
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// shared-state-space accesses at 32-bit addresses: no generic-address arithmetic in the column loops.  The loads
// are deliberately NOT volatile so that the compiler may hoist the loads of the next columns above the stores /
// warp reduction of the current one (the column chains are independent; with volatile loads each column was one
// serial LDS -> FADD -> STS -> CREDUX chain).  They only ever read data that is constant while they can execute:
// every address derives from an image base that is laundered through a volatile asm AFTER the barrier / mbarrier
// wait that publishes the data, so the data dependence keeps them behind it.
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t launder(uint32_t x) {
  asm volatile("mov.u32 %0, %0;\n" : "+r"(x)::"memory");
  return x;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory"); }

// ---- per-term column evaluation ----------------------------------------------------------------------
// A term is evaluated by ONE warp (lane = env): its prologue (sources, scalars, command gate) runs once per tile
// and the column loop is a handful of instructions per column.
struct TermCtx {
  const uint8_t* ids;  // joint / body / peak-slot ids of the term (kernel-parameter memory)
  int n_ids, n_cols;
  uint32_t x0, x1;     // shared addresses of this lane's row of source 0 / source 1
  uint32_t u1;         // source 1 holds bytes (bool tensor)
  uint32_t peaks;      // shared address of peak_table[0][lane]
  uint32_t out;        // shared address of tile[col0][lane]
  uint32_t colmax;     // shared address of tilemax[col0]
  float p0, p1, p2, gate;
  bool live, lane0;
};

template <int MODE>
__device__ __forceinline__ void emit(const TermCtx& c, int lc, float v) {
  sts_f32(c.out + lc * (kTile * 4), v);
  if (MODE != kEvalRowMajor) {
    const float m = warp_max_f32(c.live ? v : -INFINITY);
    if (c.lane0) sts_f32(c.colmax + lc * 4, m);  // this tile's column maximum; folded into the CTA's in phase C
  }
}

// Operation order follows the cited reference lines; every intermediate is rounded to fp32 exactly where
// torch materialises a tensor.
template <int OP>
__device__ __forceinline__ float column_value(const TermCtx& c, int lc) {
  const uint32_t id = c.ids[lc];
  switch (OP) {
    case CATB200_OP_ABS_MINUS:  // constraints.py:30,64,75,85
      return __fsub_rn(fabsf(lds_f32(c.x0 + id * 4)), c.p0);
    case CATB200_OP_ABSDIFF_MINUS:  // constraints.py:176-181
      return __fsub_rn(fabsf(__fsub_rn(lds_f32(c.x0 + id * 4), lds_f32(c.x1 + id * 4))), c.p0);
    case CATB200_OP_ABSDIFF_MINUS_GATE_Y:  // constraints.py:42-53
      return __fmul_rn(__fsub_rn(fabsf(__fsub_rn(lds_f32(c.x0 + id * 4), lds_f32(c.x1 + id * 4))), c.p0), c.gate);
    case CATB200_OP_ACTION_RATE:  // constraints.py:191-198 (true division by step_dt)
      return __fsub_rn(__fdiv_rn(fabsf(__fsub_rn(lds_f32(c.x0 + id * 4), lds_f32(c.x1 + id * 4))), c.p1), c.p0);
    case CATB200_OP_COMPONENT_GT:  // constraints.py:94
      return lds_f32(c.x0 + id * 4) > c.p0 ? 1.0f : 0.0f;
    case CATB200_OP_CONTACT_ANY: {  // constraints.py:103-110; ids index the peak table
      bool any = false;
      for (int b = 0; b < c.n_ids; ++b) any |= lds_f32(c.peaks + c.ids[b] * (kTile * 4)) > c.p0;
      return any ? 1.0f : 0.0f;
    }
    case CATB200_OP_NORM2_MINUS:  // constraints.py:119
      return __fsub_rn(norm2(lds_f32(c.x0), lds_f32(c.x0 + 4)), c.p0);
    case CATB200_OP_AIR_TIME: {  // constraints.py:129-141
      const bool touchdown = c.u1 ? lds_u8(c.x1 + id) != 0u : lds_f32(c.x1 + id * 4) != 0.0f;
      return __fmul_rn(__fmul_rn(__fsub_rn(c.p0, lds_f32(c.x0 + id * 4)), touchdown ? 1.0f : 0.0f), c.gate);
    }
    case CATB200_OP_N_CONTACT: {  // constraints.py:151-168
      int n = 0;
      for (int b = 0; b < c.n_ids; ++b) n += lds_f32(c.peaks + c.ids[b] * (kTile * 4)) > c.p2 ? 1 : 0;
      return __fmul_rn(fabsf((float)n - c.p0), c.gate);
    }
    case CATB200_OP_FORCE_PEAK_MINUS:  // constraints.py:207-210
      return __fsub_rn(lds_f32(c.peaks + id * (kTile * 4)), c.p0);
    case CATB200_OP_LIMIT_MINUS:  // constraints.py:220
      return __fsub_rn(c.p0, lds_f32(c.x0 + id * 4));
    case CATB200_OP_ABS_MINUS_GATE_STILL:  // constraints.py:231-235
      return __fmul_rn(__fsub_rn(fabsf(lds_f32(c.x0 + id * 4)), c.p0), c.gate);
    default:
      return 0.0f;
  }
}

// Columns are evaluated four at a time: all their loads are issued before the first store (ptxas cannot prove that the
// constraint tile and the source rows do not alias, so a load never moves above an earlier store on its own).
template <int OP, int MODE>
__device__ __forceinline__ void term_columns(const TermCtx& c) {
  int lc = 0;
  for (; lc + 4 <= c.n_cols; lc += 4) {
    const float v0 = column_value<OP>(c, lc), v1 = column_value<OP>(c, lc + 1);
    const float v2 = column_value<OP>(c, lc + 2), v3 = column_value<OP>(c, lc + 3);
    emit<MODE>(c, lc, v0);
    emit<MODE>(c, lc + 1, v1);
    emit<MODE>(c, lc + 2, v2);
    emit<MODE>(c, lc + 3, v3);
  }
  for (; lc < c.n_cols; ++lc) emit<MODE>(c, lc, column_value<OP>(c, lc));
}

// user terms already evaluated to [N, J]: float or bool (CaT.add's cast, :45-47)
template <int MODE>
__device__ __forceinline__ void generic_columns(const TermCtx& c, bool is_u8) {
  for (int lc = 0; lc < c.n_cols; ++lc) {
    const uint32_t id = c.ids[lc];
    emit<MODE>(c, lc, is_u8 ? (lds_u8(c.x0 + id) ? 1.0f : 0.0f) : lds_f32(c.x0 + id * 4));
  }
}

// probability of one column value (constraint_manager.py:64-72)
__device__ __forceinline__ float violation_prob(float c, float rm, float min_p, float span) {
  if (!(c > 0.0f)) return 0.0f;
  float x = __fdiv_rn(c, rm);
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return __fadd_rn(min_p, __fmul_rn(x, span));
}

struct ResetScratch {  // per slot: double sums + count; zero between launches
  double v[CATB200_MAX_TERMS], p[CATB200_MAX_TERMS];
  unsigned long long cnt[CATB200_MAX_TERMS];
  unsigned int ticket;                       // cat_reset_kernel (small grids)
  unsigned int tickets[1 + kTicketGroups];   // fused reset inside cat_apply_kernel (grids of up to N / 32 CTAs)
};

// episode statistics of the envs being reset, from the per-slot accumulators (constraint_manager.py:197-209)
__device__ __forceinline__ void reset_finalize(ResetScratch* sc, int n_slots, float* out) {
  for (int s2 = threadIdx.x; s2 < n_slots; s2 += blockDim.x) {
    const double v = __longlong_as_double(atomicExch((unsigned long long*)&sc->v[s2], 0ull));
    const double p = __longlong_as_double(atomicExch((unsigned long long*)&sc->p[s2], 0ull));
    const unsigned long long c = atomicExch(&sc->cnt[s2], 0ull);
    // empty selection -> mean of nothing = NaN, like torch
    const float mv = (float)(v / (double)c), mp = (float)(p / (double)c);
    out[2 * s2] = __fmul_rn(mv, 100.0f);
    out[2 * s2 + 1] = mp;
  }
}

// Persistent CTAs walk the 32-env tiles blockIdx.x, blockIdx.x + gridDim.x, ...; blockDim.x = 32 * G (G = 1, 2, 4
// or 8 warps share a tile), lane = env.  Two shared-memory images: while tile i is evaluated out of one, the bulk
// async copies of tile i+1 land in the other and the bulk async store of tile i-1's constraint tile drains.
template <int MODE>
__global__ void __launch_bounds__(kEvalMaxWarps * 32)
cat_eval_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                int num_envs, int n_groups, float* __restrict__ running_max, int* __restrict__ rm_init,
                CatWorkspace ws, float* __restrict__ out_rowmajor, const __grid_constant__ ApplyArgs ap) {
  // kEvalStep / kEvalRowMajor: 2 x [source rows | peak table | constraint tile];
  // kEvalFused: 2 x [source rows | peak table], then one constraint tile per tile this CTA walks (they stay here)
  extern __shared__ __align__(128) uint8_t smem_all[];
  __shared__ float s_colmax[CATB200_MAX_COLS];   // column maxima over all tiles of this CTA
  __shared__ float s_tilemax[CATB200_MAX_COLS];  // column maxima of the tile in flight (plain stores in the column loop)
  __shared__ __align__(8) unsigned long long s_bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5, n_threads = blockDim.x;
  const int K = plan.n_cols;
  const int n_tiles = (num_envs + kTile - 1) / kTile;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
  const uint32_t cm_base = launder((uint32_t)__cvta_generic_to_shared(s_tilemax));  // opaque: stays in a register
  const int image_bytes = MODE == kEvalFused ? plan.smem_ctile_off : plan.smem_bytes;
  const int ctile_bytes = K * kTile * 4;

  // lane s describes source s (n_sources <= 16): a source tile can use the bulk copy engine if it is a full tile,
  // contiguous and 16-byte aligned
  const uint8_t* my_ptr = nullptr;
  int my_len = 0, my_stride = 0, my_es = 4, my_off = 0;
  if (lane < plan.n_sources) {
    const catb200_source_t& src = plan.sources[lane];
    my_ptr = static_cast<const uint8_t*>(src.ptr);
    my_len = src.row_len;
    my_stride = src.row_stride;
    my_es = src.dtype == CATB200_U8 ? 1 : 4;
    my_off = src.smem_off;
  }
  auto bulk_ok = [&](int tile) -> bool {
    const uintptr_t g = reinterpret_cast<uintptr_t>(my_ptr) + (size_t)tile * kTile * my_stride * my_es;
    return lane < plan.n_sources && (tile + 1) * kTile <= num_envs && my_stride == my_len && (g & 15) == 0;
  };
  // warp 0: every source lane arrives on buffer b's barrier (with its byte count when it bulk-copies) and starts its copy
  auto issue = [&](int tile, int b) {
    if (lane < plan.n_sources) {
      const uint32_t bar = bar0 + 8 * b;
      if (bulk_ok(tile)) {
        const uint32_t bytes = kTile * my_len * my_es;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
        bulk_copy_g2s((uint32_t)__cvta_generic_to_shared(smem_all) + b * image_bytes + my_off,
                      my_ptr + (size_t)tile * kTile * my_len * my_es, bytes, bar);
      } else {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
      }
    }
  };

  for (int c = threadIdx.x; c < K; c += n_threads) s_colmax[c] = -INFINITY;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar0), "r"(plan.n_sources));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar0 + 8), "r"(plan.n_sources));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (warp == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);

  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
  const int b = it & 1;
  uint8_t* smem = smem_all + (size_t)b * image_bytes;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
  // this tile's [K][32] constraint tile: inside the image, or (fused) the it-th of the tiles kept behind the two images
  const uint8_t* ctile_g = MODE == kEvalFused ? smem_all + 2 * (size_t)image_bytes + (size_t)it * ctile_bytes : smem + plan.smem_ctile_off;
  const float* s_c = reinterpret_cast<const float*>(ctile_g);
  const uint32_t ctile_s = (uint32_t)__cvta_generic_to_shared(ctile_g);
  const int tile0 = tile * kTile;
  const int rows = min(kTile, num_envs - tile0);
  // prefetch the next tile into the other image: every warp left it behind at the barrier that ended the previous
  // iteration; its constraint tile may still be draining (bulk store of iteration it-1), which only reads it
  if (warp == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, b ^ 1);
  // sources that cannot use the bulk engine (strided views, unaligned bases, the ragged last tile): cooperative copy
  const uint32_t coop_mask = ~__ballot_sync(0xffffffffu, bulk_ok(tile)) & ((1u << plan.n_sources) - 1u);
  for (uint32_t m = coop_mask; m; m &= m - 1) {
    const catb200_source_t& src = plan.sources[__ffs(m) - 1];
    const int total = rows * src.row_len;
    if (src.dtype == CATB200_F32) {
      const float* g = static_cast<const float*>(src.ptr);
      float* dst = reinterpret_cast<float*>(smem + src.smem_off);
      for (int f = threadIdx.x; f < total; f += n_threads) {
        const int r = f / src.row_len, e = f - r * src.row_len;
        dst[f] = __ldg(g + (size_t)(tile0 + r) * src.row_stride + e);
      }
    } else {
      const uint8_t* g = static_cast<const uint8_t*>(src.ptr);
      for (int f = threadIdx.x; f < total; f += n_threads) {
        const int r = f / src.row_len, e = f - r * src.row_len;
        smem[src.smem_off + f] = g[(size_t)(tile0 + r) * src.row_stride + e];
      }
    }
  }
  if (coop_mask) __syncthreads();  // (CTA-uniform) cooperative copies visible
  // wait for this image's bulk copies: one lane per warp sleeps on the mbarrier (suspend-time hint), then every
  // lane performs one already-satisfied acquire of its own
  const uint32_t parity = (it >> 1) & 1;
  auto try_wait = [&]() -> uint32_t {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar0 + 8 * b), "r"(parity)
        : "memory");
    return done;
  };
  if (lane == 0) {
    while (!try_wait()) {
    }
  }
  __syncwarp();
  while (!try_wait()) {
  }

  const bool live = lane < rows;
  const uint32_t row = live ? lane : 0;
  const uint32_t ibase = launder(sbase);  // loads of the staged rows must stay behind the wait above
  const uint32_t peaks_w = sbase + plan.smem_peak_off + lane * 4;  // peak_table[0][lane]

  // ---- phase A: contact-force peaks, once per (history tensor, body) pair referenced by any term (dealt to the
  //      warps): max over the history axis of |F[h, body, :]| (constraints.py:102-107,151-158,207-209).  The
  //      correctly rounded square root is monotone, so max_h sqrt(s_h) == sqrt(max_h s_h) bit for bit: one
  //      square root per body.
  for (int p = warp; p < plan.n_peaks; p += n_warps) {
    const catb200_source_t& src = plan.sources[plan.peak_src[p]];
    const uint32_t hstride = src.aux * 12;
    const uint32_t end = ibase + src.smem_off + (row + 1) * (src.row_len * 4);
    float peak = 0.0f;  // sums of squares are >= 0
    for (uint32_t a = ibase + src.smem_off + row * (src.row_len * 4) + plan.peak_body[p] * 12; a < end; a += hstride)
      peak = fmaxf(peak, sumsq3(lds_f32(a), lds_f32(a + 4), lds_f32(a + 8)));
    sts_f32(peaks_w + p * (kTile * 4), __fsqrt_rn(peak));
  }
  __syncthreads();
  const uint32_t peaks = launder(peaks_w);  // loads of the peak table stay behind this barrier

  // ---- phase B: terms, dealt to the warps (term ti -> warp ti mod G).  lane = env, so the op dispatch never
  //      diverges; the op switch sits outside the column loop.
  //      Dealing order: finalize ranked the terms by descending cost; rank k goes to warp k mod G on even rounds
  //      and to the mirrored warp on odd rounds (snake), which balances the warps for any G.
  for (int k0 = 0; k0 < plan.n_terms; k0 += 2 * n_warps) {
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int k = half ? k0 + 2 * n_warps - 1 - warp : k0 + warp;
    if (k >= plan.n_terms) continue;
    const catb200_term_t& t2 = plan.terms[plan.terms[k].reserved2];
    const int op = t2.op;
    const catb200_source_t& s0 = plan.sources[t2.src0];
    TermCtx c;
    c.ids = t2.ids;
    c.n_ids = t2.n_ids;
    c.n_cols = t2.n_cols;
    c.p0 = t2.p0;
    c.p1 = t2.p1;
    c.p2 = t2.p2;
    c.live = live;
    c.lane0 = lane == 0;
    c.peaks = peaks;
    c.out = ctile_s + (t2.col_offset * kTile + lane) * 4;
    c.colmax = cm_base + t2.col_offset * 4;
    const bool u0 = s0.dtype == CATB200_U8;
    c.x0 = ibase + s0.smem_off + row * (s0.row_len * (u0 ? 1 : 4));
    c.x1 = c.x0;
    c.u1 = 0;
    if (t2.src1 != 0xff) {
      const catb200_source_t& s1 = plan.sources[t2.src1];
      c.u1 = s1.dtype == CATB200_U8;
      c.x1 = ibase + s1.smem_off + row * (s1.row_len * (c.u1 ? 1 : 4));
    }
    c.gate = 1.0f;  // command-dependent factor shared by all columns of the term
    if (t2.src2 != 0xff) {
      const catb200_source_t& sc = plan.sources[t2.src2];
      const uint32_t a = ibase + sc.smem_off + row * (sc.row_len * 4);
      if (op == CATB200_OP_ABSDIFF_MINUS_GATE_Y) {
        c.gate = fabsf(lds_f32(a + 4)) < c.p1 ? 1.0f : 0.0f;  // constraints.py:46-53
      } else {
        const float cn = norm3(lds_f32(a), lds_f32(a + 4), lds_f32(a + 8));
        c.gate = op == CATB200_OP_ABS_MINUS_GATE_STILL ? (cn < c.p1 ? 1.0f : 0.0f) : (cn > c.p1 ? 1.0f : 0.0f);
      }
    }
    switch (op) {
      case CATB200_OP_GENERIC: generic_columns<MODE>(c, u0); break;
      case CATB200_OP_ABS_MINUS: term_columns<CATB200_OP_ABS_MINUS, MODE>(c); break;
      case CATB200_OP_ABSDIFF_MINUS: term_columns<CATB200_OP_ABSDIFF_MINUS, MODE>(c); break;
      case CATB200_OP_ABSDIFF_MINUS_GATE_Y: term_columns<CATB200_OP_ABSDIFF_MINUS_GATE_Y, MODE>(c); break;
      case CATB200_OP_ACTION_RATE: term_columns<CATB200_OP_ACTION_RATE, MODE>(c); break;
      case CATB200_OP_COMPONENT_GT: term_columns<CATB200_OP_COMPONENT_GT, MODE>(c); break;
      case CATB200_OP_CONTACT_ANY: term_columns<CATB200_OP_CONTACT_ANY, MODE>(c); break;
      case CATB200_OP_NORM2_MINUS: term_columns<CATB200_OP_NORM2_MINUS, MODE>(c); break;
      case CATB200_OP_AIR_TIME: term_columns<CATB200_OP_AIR_TIME, MODE>(c); break;
      case CATB200_OP_N_CONTACT: term_columns<CATB200_OP_N_CONTACT, MODE>(c); break;
      case CATB200_OP_FORCE_PEAK_MINUS: term_columns<CATB200_OP_FORCE_PEAK_MINUS, MODE>(c); break;
      case CATB200_OP_LIMIT_MINUS: term_columns<CATB200_OP_LIMIT_MINUS, MODE>(c); break;
      case CATB200_OP_ABS_MINUS_GATE_STILL: term_columns<CATB200_OP_ABS_MINUS_GATE_STILL, MODE>(c); break;
      default: break;
    }
  }
  }

  if (MODE == kEvalRowMajor) {
    // debug / stand-alone path: transpose the tile into the row-major [N][K] matrix (coalesced global writes)
    __syncthreads();
    for (int f = threadIdx.x; f < rows * K; f += n_threads) {
      const int r = f / K, col = f - r * K;
      out_rowmajor[(size_t)(tile0 + r) * K + col] = s_c[col * kTile + r];
    }
    __syncthreads();
    continue;
  }

  // ---- phase C: the tile leaves as one bulk async store (generic-proxy writes fenced for the async proxy first).
  //      Before that, the store issued two iterations ago out of this same image must have finished reading it --
  //      thread 0 waited for that right after issuing the previous one (at most 1 group pending).
  //      Fused: the tile stays where it is.
  if (MODE != kEvalFused) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();  // also: every warp is done with this image's sources before the next iteration's prefetch
  if (MODE != kEvalFused && threadIdx.x == 0) {
    float* dst = ws.c_t + (size_t)tile * K * kTile;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst),
                 "r"(ctile_s), "r"((uint32_t)ctile_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");  // the other image's constraint tile is free again
  }
  // fold this tile's column maxima into the CTA's (the next tile overwrites s_tilemax only after its peak barrier)
  for (int c = threadIdx.x; c < K; c += n_threads) s_colmax[c] = fmaxf(s_colmax[c], s_tilemax[c]);
  }  // tiles
  if (MODE == kEvalRowMajor) return;

  // ---- the CTA folds its column maxima into the scratch rows, then the last CTA of the grid folds everything
  //      into the Polyak running max (constraint_manager.py:55-61)
  __syncthreads();
  // n_groups scratch rows (power of two, chosen by the host: a few hundred CTAs per row keep the atomic queues short)
  uint32_t* grow = ws.colmax + (blockIdx.x & (n_groups - 1)) * CATB200_MAX_COLS;
  for (int c = threadIdx.x; c < K; c += n_threads) {
    if ((int)blockIdx.x >= n_tiles) break;
    const uint32_t key = float_to_ordered(s_colmax[c]);
    // most CTAs cannot raise the maximum any more: a plain (L2) read first keeps the atomic units idle
    if (key > __ldcg(grow + c)) atomicMax(grow + c, key);
  }
  if (MODE != kEvalFused && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");  // smem read out before exit
  // fused: the generation of the grid barrier's flag, read BEFORE this CTA counts as arrived (it cannot change earlier)
  volatile unsigned int* gen_flag = ws.ticket + 48;
  unsigned int gen0 = 0;
  if (MODE == kEvalFused && threadIdx.x == 0) gen0 = *gen_flag;
  const bool last_cta = last_block_ticket_grouped(ws.ticket, gridDim.x);
  if (last_cta) {
    for (int col = threadIdx.x; col < K; col += n_threads) {
      uint32_t key = 0u;
      int gi = 0;
      for (; gi + 4 <= n_groups; gi += 4) {  // 4 independent exchanges in flight
        const uint32_t k0 = atomicExch(&ws.colmax[gi * CATB200_MAX_COLS + col], 0u);
        const uint32_t k1 = atomicExch(&ws.colmax[(gi + 1) * CATB200_MAX_COLS + col], 0u);
        const uint32_t k2 = atomicExch(&ws.colmax[(gi + 2) * CATB200_MAX_COLS + col], 0u);
        const uint32_t k3 = atomicExch(&ws.colmax[(gi + 3) * CATB200_MAX_COLS + col], 0u);
        key = max(max(key, max(k0, k1)), max(k2, k3));
      }
      for (; gi < n_groups; ++gi) key = max(key, atomicExch(&ws.colmax[gi * CATB200_MAX_COLS + col], 0u));
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
  if (MODE != kEvalFused) return;

  // ---- single launch: grid barrier (every CTA is resident: the host sized the grid with the occupancy calculator),
  //      then the apply phase straight from the constraint tiles in shared memory -- no C_T round trip, no second launch.
  //      The last CTA has just written running_max; it publishes a new generation of the flag, the others wait for it.
  __syncthreads();
  if (threadIdx.x == 0) {
    if (last_cta) {
      __threadfence();
      atomicExch(const_cast<unsigned int*>(gen_flag), gen0 + 1u);
    } else {
      const long long t0 = clock64();
      while (*gen_flag == gen0) {
        if (clock64() - t0 > 4000000000ll) __trap();  // a CTA that never became resident: fail the launch, do not hang
      }
      __threadfence();
    }
  }
  __syncthreads();
  __shared__ float s_part[kEvalMaxWarps][kTile];
  for (int c = threadIdx.x; c < K; c += n_threads) s_tilemax[c] = __ldcg(running_max + c);  // reuse: the new running max
  __syncthreads();
  const float* s_rm = s_tilemax;
  const bool fused_reset = ap.episode_length != nullptr;
  it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const float* ct = reinterpret_cast<const float*>(smem_all + 2 * (size_t)image_bytes + (size_t)it * ctile_bytes) + lane;
    const int i = tile * kTile + lane;
    const bool live = i < num_envs;
    float overall = -INFINITY;
    const bool resetting = fused_reset && live && ap.reset_buf[i] != 0;
    const unsigned reset_lanes = fused_reset ? __ballot_sync(0xffffffffu, resetting) : 0u;
    const float ep_len = resetting ? (float)ap.episode_length[i] : 1.0f;  // int64 -> float like torch's float / long promotion
    for (int slot = warp; slot < plan.n_slots; slot += n_warps) {
      float es2 = 0.0f, mv2 = 0.0f;
      if (live) {
        const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
        const float span = prm.span_dev ? __ldg(prm.span_dev + slot) : prm.span[slot];
        const size_t k = (size_t)slot * num_envs + i;
        const float es = ap.episode_sums[k], mv = ap.mean_values[k];
        float tmax = -INFINITY;
        for (int col = c0; col < c1; ++col) tmax = fmaxf(tmax, violation_prob(ct[col * kTile], s_rm[col], prm.min_p, span));
        es2 = __fadd_rn(es, tmax > 0.0f ? 1.0f : 0.0f);  // :226
        mv2 = __fadd_rn(mv, tmax);                         // :227
        ap.episode_sums[k] = resetting ? 0.0f : es2;
        ap.mean_values[k] = resetting ? 0.0f : mv2;
        overall = fmaxf(overall, tmax);
      }
      if (reset_lanes) {  // warp-uniform: some env of this tile resets (every lane of the warp takes part in the sums)
        const double av = warp_sum(resetting ? (double)__fdiv_rn(es2, ep_len) : 0.0);
        const double apv = warp_sum(resetting ? (double)__fdiv_rn(mv2, ep_len) : 0.0);
        if (lane == 0) {
          atomicAdd(&ap.rsc->v[slot], av);
          atomicAdd(&ap.rsc->p[slot], apv);
          atomicAdd(&ap.rsc->cnt[slot], (unsigned long long)__popc(reset_lanes));
        }
      }
    }
    s_part[warp][lane] = overall;
    __syncthreads();
    if (warp == 0 && live) {
      for (int w = 1; w < n_warps; ++w) overall = fmaxf(overall, s_part[w][lane]);
      ap.cstr_prob[i] = overall;
      if (ap.raw_reward != nullptr) {
        // cat_env.py:102-107: reward = clip(reward * (1 - p), min=0); dones = p; :121 dones[reset] = 1
        ap.reward_out[i] = fmaxf(__fmul_rn(ap.raw_reward[i], __fsub_rn(1.0f, overall)), 0.0f);
        ap.dones_out[i] = (ap.reset_buf != nullptr && ap.reset_buf[i]) ? 1.0f : overall;
      }
    }
    __syncthreads();  // s_part is rewritten by the next tile
  }
  if (fused_reset && last_block_ticket_grouped(ap.rsc->tickets, gridDim.x)) reset_finalize(ap.rsc, plan.n_slots, ap.reset_out);
}

// One CTA per 32 envs (lane = env); the warps split the statistics slots (terms) so that the dependent
// load chains are 4x shorter and 4x more loads are in flight than with one thread per env.
// With `episode_length` (fused reset, the order of CaTEnv.step: compute() at cat_env.py:100, then
// _reset_idx -> ConstraintManager.reset at :181): envs whose reset_buf is set contribute their just-updated
// statistics / episode length to the per-term means (constraint_manager.py:197-209) and are zeroed, and the
// last CTA writes the 2 * n_slots means -- ConstraintManager.reset(reset ids) without a second launch.
__global__ void __launch_bounds__(kApplyMaxWarps * 32)
cat_apply_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ c_t,
                 float* __restrict__ episode_sums, float* __restrict__ mean_values,
                 float* __restrict__ cstr_prob, const float* __restrict__ raw_reward,
                 const uint8_t* __restrict__ reset_buf, float* __restrict__ reward_out,
                 float* __restrict__ dones_out, const int64_t* __restrict__ episode_length,
                 float* __restrict__ reset_out, ResetScratch* __restrict__ rsc) {
  __shared__ float s_rm[CATB200_MAX_COLS];
  __shared__ float s_part[kApplyMaxWarps][kTile];
  const int n_warps = blockDim.x >> 5;
  for (int c = threadIdx.x; c < plan.n_cols; c += blockDim.x) s_rm[c] = running_max[c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kTile + lane;
  const bool live = i < num_envs;
  const float* ct = c_t + (size_t)blockIdx.x * plan.n_cols * kTile + lane;  // this env's column of the [K][32] tile
  float overall = -INFINITY;
  const bool fused_reset = episode_length != nullptr;
  const bool resetting = fused_reset && live && reset_buf[i] != 0;
  const unsigned reset_lanes = fused_reset ? __ballot_sync(0xffffffffu, resetting) : 0u;
  const float ep_len = resetting ? (float)episode_length[i] : 1.0f;  // int64 -> float like torch's float / long promotion
  for (int slot = warp; slot < plan.n_slots; slot += n_warps) {
    float es2 = 0.0f, mv2 = 0.0f;
    if (live) {
      const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
      const float span = prm.span_dev ? __ldg(prm.span_dev + slot) : prm.span[slot];
      const size_t k = (size_t)slot * num_envs + i;
      const float es = episode_sums[k], mv = mean_values[k];  // requested before the columns: all in flight together
      float tmax = -INFINITY;
      int col = c0;
      for (; col + 4 <= c1; col += 4) {
        const float v0 = __ldcs(ct + col * kTile);
        const float v1 = __ldcs(ct + (col + 1) * kTile);
        const float v2 = __ldcs(ct + (col + 2) * kTile);
        const float v3 = __ldcs(ct + (col + 3) * kTile);
        tmax = fmaxf(tmax, fmaxf(fmaxf(violation_prob(v0, s_rm[col], prm.min_p, span), violation_prob(v1, s_rm[col + 1], prm.min_p, span)),
                                 fmaxf(violation_prob(v2, s_rm[col + 2], prm.min_p, span), violation_prob(v3, s_rm[col + 3], prm.min_p, span))));
      }
      for (; col < c1; ++col) tmax = fmaxf(tmax, violation_prob(__ldcs(ct + col * kTile), s_rm[col], prm.min_p, span));
      es2 = __fadd_rn(es, tmax > 0.0f ? 1.0f : 0.0f);  // :226
      mv2 = __fadd_rn(mv, tmax);                         // :227
      episode_sums[k] = resetting ? 0.0f : es2;
      mean_values[k] = resetting ? 0.0f : mv2;
      overall = fmaxf(overall, tmax);
    }
    if (reset_lanes) {  // warp-uniform: some env of this tile resets (every lane of the warp takes part in the sums)
      const double av = warp_sum(resetting ? (double)__fdiv_rn(es2, ep_len) : 0.0);
      const double ap = warp_sum(resetting ? (double)__fdiv_rn(mv2, ep_len) : 0.0);
      if (lane == 0) {
        atomicAdd(&rsc->v[slot], av);
        atomicAdd(&rsc->p[slot], ap);
        atomicAdd(&rsc->cnt[slot], (unsigned long long)__popc(reset_lanes));
      }
    }
  }
  s_part[warp][lane] = overall;
  __syncthreads();
  if (warp == 0 && live) {
    for (int w = 1; w < n_warps; ++w) overall = fmaxf(overall, s_part[w][lane]);
    cstr_prob[i] = overall;
    if (raw_reward != nullptr) {
      // cat_env.py:102-107: reward = clip(reward * (1 - p), min=0); dones = p; :121 dones[reset] = 1
      reward_out[i] = fmaxf(__fmul_rn(raw_reward[i], __fsub_rn(1.0f, overall)), 0.0f);
      dones_out[i] = (reset_buf != nullptr && reset_buf[i]) ? 1.0f : overall;
    }
  }
  if (fused_reset && last_block_ticket_grouped(rsc->tickets, gridDim.x)) reset_finalize(rsc, plan.n_slots, reset_out);
}

__global__ void __launch_bounds__(kApplyThreads)
cat_probs_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ raw,
                 float* __restrict__ probs_out) {
  const int i = blockIdx.x * kApplyThreads + threadIdx.x;
  if (i >= num_envs) return;
  for (int slot = 0; slot < plan.n_slots; ++slot) {
    const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
    for (int col = c0; col < c1; ++col) {
      const size_t e = (size_t)i * plan.n_cols + col;
      probs_out[e] = violation_prob(raw[e], running_max[col], prm.min_p, prm.span_dev ? __ldg(prm.span_dev + slot) : prm.span[slot]);
    }
  }
}

// ---- ConstraintManager.reset (constraint_manager.py:190-211) -----------------------------------------
// grid = n_slots CTAs; each reduces its statistics row over the selected envs in double precision
// (torch's own fp32 reduction order is implementation defined; parity tolerance 1e-5 relative) and
// then zeroes the selected entries.
constexpr int kResetThreads = 256;
constexpr int kResetChunk = 1024;  // selected envs (or mask entries) per CTA


// grid = (n_slots, chunks): every CTA reduces one statistics row over one chunk of the selection in double
// precision (torch's own fp32 reduction order is implementation defined; parity tolerance 1e-5 relative),
// zeroes the selected entries and adds its partial sums to the slot's accumulators; the last CTA turns the
// accumulators into the two means per slot.
__global__ void __launch_bounds__(kResetThreads)
cat_reset_kernel(const int64_t* __restrict__ env_ids, int n_ids, const uint8_t* __restrict__ mask,
                 const int64_t* __restrict__ episode_length, int num_envs, int n_slots, float* __restrict__ episode_sums,
                 float* __restrict__ mean_values, float* __restrict__ out, ResetScratch* __restrict__ sc) {
  const int slot = blockIdx.x;
  float* sums = episode_sums + (size_t)slot * num_envs;
  float* means = mean_values + (size_t)slot * num_envs;
  double acc_v = 0.0, acc_p = 0.0;
  unsigned long long cnt = 0;
  const int total = env_ids ? n_ids : num_envs;
  const int begin = blockIdx.y * kResetChunk, end = min(total, begin + kResetChunk);
  for (int k = begin + threadIdx.x; k < end; k += kResetThreads) {
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
  acc_v = warp_sum(acc_v);
  acc_p = warp_sum(acc_p);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicAdd(&sc->v[slot], acc_v);
    atomicAdd(&sc->p[slot], acc_p);
    atomicAdd(&sc->cnt[slot], cnt);
  }
  if (last_block_ticket(&sc->ticket, gridDim.x * gridDim.y)) reset_finalize(sc, n_slots, out);
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
  int off = 0;  // bytes within a tile's shared-memory image
  for (int s = 0; s < plan->n_sources; ++s) {
    catb200_source_t& src = plan->sources[s];
    if (src.row_len <= 0 || src.row_len > 4096 || src.row_stride < src.row_len) return CATB200_ERR_INVALID_ARGUMENT;
    if (src.dtype != CATB200_F32 && src.dtype != CATB200_U8) return CATB200_ERR_UNSUPPORTED;
    if (src.aux < 0 || (src.aux > 0 && src.row_len % (3 * src.aux) != 0)) return CATB200_ERR_INVALID_ARGUMENT;
    src.smem_off = off;  // 16-byte aligned: destination of a bulk async copy
    off += (kTile * src.row_len * (src.dtype == CATB200_U8 ? 1 : 4) + 15) & ~15;
    src.magic = 0;
  }
  plan->n_peaks = 0;
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
    if (contact_op) {
      // body ids -> slots of the shared peak table (one entry per distinct (history tensor, body) pair)
      if (term.reserved != 0) return CATB200_ERR_INVALID_ARGUMENT;  // plan was already finalized
      for (int k = 0; k < term.n_ids; ++k) {
        int slot = -1;
        for (int p = 0; p < plan->n_peaks; ++p)
          if (plan->peak_src[p] == term.src0 && plan->peak_body[p] == term.ids[k]) slot = p;
        if (slot < 0) {
          if (plan->n_peaks >= CATB200_MAX_PEAKS) return CATB200_ERR_UNSUPPORTED;
          slot = plan->n_peaks++;
          plan->peak_src[slot] = term.src0;
          plan->peak_body[slot] = term.ids[k];
        }
        term.ids[k] = (uint8_t)slot;
      }
      term.reserved = 1;
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
  // evaluation order: rank of every term by descending cost (prologue ~ 6 columns' worth; IEEE division doubles a
  // column).  The kernel deals the ranks to its warps in snake order, which balances them for any warp count.
  for (int t = 0; t < plan->n_terms; ++t) {
    auto cost = [&](int i) {
      const catb200_term_t& x = plan->terms[i];
      return 6 + (int)x.n_cols * (x.op == CATB200_OP_ACTION_RATE ? 2 : 1);
    };
    int rank = 0;
    for (int u = 0; u < plan->n_terms; ++u)
      if (cost(u) > cost(t) || (cost(u) == cost(t) && u < t)) ++rank;
    plan->terms[rank].reserved2 = (uint16_t)t;  // entry k holds the index of the term with rank k
  }
  plan->slot_col_begin[slots] = (uint16_t)col;
  plan->n_cols = col;
  plan->n_slots = slots;
  plan->smem_peak_off = off;
  off += plan->n_peaks * kTile * 4;
  off = (off + 127) & ~127;
  plan->smem_ctile_off = off;  // [n_cols][32] fp32 constraint tile, source of the bulk async store
  off += plan->n_cols * kTile * 4;
  plan->smem_bytes = off;
  if (2 * plan->smem_bytes > 200 * 1024) return CATB200_ERR_UNSUPPORTED;  // two tile images must fit a CTA
  return CATB200_OK;
}

size_t catb200_cat_workspace_bytes(int32_t num_envs, int32_t n_cols) {
  if (num_envs < 0 || n_cols < 0) return 0;
  const size_t n_tiles = ((size_t)num_envs + kTile - 1) / kTile;  // C_T is tile-major: whole tiles
  return 256 + align256(sizeof(uint32_t) * CATB200_MAX_COLS * kMaxGroups) + align256(sizeof(float) * n_tiles * kTile * n_cols);
}

// warps sharing one 32-env tile (terms are dealt to them).  CATB200_EVAL_WARPS overrides.
static int eval_warps(int n_terms) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = std::getenv("CATB200_EVAL_WARPS");
    const int v = e ? std::atoi(e) : 0;
    forced = (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) ? v : 0;
  }
  if (forced) return forced;
  int g = 1;  // a term is evaluated by one warp: no point in more warps than terms (measured: 16 >= 8 > 4 at every size)
  while (g < kEvalMaxWarps && g < n_terms) g <<= 1;
  return g;
}

// apply kernel: one warp per statistics slot up to 16 (CATB200_APPLY_WARPS overrides)
static int apply_warps(int n_slots, int n_tiles) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = std::getenv("CATB200_APPLY_WARPS");
    const int v = e ? std::atoi(e) : 0;
    forced = (v >= 1 && v <= kApplyMaxWarps) ? v : 0;
  }
  if (forced) return forced;
  // few tiles: the launch is one latency chain per CTA, a warp per slot shortens it (4.6 vs 6.4 us at 4096 envs);
  // many tiles: 8 warps keep more CTAs resident and move more bytes (248 vs 314 us at 1 M envs)
  return n_tiles <= 2 * kNumSMs ? max(1, min(kApplyMaxWarps, n_slots)) : max(1, min(8, n_slots));
}

static int launch_eval(const catb200_plan_t* plan, const catb200_cat_params_t* prm, int num_envs, float* running_max,
                       int* rm_init, CatWorkspace ws, float* out_rowmajor, int mode, cudaStream_t stream) {
  const size_t smem = 2 * (size_t)plan->smem_bytes;  // two images of [source rows | peak table | constraint tile]
  const int n_tiles = (num_envs + kTile - 1) / kTile;
  const int threads = 32 * eval_warps(plan->n_terms);
  // persistent grid: as many CTAs as stay resident (shared memory, 2048 threads per SM), never more than tiles
  const int per_sm = (int)max((size_t)1, min((size_t)(2048 / threads), (size_t)(227 * 1024) / (smem + 3 * 1024)));
  const int grid = min(n_tiles, kNumSMs * per_sm);
  int n_groups = 1;  // scratch rows for the cross-CTA column maxima (power of two, <= kMaxGroups)
  while (n_groups < kMaxGroups && n_groups * 256 <= grid) n_groups <<= 1;
  if (mode == kEvalStep) {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalStep>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalStep><<<grid, threads, smem, stream>>>(*plan, *prm, num_envs, n_groups, running_max, rm_init, ws, nullptr, ApplyArgs{});
  } else {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalRowMajor>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalRowMajor><<<grid, threads, smem, stream>>>(*plan, *prm, num_envs, n_groups, nullptr, nullptr, ws, out_rowmajor, ApplyArgs{});
  }
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

// Single-launch step (cat_eval_kernel<kEvalFused>): evaluation, grid barrier, apply phase out of shared memory.  Possible
// while every CTA of the grid is resident at once AND the constraint tiles of all the tiles a CTA walks fit behind its two
// staging images: up to ~70 k envs with the Solo12 plan (14 tiles of 10 KiB per CTA).  Returns the grid (0: use the
// two-launch path) and the dynamic shared memory.  CATB200_CAT_FUSED=0 disables it.
static int fused_geometry(const catb200_plan_t* plan, int num_envs, int threads, size_t* smem_out) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = std::getenv("CATB200_CAT_FUSED");
    enabled = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;
  }
  if (!enabled) return 0;
  const int n_tiles = (num_envs + kTile - 1) / kTile;
  const size_t image = (size_t)plan->smem_ctile_off, ctile = (size_t)plan->n_cols * kTile * 4;
  const size_t budget = 227 * 1024 - 8 * 1024;  // static shared memory (column maxima, partial maxima) + reserve
  for (int per_sm = min(4, 2048 / threads); per_sm >= 1; --per_sm) {
    const int grid = min(n_tiles, kNumSMs * per_sm);
    const int per_cta = (n_tiles + grid - 1) / grid;
    // Measured on the B200 (profiles/README.md, round 2): with one tile per CTA the single launch matches eval + apply
    // (19.3 vs 18.9 us inside the step graph at 4096 envs, one launch and the C_T round trip fewer); with several tiles per
    // CTA the apply phase runs out of 16 warps per SM and loses (95 vs 56 us at 65536 envs).  CATB200_CAT_FUSED=2 forces it.
    if (per_cta > 1 && enabled != 2) continue;
    const size_t smem = 2 * image + (size_t)per_cta * ctile;
    if ((smem + 1024) * per_sm > budget) continue;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(cat_eval_kernel<kEvalFused>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      continue;
    int resident = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cat_eval_kernel<kEvalFused>, threads, smem) != cudaSuccess) continue;
    if (resident * kNumSMs < grid) continue;
    *smem_out = smem;
    return grid;
  }
  return 0;
}

static int cat_step_impl(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                         float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                         float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf, float* reward_out,
                         float* dones_out, void* workspace, size_t workspace_bytes, const int64_t* episode_length,
                         float* reset_out, void* reset_workspace, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !rm_init || !episode_sums || !mean_values || !cstr_prob ||
      !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_cols <= 0 || plan->smem_bytes <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  if (raw_reward && (!reward_out || !dones_out)) return CATB200_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < catb200_cat_workspace_bytes(num_envs, plan->n_cols)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  CatWorkspace ws = carve(workspace, num_envs);
  {
    const int threads = 32 * eval_warps(plan->n_terms);
    size_t smem = 0;
    const int fgrid = fused_geometry(plan, num_envs, threads, &smem);
    if (fgrid > 0) {
      int n_groups = 1;
      while (n_groups < kMaxGroups && n_groups * 256 <= fgrid) n_groups <<= 1;
      ApplyArgs ap = {episode_sums, mean_values, cstr_prob, raw_reward, reset_buf, reward_out, dones_out, episode_length,
                      reset_out, static_cast<ResetScratch*>(reset_workspace)};
      cat_eval_kernel<kEvalFused><<<fgrid, threads, smem, st>>>(*plan, *params, num_envs, n_groups, running_max, rm_init, ws, nullptr, ap);
      CATB200_LAUNCH_CHECK();
      return CATB200_OK;
    }
  }
  int rc = launch_eval(plan, params, num_envs, running_max, rm_init, ws, nullptr, kEvalStep, st);
  if (rc != CATB200_OK) return rc;
  const int grid = (num_envs + kTile - 1) / kTile;
  cat_apply_kernel<<<grid, 32 * apply_warps(plan->n_slots, grid), 0, st>>>(
      *plan, *params, num_envs, running_max, ws.c_t, episode_sums, mean_values, cstr_prob, raw_reward, reset_buf,
      reward_out, dones_out, episode_length, reset_out, static_cast<ResetScratch*>(reset_workspace));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_step(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                     float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                     float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf, float* reward_out,
                     float* dones_out, void* workspace, size_t workspace_bytes, void* stream) {
  return cat_step_impl(plan, params, num_envs, running_max, rm_init, episode_sums, mean_values, cstr_prob, raw_reward,
                       reset_buf, reward_out, dones_out, workspace, workspace_bytes, nullptr, nullptr, nullptr, stream);
}

int catb200_cat_step_reset(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                           float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                           float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf, float* reward_out,
                           float* dones_out, void* workspace, size_t workspace_bytes, const int64_t* episode_length,
                           float* reset_out, void* reset_workspace, size_t reset_workspace_bytes, void* stream) {
  if (!reset_buf || !episode_length || !reset_out || !reset_workspace) return CATB200_ERR_INVALID_ARGUMENT;
  if (reset_workspace_bytes < sizeof(ResetScratch)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  return cat_step_impl(plan, params, num_envs, running_max, rm_init, episode_sums, mean_values, cstr_prob, raw_reward,
                       reset_buf, reward_out, dones_out, workspace, workspace_bytes, episode_length, reset_out,
                       reset_workspace, stream);
}

int catb200_cat_eval_terms(const catb200_plan_t* plan, int32_t num_envs, float* out, void* stream) {
  if (!plan || num_envs <= 0 || !out || plan->n_cols <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  catb200_cat_params_t dummy = {};
  CatWorkspace ws = {};
  return launch_eval(plan, &dummy, num_envs, nullptr, nullptr, ws, out, kEvalRowMajor, as_stream(stream));
}

int catb200_cat_probs(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                      const float* running_max, float* probs_out, const float* raw, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !probs_out || !raw) return CATB200_ERR_INVALID_ARGUMENT;
  const int grid = (num_envs + kApplyThreads - 1) / kApplyThreads;
  cat_probs_kernel<<<grid, kApplyThreads, 0, as_stream(stream)>>>(*plan, *params, num_envs, running_max, raw, probs_out);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

size_t catb200_cat_reset_workspace_bytes(void) { return sizeof(ResetScratch); }

int catb200_cat_reset_stats(const int64_t* env_ids, int32_t n_ids, const uint8_t* mask, const int64_t* episode_length,
                            int32_t num_envs, int32_t n_slots, float* episode_sums, float* mean_values, float* out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!episode_length || num_envs <= 0 || n_slots <= 0 || n_slots > CATB200_MAX_TERMS || !episode_sums || !mean_values ||
      !out || !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (env_ids && n_ids < 0) return CATB200_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < sizeof(ResetScratch)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  const int total = env_ids ? n_ids : num_envs;
  const int chunks = max(1, (total + kResetChunk - 1) / kResetChunk);
  cat_reset_kernel<<<dim3(n_slots, chunks), kResetThreads, 0, as_stream(stream)>>>(
      env_ids, n_ids, mask, episode_length, num_envs, n_slots, episode_sums, mean_values, out,
      static_cast<ResetScratch*>(workspace));
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

}  // extern "C"

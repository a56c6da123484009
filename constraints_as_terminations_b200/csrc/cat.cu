// Per-step CaT path on sm_100a: constraint terms -> termination probabilities.
//
// Replaces, in two launches, the ~250-300 eager kernels + 13 host syncs of the reference's
// ConstraintManager.compute() (U/cat/constraint_manager.py:213-229 driving constraints.py:23-235 and
// CaT.add/get_probs :39-82) and the reward/dones lines of CaTEnv.step (U/cat/cat_env.py:102-121).
//
// Data flow (N envs, K constraint columns, S statistics slots):
//   cat_eval_kernel : persistent CTAs walk 32-env tiles.  The tile's rows of every source tensor are copied
//                     verbatim into shared memory by the bulk async-copy engine (cp.async.bulk + mbarrier, one
//                     copy per source, no per-element staging instructions), double buffered: tile i+1 lands
//                     while tile i is evaluated.  Contact-force peaks are computed once per (history tensor,
//                     body) pair into a shared table; then warp w evaluates columns w, w+W, ... of every term
//                     (lane = env, warp-uniform op dispatch hoisted out of the column loop, term-level gates
//                     computed once), stores the raw constraint column-major into the workspace (C_T[K][N],
//                     coalesced) and keeps the column maxima in shared memory (redux.sync per column).  At the
//                     end each CTA issues one atomicMax per column into one of up to 64 scratch rows and the
//                     last CTA (two-level ticket) applies the clamp + Polyak update to running_max[K] (:55-61).
//   cat_apply_kernel: one CTA per 32 envs (lane = env), 8 warps split the statistics slots.  Reads the K
//                     constraint values back (coalesced, L2 hits), maps violations to probabilities (:64-72),
//                     takes the per-term and overall row max (:82,:225), updates the two per-term episode
//                     statistics (:226-227) and writes cstr_prob plus, optionally, the scaled reward / float dones.
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

constexpr int kTile = 32;          // envs per CTA in the eval kernel (one per lane)
constexpr int kEvalWarps = 8;      // warps per CTA of the apply kernel (and of the eval kernel at large N)
constexpr int kEvalThreads = kEvalWarps * 32;
constexpr int kEvalWarpsSmallN = 16;  // eval kernel at small N: shorter per-tile critical path (latency bound there)
constexpr int kEvalThreadsMax = kEvalWarpsSmallN * 32;
constexpr int kApplyThreads = 64;

// ---- staged-source accessors ----------------------------------------------------------------------
// A tile's rows of every source are copied verbatim (row pitch = row_len) into shared memory by the bulk
// async-copy engine (cp.async.bulk + mbarrier, i.e. TMA's 1-D path): zero per-element staging instructions.
struct SrcView {
  const uint8_t* base;  // shared-memory bytes of this source's tile
  int row_len;
  int is_u8;
  __device__ __forceinline__ float at(int row, int e) const {
    return is_u8 ? (base[row * row_len + e] ? 1.0f : 0.0f) : reinterpret_cast<const float*>(base)[row * row_len + e];
  }
  // sources the library ops read as fp32 by construction (state tensors); bool / u8 only reach `at`
  __device__ __forceinline__ float f32(int row, int e) const { return reinterpret_cast<const float*>(base)[row * row_len + e]; }
};

__device__ __forceinline__ SrcView view_of(const catb200_plan_t& plan, const uint8_t* smem, int s) {
  const catb200_source_t& src = plan.sources[s];
  return SrcView{smem + src.smem_off, src.row_len, src.dtype == CATB200_U8};
}

// sqrt(x^2 + y^2 + z^2) the way torch.norm reduces a short contiguous dim on CPU and CUDA:
// sequential fused multiply-adds from a zero accumulator, then a correctly rounded sqrt.
__device__ __forceinline__ float norm3(float x, float y, float z) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  acc = __fmaf_rn(z, z, acc);
  return __fsqrt_rn(acc);
}
__device__ __forceinline__ float norm2(float x, float y) {
  float acc = __fmul_rn(x, x);
  acc = __fmaf_rn(y, y, acc);
  return __fsqrt_rn(acc);
}

// max over the history axis of |F[h, body, :]| for one body (constraints.py:102-107,151-158,207-209)
__device__ __forceinline__ float force_peak(const SrcView& v, int bodies, int row, int body) {
  const int H = v.row_len / (3 * bodies);
  float peak = -INFINITY;
  for (int h = 0; h < H; ++h) {
    const int e = (h * bodies + body) * 3;
    peak = fmaxf(peak, norm3(v.f32(row, e), v.f32(row, e + 1), v.f32(row, e + 2)));
  }
  return peak;
}

// Same-address atomics serialise in L2 (tens of ns each): with one scratch word per column, 32 768 CTAs
// (1 M envs) would queue 32 768 deep.  CTAs therefore fold their column maxima into one of kMaxGroups
// scratch rows (blockIdx % groups) and the last CTA reduces the rows.
constexpr int kMaxGroups = 64;

struct CatWorkspace {
  // layout inside the caller's workspace (all 256-byte aligned)
  unsigned int* ticket;   // 1 word (padded)
  uint32_t* colmax;       // [kMaxGroups][CATB200_MAX_COLS] ordered-float column maxima, 0 between launches
  float* c_t;             // [K][N] raw constraints, column-major
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

enum EvalMode { kEvalStep = 0, kEvalRowMajor = 1 };

// ---- per-term column evaluation ----------------------------------------------------------------------
struct TermCtx {
  const uint8_t* ids;  // joint / body / peak-slot ids of the term (kernel-parameter memory)
  int n_ids, first, n_cols, col0, row, lane, n_warps;
  float p0, p1, p2, gate;
  const float* peaks;
  SrcView v0, v1;
};

template <int MODE>
struct ColumnSink {  // where a column value of this tile goes: C_T + column max, or the row-major debug matrix
  bool live;
  int lane, n_cols, num_envs;
  float* ct;         // &C_T[0][tile0 + lane]
  uint32_t* colmax;  // this CTA's shared-memory column maxima (ordered-float keys), one writer warp per column
  float* out_row;    // &out[tile0 + lane][0]
  __device__ __forceinline__ void emit(int col, float c) const {
    if (MODE == kEvalRowMajor) {
      if (live) out_row[col] = c;
    } else {
      if (live) ct[(size_t)col * num_envs] = c;
      const uint32_t m = __reduce_max_sync(0xffffffffu, live ? float_to_ordered(c) : 0u);
      // column `col` is always handled by the same warp of this CTA: a plain read-modify-write suffices
      if (lane == 0 && m > colmax[col]) colmax[col] = m;
    }
  }
};

// Operation order follows the cited reference lines; every intermediate is rounded to fp32 exactly where
// torch materialises a tensor.
template <int OP>
__device__ __forceinline__ float column_value(const TermCtx& c, int lc) {
  const int id = c.ids[lc];
  const int row = c.row;
  switch (OP) {
    case CATB200_OP_GENERIC:
      return c.v0.at(row, id);
    case CATB200_OP_ABS_MINUS:  // constraints.py:30,64,75,85
      return __fsub_rn(fabsf(c.v0.f32(row, id)), c.p0);
    case CATB200_OP_ABSDIFF_MINUS:  // constraints.py:176-181
      return __fsub_rn(fabsf(__fsub_rn(c.v0.f32(row, id), c.v1.f32(row, id))), c.p0);
    case CATB200_OP_ABSDIFF_MINUS_GATE_Y:  // constraints.py:42-53
      return __fmul_rn(__fsub_rn(fabsf(__fsub_rn(c.v0.f32(row, id), c.v1.f32(row, id))), c.p0), c.gate);
    case CATB200_OP_ACTION_RATE:  // constraints.py:191-198 (true division by step_dt)
      return __fsub_rn(__fdiv_rn(fabsf(__fsub_rn(c.v0.f32(row, id), c.v1.f32(row, id))), c.p1), c.p0);
    case CATB200_OP_COMPONENT_GT:  // constraints.py:94
      return c.v0.f32(row, id) > c.p0 ? 1.0f : 0.0f;
    case CATB200_OP_CONTACT_ANY: {  // constraints.py:103-110; ids index the peak table
      bool any = false;
      for (int b = 0; b < c.n_ids; ++b) any |= c.peaks[c.ids[b] * kTile + c.lane] > c.p0;
      return any ? 1.0f : 0.0f;
    }
    case CATB200_OP_NORM2_MINUS:  // constraints.py:119
      return __fsub_rn(norm2(c.v0.f32(row, 0), c.v0.f32(row, 1)), c.p0);
    case CATB200_OP_AIR_TIME: {  // constraints.py:129-141
      const float td = c.v1.at(row, id) != 0.0f ? 1.0f : 0.0f;
      return __fmul_rn(__fmul_rn(__fsub_rn(c.p0, c.v0.f32(row, id)), td), c.gate);
    }
    case CATB200_OP_N_CONTACT: {  // constraints.py:151-168
      int n = 0;
      for (int b = 0; b < c.n_ids; ++b) n += c.peaks[c.ids[b] * kTile + c.lane] > c.p2 ? 1 : 0;
      return __fmul_rn(fabsf((float)n - c.p0), c.gate);
    }
    case CATB200_OP_FORCE_PEAK_MINUS:  // constraints.py:207-210
      return __fsub_rn(c.peaks[id * kTile + c.lane], c.p0);
    case CATB200_OP_LIMIT_MINUS:  // constraints.py:220
      return __fsub_rn(c.p0, c.v0.f32(row, id));
    case CATB200_OP_ABS_MINUS_GATE_STILL:  // constraints.py:231-235
      return __fmul_rn(__fsub_rn(fabsf(c.v0.f32(row, id)), c.p0), c.gate);
    default:
      return 0.0f;
  }
}

template <int OP, int MODE>
__device__ __forceinline__ void term_columns(const TermCtx& c, const ColumnSink<MODE>& sink) {
  for (int lc = c.first; lc < c.n_cols; lc += c.n_warps) sink.emit(c.col0 + lc, column_value<OP>(c, lc));
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Persistent, double-buffered: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; while tile i is
// evaluated out of one shared-memory buffer the bulk async copies of tile i+1 land in the other.  Column
// maxima are kept per CTA in shared memory and flushed with one atomicMax per column per CTA at the end.
template <int MODE>
__global__ void __launch_bounds__(kEvalThreadsMax)
cat_eval_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                int num_envs, float* __restrict__ running_max, int* __restrict__ rm_init,
                CatWorkspace ws, float* __restrict__ out_rowmajor) {
  extern __shared__ __align__(128) uint8_t smem_all[];
  __shared__ uint32_t s_colmax[CATB200_MAX_COLS];
  __shared__ __align__(8) unsigned long long s_bar[2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_warps = blockDim.x >> 5, n_threads = blockDim.x;  // 8 or 16 warps (power of two)
  const int n_tiles = (num_envs + kTile - 1) / kTile;
  const int buf_bytes = plan.smem_bar_off;  // sources + peak table of one buffer (16-byte multiple)
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);

  for (int c = threadIdx.x; c < plan.n_cols; c += n_threads) s_colmax[c] = 0u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // a source tile can use the bulk copy engine if it is contiguous and 16-byte aligned (full tiles only)
  auto bulk_ok = [&](int s, int tile0, int rows) -> bool {
    const catb200_source_t& src = plan.sources[s];
    const int es = src.dtype == CATB200_U8 ? 1 : 4;
    const uint8_t* g = static_cast<const uint8_t*>(src.ptr) + (size_t)tile0 * src.row_stride * es;
    return rows == kTile && src.row_stride == src.row_len && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  };
  // thread 0: arm buffer `b`'s barrier and start the bulk copies of tile `t`
  auto issue = [&](int t, int b) {
    const int tile0 = t * kTile, rows = min(kTile, num_envs - tile0);
    uint8_t* dst = smem_all + (size_t)b * buf_bytes;
    uint32_t total = 0;
    for (int s = 0; s < plan.n_sources; ++s)
      if (bulk_ok(s, tile0, rows)) total += kTile * plan.sources[s].row_len * (plan.sources[s].dtype == CATB200_U8 ? 1 : 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar0 + 8 * b), "r"(total) : "memory");
    for (int s = 0; s < plan.n_sources; ++s) {
      if (!bulk_ok(s, tile0, rows)) continue;
      const catb200_source_t& src = plan.sources[s];
      const int es = src.dtype == CATB200_U8 ? 1 : 4;
      bulk_copy_g2s((uint32_t)__cvta_generic_to_shared(dst + src.smem_off),
                    static_cast<const uint8_t*>(src.ptr) + (size_t)tile0 * src.row_len * es, kTile * src.row_len * es,
                    bar0 + 8 * b);
    }
  };

  if (threadIdx.x == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int b = it & 1;
    const uint8_t* smem = smem_all + (size_t)b * buf_bytes;
    uint8_t* smem_w = smem_all + (size_t)b * buf_bytes;
    float* peaks = reinterpret_cast<float*>(smem_w + plan.smem_peak_off);  // [n_peaks][32]
    const int tile0 = t * kTile;
    const int rows = min(kTile, num_envs - tile0);
    // prefetch the next tile into the other buffer (its previous contents were consumed before the barrier that
    // ended the previous iteration)
    if (threadIdx.x == 0 && t + (int)gridDim.x < n_tiles) issue(t + gridDim.x, b ^ 1);
    // sources that cannot use the bulk engine (strided views, unaligned bases, the ragged last tile): cooperative copy
    for (int s = 0; s < plan.n_sources; ++s) {
      if (bulk_ok(s, tile0, rows)) continue;
      const catb200_source_t& src = plan.sources[s];
      const int total = rows * src.row_len;
      if (src.dtype == CATB200_F32) {
        const float* g = static_cast<const float*>(src.ptr);
        float* dst = reinterpret_cast<float*>(smem_w + src.smem_off);
        for (int f = threadIdx.x; f < total; f += n_threads) {
          const int r = f / src.row_len, e = f - r * src.row_len;
          dst[f] = __ldg(g + (size_t)(tile0 + r) * src.row_stride + e);
        }
      } else {
        const uint8_t* g = static_cast<const uint8_t*>(src.ptr);
        for (int f = threadIdx.x; f < total; f += n_threads) {
          const int r = f / src.row_len, e = f - r * src.row_len;
          smem_w[src.smem_off + f] = g[(size_t)(tile0 + r) * src.row_stride + e];
        }
      }
    }
    // wait for this buffer's bulk copies: one lane sleeps on the mbarrier (suspend-time hint), the CTA barrier
    // releases the rest, then every thread performs one already-satisfied acquire of its own
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
    if (threadIdx.x == 0) {
      while (!try_wait()) {
      }
    }
    __syncthreads();
    while (!try_wait()) {
    }

    const bool live = lane < rows;
    const int row = live ? lane : 0;

    // ---- phase A: contact-force peaks, once per (history tensor, body) pair referenced by any term
    for (int p = warp; p < plan.n_peaks; p += n_warps) {
      const int s = plan.peak_src[p];
      peaks[p * kTile + lane] = force_peak(view_of(plan, smem, s), plan.sources[s].aux, row, plan.peak_body[p]);
    }
    __syncthreads();

    // ---- phase B: terms.  Term-level scalars and gates are computed once per term; the term's columns are
    //      dealt to the warps by global column index (lane = env, so the op dispatch never diverges) and the
    //      op switch sits outside the column loop.
    ColumnSink<MODE> sink;
    sink.live = live;
    sink.lane = lane;
    sink.n_cols = plan.n_cols;
    sink.num_envs = num_envs;
    sink.ct = MODE == kEvalStep ? ws.c_t + tile0 + lane : nullptr;
    sink.colmax = s_colmax;
    sink.out_row = MODE == kEvalRowMajor ? out_rowmajor + (size_t)(tile0 + lane) * plan.n_cols : nullptr;
    for (int ti = 0; ti < plan.n_terms; ++ti) {
      const catb200_term_t& t2 = plan.terms[ti];
      const int n_cols = t2.n_cols, col0 = t2.col_offset, op = t2.op;
      const int first = (warp - (col0 & (n_warps - 1))) & (n_warps - 1);  // first local column of this warp
      if (first >= n_cols) continue;
      TermCtx c;
      c.ids = t2.ids;
      c.n_ids = t2.n_ids;
      c.first = first;
      c.n_cols = n_cols;
      c.col0 = col0;
      c.row = row;
      c.lane = lane;
      c.n_warps = n_warps;
      c.p0 = t2.p0;
      c.p1 = t2.p1;
      c.p2 = t2.p2;
      c.peaks = peaks;
      c.v0 = view_of(plan, smem, t2.src0);
      c.v1 = t2.src1 != 0xff ? view_of(plan, smem, t2.src1) : c.v0;
      c.gate = 1.0f;  // command-dependent factor shared by all columns of the term
      if (t2.src2 != 0xff) {
        const SrcView vc = view_of(plan, smem, t2.src2);
        if (op == CATB200_OP_ABSDIFF_MINUS_GATE_Y) {
          c.gate = fabsf(vc.f32(row, 1)) < c.p1 ? 1.0f : 0.0f;  // constraints.py:46-53
        } else {
          const float cn = norm3(vc.f32(row, 0), vc.f32(row, 1), vc.f32(row, 2));
          c.gate = op == CATB200_OP_ABS_MINUS_GATE_STILL ? (cn < c.p1 ? 1.0f : 0.0f) : (cn > c.p1 ? 1.0f : 0.0f);
        }
      }
      switch (op) {
        case CATB200_OP_GENERIC: term_columns<CATB200_OP_GENERIC>(c, sink); break;
        case CATB200_OP_ABS_MINUS: term_columns<CATB200_OP_ABS_MINUS>(c, sink); break;
        case CATB200_OP_ABSDIFF_MINUS: term_columns<CATB200_OP_ABSDIFF_MINUS>(c, sink); break;
        case CATB200_OP_ABSDIFF_MINUS_GATE_Y: term_columns<CATB200_OP_ABSDIFF_MINUS_GATE_Y>(c, sink); break;
        case CATB200_OP_ACTION_RATE: term_columns<CATB200_OP_ACTION_RATE>(c, sink); break;
        case CATB200_OP_COMPONENT_GT: term_columns<CATB200_OP_COMPONENT_GT>(c, sink); break;
        case CATB200_OP_CONTACT_ANY: term_columns<CATB200_OP_CONTACT_ANY>(c, sink); break;
        case CATB200_OP_NORM2_MINUS: term_columns<CATB200_OP_NORM2_MINUS>(c, sink); break;
        case CATB200_OP_AIR_TIME: term_columns<CATB200_OP_AIR_TIME>(c, sink); break;
        case CATB200_OP_N_CONTACT: term_columns<CATB200_OP_N_CONTACT>(c, sink); break;
        case CATB200_OP_FORCE_PEAK_MINUS: term_columns<CATB200_OP_FORCE_PEAK_MINUS>(c, sink); break;
        case CATB200_OP_LIMIT_MINUS: term_columns<CATB200_OP_LIMIT_MINUS>(c, sink); break;
        case CATB200_OP_ABS_MINUS_GATE_STILL: term_columns<CATB200_OP_ABS_MINUS_GATE_STILL>(c, sink); break;
        default: break;
      }
    }
    __syncthreads();  // every warp is done with this buffer before the next iteration's prefetch overwrites it
  }

  if (MODE == kEvalStep) {
    // ---- flush this CTA's column maxima (one atomic per column per CTA, spread over scratch rows), then the
    //      last CTA folds everything into the Polyak running max (constraint_manager.py:55-61)
    int n_groups = 1;  // power of two; a few hundred CTAs per scratch row keep the atomic queues short
    while (n_groups < kMaxGroups && n_groups * 256 <= (int)gridDim.x) n_groups <<= 1;
    uint32_t* grow = ws.colmax + (blockIdx.x & (n_groups - 1)) * CATB200_MAX_COLS;
    for (int c = threadIdx.x; c < plan.n_cols; c += n_threads)
      if (s_colmax[c] != 0u) atomicMax(grow + c, s_colmax[c]);
    if (last_block_ticket_grouped(ws.ticket, gridDim.x)) {
      const int groups = n_groups;
      for (int col = threadIdx.x; col < plan.n_cols; col += n_threads) {
        uint32_t key = 0u;
        int gi = 0;
        for (; gi + 4 <= groups; gi += 4) {  // 4 independent exchanges in flight
          const uint32_t k0 = atomicExch(&ws.colmax[gi * CATB200_MAX_COLS + col], 0u);
          const uint32_t k1 = atomicExch(&ws.colmax[(gi + 1) * CATB200_MAX_COLS + col], 0u);
          const uint32_t k2 = atomicExch(&ws.colmax[(gi + 2) * CATB200_MAX_COLS + col], 0u);
          const uint32_t k3 = atomicExch(&ws.colmax[(gi + 3) * CATB200_MAX_COLS + col], 0u);
          key = max(max(key, max(k0, k1)), max(k2, k3));
        }
        for (; gi < groups; ++gi) key = max(key, atomicExch(&ws.colmax[gi * CATB200_MAX_COLS + col], 0u));
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
  }
}

// probability of one column value (constraint_manager.py:64-72)
__device__ __forceinline__ float violation_prob(float c, float rm, float min_p, float span) {
  if (!(c > 0.0f)) return 0.0f;
  float x = __fdiv_rn(c, rm);
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return __fadd_rn(min_p, __fmul_rn(x, span));
}

// One CTA per 32 envs (lane = env); the 4 warps split the statistics slots (terms) so that the dependent
// load chains are 4x shorter and 4x more loads are in flight than with one thread per env.
__global__ void __launch_bounds__(kEvalThreads)
cat_apply_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ c_t,
                 float* __restrict__ episode_sums, float* __restrict__ mean_values,
                 float* __restrict__ cstr_prob, const float* __restrict__ raw_reward,
                 const uint8_t* __restrict__ reset_buf, float* __restrict__ reward_out,
                 float* __restrict__ dones_out) {
  __shared__ float s_rm[CATB200_MAX_COLS];
  __shared__ float s_part[kEvalWarps][kTile];
  for (int c = threadIdx.x; c < plan.n_cols; c += kEvalThreads) s_rm[c] = running_max[c];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kTile + lane;
  const bool live = i < num_envs;
  float overall = -INFINITY;
  if (live) {
    for (int slot = warp; slot < plan.n_slots; slot += kEvalWarps) {
      const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
      const float span = prm.span[slot];
      const size_t k = (size_t)slot * num_envs + i;
      const float es = episode_sums[k], mv = mean_values[k];  // requested before the columns: all in flight together
      float tmax = -INFINITY;
      int col = c0;
      for (; col + 4 <= c1; col += 4) {
        const float v0 = __ldcs(c_t + (size_t)col * num_envs + i);
        const float v1 = __ldcs(c_t + (size_t)(col + 1) * num_envs + i);
        const float v2 = __ldcs(c_t + (size_t)(col + 2) * num_envs + i);
        const float v3 = __ldcs(c_t + (size_t)(col + 3) * num_envs + i);
        tmax = fmaxf(tmax, fmaxf(fmaxf(violation_prob(v0, s_rm[col], prm.min_p, span), violation_prob(v1, s_rm[col + 1], prm.min_p, span)),
                                 fmaxf(violation_prob(v2, s_rm[col + 2], prm.min_p, span), violation_prob(v3, s_rm[col + 3], prm.min_p, span))));
      }
      for (; col < c1; ++col) tmax = fmaxf(tmax, violation_prob(__ldcs(c_t + (size_t)col * num_envs + i), s_rm[col], prm.min_p, span));
      episode_sums[k] = __fadd_rn(es, tmax > 0.0f ? 1.0f : 0.0f);  // :226
      mean_values[k] = __fadd_rn(mv, tmax);                          // :227
      overall = fmaxf(overall, tmax);
    }
  }
  s_part[warp][lane] = overall;
  __syncthreads();
  if (warp != 0 || !live) return;
#pragma unroll
  for (int w = 1; w < kEvalWarps; ++w) overall = fmaxf(overall, s_part[w][lane]);
  cstr_prob[i] = overall;
  if (raw_reward != nullptr) {
    // cat_env.py:102-107: reward = clip(reward * (1 - p), min=0); dones = p; :121 dones[reset] = 1
    reward_out[i] = fmaxf(__fmul_rn(raw_reward[i], __fsub_rn(1.0f, overall)), 0.0f);
    dones_out[i] = (reset_buf != nullptr && reset_buf[i]) ? 1.0f : overall;
  }
}

__global__ void __launch_bounds__(kApplyThreads)
cat_probs_kernel(const __grid_constant__ catb200_plan_t plan, const __grid_constant__ catb200_cat_params_t prm,
                 int num_envs, const float* __restrict__ running_max, const float* __restrict__ c_t,
                 float* __restrict__ probs_out) {
  const int i = blockIdx.x * kApplyThreads + threadIdx.x;
  if (i >= num_envs) return;
  for (int slot = 0; slot < plan.n_slots; ++slot) {
    const int c0 = plan.slot_col_begin[slot], c1 = plan.slot_col_begin[slot + 1];
    for (int col = c0; col < c1; ++col) {
      const float c = c_t[(size_t)col * num_envs + i];
      probs_out[(size_t)i * plan.n_cols + col] = violation_prob(c, running_max[col], prm.min_p, prm.span[slot]);
    }
  }
}

// ---- ConstraintManager.reset (constraint_manager.py:190-211) -----------------------------------------
// grid = n_slots CTAs; each reduces its statistics row over the selected envs in double precision
// (torch's own fp32 reduction order is implementation defined; parity tolerance 1e-5 relative) and
// then zeroes the selected entries.
constexpr int kResetThreads = 256;
constexpr int kResetChunk = 1024;  // selected envs (or mask entries) per CTA

struct ResetScratch {  // per slot: double sums + count; zero between launches
  double v[CATB200_MAX_TERMS], p[CATB200_MAX_TERMS];
  unsigned long long cnt[CATB200_MAX_TERMS];
  unsigned int ticket;
};

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
  if (last_block_ticket(&sc->ticket, gridDim.x * gridDim.y)) {
    for (int s2 = threadIdx.x; s2 < n_slots; s2 += kResetThreads) {
      const double v = __longlong_as_double(atomicExch((unsigned long long*)&sc->v[s2], 0ull));
      const double p = __longlong_as_double(atomicExch((unsigned long long*)&sc->p[s2], 0ull));
      const unsigned long long c = atomicExch(&sc->cnt[s2], 0ull);
      // empty selection -> mean of nothing = NaN, like torch
      const float mv = (float)(v / (double)c), mp = (float)(p / (double)c);
      out[2 * s2] = __fmul_rn(mv, 100.0f);
      out[2 * s2 + 1] = mp;
    }
  }
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
  plan->slot_col_begin[slots] = (uint16_t)col;
  plan->n_cols = col;
  plan->n_slots = slots;
  plan->smem_peak_off = off;
  off += plan->n_peaks * kTile * 4;
  off = (off + 15) & ~15;
  plan->smem_bar_off = off;
  off += 16;
  plan->smem_bytes = off;
  if (2 * plan->smem_bar_off > 200 * 1024) return CATB200_ERR_UNSUPPORTED;  // two staging buffers must fit
  return CATB200_OK;
}

size_t catb200_cat_workspace_bytes(int32_t num_envs, int32_t n_cols) {
  if (num_envs < 0 || n_cols < 0) return 0;
  return 256 + align256(sizeof(uint32_t) * CATB200_MAX_COLS * kMaxGroups) + align256(sizeof(float) * (size_t)num_envs * n_cols);
}

static int launch_eval(const catb200_plan_t* plan, const catb200_cat_params_t* prm, int num_envs, float* running_max,
                       int* rm_init, CatWorkspace ws, float* out_rowmajor, int mode, cudaStream_t stream) {
  const size_t smem = 2 * (size_t)plan->smem_bar_off;  // two staging buffers (sources + peak table each)
  const int n_tiles = (num_envs + kTile - 1) / kTile;
  // persistent grid: as many CTAs as fit (shared memory / 2048 threads per SM), never more than tiles
  // small N (about one tile per SM-resident CTA): 16 warps shorten the per-tile critical path; large N: 8 warps
  // spend fewer instructions on per-warp term prologues (throughput bound there)
  const int threads = n_tiles <= 4 * kNumSMs ? kEvalThreadsMax : kEvalThreads;
  const int per_sm = (int)max((size_t)1, min((size_t)(2048 / threads), (size_t)(220 * 1024) / max(smem + 2048, (size_t)1)));
  const int grid = min(n_tiles, kNumSMs * per_sm);
  if (mode == kEvalStep) {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalStep>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalStep><<<grid, threads, smem, stream>>>(*plan, *prm, num_envs, running_max, rm_init, ws, nullptr);
  } else {
    if (smem > 48 * 1024)
      CATB200_CUDA_TRY(cudaFuncSetAttribute(cat_eval_kernel<kEvalRowMajor>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cat_eval_kernel<kEvalRowMajor><<<grid, threads, smem, stream>>>(*plan, *prm, num_envs, nullptr, nullptr, ws, out_rowmajor);
  }
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_step(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                     float* running_max, int32_t* rm_init, float* episode_sums, float* mean_values,
                     float* cstr_prob, const float* raw_reward, const uint8_t* reset_buf, float* reward_out,
                     float* dones_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !rm_init || !episode_sums || !mean_values || !cstr_prob ||
      !workspace)
    return CATB200_ERR_INVALID_ARGUMENT;
  if (plan->n_cols <= 0 || plan->smem_bytes <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  if (raw_reward && (!reward_out || !dones_out)) return CATB200_ERR_INVALID_ARGUMENT;
  if (workspace_bytes < catb200_cat_workspace_bytes(num_envs, plan->n_cols)) return CATB200_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = as_stream(stream);
  CatWorkspace ws = carve(workspace, num_envs);
  int rc = launch_eval(plan, params, num_envs, running_max, rm_init, ws, nullptr, kEvalStep, st);
  if (rc != CATB200_OK) return rc;
  const int grid = (num_envs + kTile - 1) / kTile;
  cat_apply_kernel<<<grid, kEvalThreads, 0, st>>>(*plan, *params, num_envs, running_max, ws.c_t, episode_sums,
                                                   mean_values, cstr_prob, raw_reward, reset_buf, reward_out,
                                                   dones_out);
  CATB200_LAUNCH_CHECK();
  return CATB200_OK;
}

int catb200_cat_eval_terms(const catb200_plan_t* plan, int32_t num_envs, float* out, void* stream) {
  if (!plan || num_envs <= 0 || !out || plan->n_cols <= 0) return CATB200_ERR_INVALID_ARGUMENT;
  catb200_cat_params_t dummy = {};
  CatWorkspace ws = {};
  return launch_eval(plan, &dummy, num_envs, nullptr, nullptr, ws, out, kEvalRowMajor, as_stream(stream));
}

int catb200_cat_probs(const catb200_plan_t* plan, const catb200_cat_params_t* params, int32_t num_envs,
                      const float* running_max, float* probs_out, const void* workspace, void* stream) {
  if (!plan || !params || num_envs <= 0 || !running_max || !probs_out || !workspace) return CATB200_ERR_INVALID_ARGUMENT;
  CatWorkspace ws = carve(const_cast<void*>(workspace), num_envs);
  const int grid = (num_envs + kApplyThreads - 1) / kApplyThreads;
  cat_probs_kernel<<<grid, kApplyThreads, 0, as_stream(stream)>>>(*plan, *params, num_envs, running_max, ws.c_t, probs_out);
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

// PTX wrappers for the Blackwell tensor-core path (tcgen05 / TMEM / TMA / mbarrier), shared by tc_gemm.cu and
// tc_fwd3.cu.  Bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mma.cuh"

namespace catb200 {

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"  // suspend-time hint: sleep, do not spin
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a barrier that is not reached within ~4 s (a protocol bug, never a legitimate wait) traps, which fails the
// launch with an error instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// The same with an L2 eviction-priority hint (createpolicy descriptors as constants: the encodings CUTLASS' TMA::CacheHintSm90
// uses).  The activations of a minibatch (2 x 235 MB in tf32) do not fit the 126 MB L2; the hints say which of them the NEXT
// launch re-reads (evict_last: the tensor a kernel has just produced, the weights) and which are dead after this read
// (evict_first), so that the former are not pushed out by the latter.
constexpr unsigned long long kL2EvictNormal = 0x1000000000000000ull;
constexpr unsigned long long kL2EvictFirst = 0x12F0000000000000ull;
constexpr unsigned long long kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, unsigned long long hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1, unsigned long long hint) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;\n" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "l"(hint)
               : "memory");
}
// shared -> global tensor store of one box (coordinates {c0 = column, c1 = row}); completion tracked by bulk groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane_base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// wait for outstanding tcgen05.ld; the registers are listed as in/out operands so that no use of them can be
// scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
// ELU for the epilogue: exp through one MUFU (ex2.approx.ftz of v * log2 e; abs error ~1e-7, far below the bf16
// rounding of the result).  __expf adds a range fix-up for results below 2^-126 (4 more instructions and two
// predicates per element); flushed to zero they give ELU = -1 exactly, which is what the fix-up would round to.
__device__ __forceinline__ float elu_fast(float v) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(t) : "f"(v * 1.4426950408889634f));
  return v > 0.0f ? v : t - 1.0f;
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in bits [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48),
// layout type in [61,64): SWIZZLE_128B = 2 (16-byte chunks XOR-ed with the row index, 8-row atoms) or
// SWIZZLE_128B_BASE32B = 1 (32-byte chunks, 4-row atoms: the only layout the tensor core accepts for MN-major
// 32-bit operands; TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = kLayoutSw128) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A / B format in [7,10) / [10,13)
// (0 = f16, 1 = bf16, 2 = tf32), a_major bit 15, b_major bit 16 (1 = MN-major), N >> 3 in [17,23), M >> 4 in [24,29).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- CTA pairs (cta_group::2): two SMs of one TPC run one 256 x N tile --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA load of a CTA pair: lands in THIS CTA's shared memory, completes on the mbarrier `bar_cluster` (a shared::cluster
// address: the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar), "h"(mask)
               : "memory");
}

// ---- operand precision of the hidden-layer GEMMs ------------------------------------------------------------
// kPrecTf32 (default): fp32 storage, operands rounded to tf32 (10-bit mantissa) where they are stored, tcgen05.mma
//   kind::tf32, fp32 accumulation -- the reference's own GPU numerics (scripts/clean_rl/train.py:86-87).
// kPrecBf16: bf16 storage and operands (kind::f16), fp32 accumulation -- half the operand bytes, 8-bit mantissa.
// Either way one shared-memory row of an operand tile is 128 bytes (SWIZZLE_128B) and one UMMA K-step is 32 bytes.
enum Prec { kPrecBf16 = 0, kPrecTf32 = 1 };

template <int PREC>
struct PrecT;
template <>
struct PrecT<kPrecBf16> {
  using T = bf16;
  static constexpr int kBK = 64;     // reduction elements per 128-byte row = per pipeline stage
  static constexpr int kUmmaK = 16;  // reduction elements per tcgen05.mma
  static constexpr uint32_t kFmt = 1;
};
template <>
struct PrecT<kPrecTf32> {
  using T = float;
  static constexpr int kBK = 32;
  static constexpr int kUmmaK = 8;
  static constexpr uint32_t kFmt = 2;
};

template <int PREC>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == kPrecBf16) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}
// the same for a CTA pair: M = 256 (128 rows per CTA), A from each CTA's own shared memory, B halves from both
template <int PREC>
__device__ __forceinline__ void umma_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (PREC == kPrecBf16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

// round-to-nearest fp32 -> tf32 (low 13 mantissa bits cleared); operands are rounded where they are stored so that
// the tensor core's own truncation of the fp32 bit pattern never discards anything
__device__ __forceinline__ float round_tf32(float x) {
  // cvt.rna.tf32.f32 (round the magnitude to a 10-bit mantissa, ties away from zero) as two integer instructions: half an
  // ulp of the kept mantissa is added to the sign-magnitude bits and the dropped bits are masked.  ptxas expands the cvt
  // to the same add + mask behind an |x| < inf test; infinities survive this form unchanged as well (0x7f800000 + 0x1000
  // masks back to 0x7f800000), so the test buys nothing on the epilogue paths, where every output element is rounded.
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// fire-and-forget vector reduction into global memory (sm_90+): 16 bytes per instruction
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;\n" ::"l"(addr), "f"(a) : "memory");
}
// 32 lanes x 8 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}

}  // namespace catb200

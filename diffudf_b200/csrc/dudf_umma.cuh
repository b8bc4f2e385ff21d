// Thin PTX wrappers for the Blackwell (sm_100a) tensor-core path: mbarrier, bulk async copy (TMA
// engine, UBLKCP), tensor memory (TMEM) and tcgen05.mma with shared-memory matrix descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// a spin-wait that cannot hang the GPU: after ~4 s it records where it was stuck and traps
static __device__ unsigned int g_watchdog_code = 0;
#define UMMA_WATCHDOG_CYCLES (8000000000ll)
#ifndef DUDF_PRODUCER_SLEEP_NS
#define DUDF_PRODUCER_SLEEP_NS 64
#endif

__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > UMMA_WATCHDOG_CYCLES) {
      g_watchdog_code = code;
      __threadfence_system();
      __trap();
    }
  }
}

// the same wait for a role that runs far ahead of its consumer (weight producer): back off between polls instead of spinning
// at full issue rate next to the epilogue warps (the board is power-capped under this kernel, tools/clock_probe.py)
__device__ __forceinline__ void mbar_wait_relaxed(void* bar, uint32_t parity, unsigned code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(DUDF_PRODUCER_SLEEP_NS);       // measured neutral for the kernel times (0 vs 64 ns), kept for the issue slots it frees
    if (clock64() - t0 > UMMA_WATCHDOG_CYCLES) {
      g_watchdog_code = code;
      __threadfence_system();
      __trap();
    }
  }
}

// asks the L2 for a contiguous global range (16-byte aligned, size a multiple of 16): no registers, no completion to wait for
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// generic-proxy writes to shared memory -> visible to the async proxy (tcgen05.mma / bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy global -> shared, completion signalled on an mbarrier (complete_tx)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same copy with an L2 cache policy (createpolicy) for the global reads
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// the same copy, delivered to the same shared-memory offset (data and complete_tx) of every CTA in cta_mask
__device__ __forceinline__ void bulk_g2s_multicast(void* dst_smem, const void* src_gmem, uint32_t bytes, void* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tensor memory ----
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, N consecutive columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// 64 consecutive columns in one instruction: columns 0-31 -> r0, 32-63 -> r1
__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t* r0, uint32_t* r1) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15]), "=r"(r0[16]), "=r"(r0[17]), "=r"(r0[18]), "=r"(r0[19]), "=r"(r0[20]), "=r"(r0[21]), "=r"(r0[22]), "=r"(r0[23]), "=r"(r0[24]), "=r"(r0[25]), "=r"(r0[26]), "=r"(r0[27]), "=r"(r0[28]), "=r"(r0[29]), "=r"(r0[30]), "=r"(r0[31]), "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15]), "=r"(r1[16]), "=r"(r1[17]), "=r"(r1[18]), "=r"(r1[19]), "=r"(r1[20]), "=r"(r1[21]), "=r"(r1[22]), "=r"(r1[23]), "=r"(r1[24]), "=r"(r1[25]), "=r"(r1[26]), "=r"(r1[27]), "=r"(r1[28]), "=r"(r1[29]), "=r"(r1[30]), "=r"(r1[31])
      : "r"(taddr)
      : "memory");
}

// ---- descriptors ----
// shared-memory matrix descriptor, 128-byte swizzle.  The tile is a stack of 8-row x 128-byte
// swizzle atoms; `sbo` = byte distance between consecutive atoms along the strided (8-row group)
// dimension, `lbo` = byte distance between atoms along the leading dimension (only consulted for
// MN-major operands wider than one atom).  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B
  return d;
}
// instruction descriptor for kind::f16, fp16 (fmt 0) or bf16 (fmt 1) operands, fp32 accumulation
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int ab_fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)ab_fmt << 7) | ((uint32_t)ab_fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same instruction issued from warp-uniform code: every lane of a converged warp calls this with identical operands,
// one elected lane issues.  Keeps the operands in uniform registers (no per-instruction election loop in SASS).
__device__ __forceinline__ void mma_f16_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_warp(void* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// Four consecutive K = 16 steps of one 64-wide k chunk in ONE elected block: A advances by 32 bytes per step inside its
// 128-byte swizzle row (descriptor + 2), B (MN-major) by two 1 KB k-groups (descriptor + 128).  `first` = 0 starts a new
// accumulation with the first step.  One election and no descriptor traffic through general registers per step.
__device__ __forceinline__ void mma_f16_ss_k64_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t first) {
  asm volatile(
      "{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.s64 a1, %1, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 a3, %1, 6;\n\t"
      "add.s64 b1, %2, 128;\n\tadd.s64 b2, %2, 256;\n\tadd.s64 b3, %2, 384;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(first)
      : "memory");
}
// Split-precision pair: the same A slice (a weight chunk's K = 16 step) against TWO B tiles (hi and lo activations), step by
// step over a 64-wide k chunk.  The first instruction of each step latches A in the collector buffer (collector::a::fill), the
// second reuses it (collector::a::lastuse): A crosses shared memory once per step instead of twice — at M = N = 128 the two
// operands of an MMA already take the whole 128 B/clk of the shared-memory port, so every byte not re-read is tensor time.
__device__ __forceinline__ void mma_f16_ss_k64_pair_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t bh_desc, uint64_t bl_desc, uint32_t idesc,
                                                         uint32_t first) {
  asm volatile(
      "{\n\t.reg .pred q, p, t;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, c1, c2, c3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 t, 0, 0;\n\t"
      "add.s64 a1, %1, 2;\n\tadd.s64 a2, %1, 4;\n\tadd.s64 a3, %1, 6;\n\t"
      "add.s64 b1, %2, 128;\n\tadd.s64 b2, %2, 256;\n\tadd.s64 b3, %2, 384;\n\t"
      "add.s64 c1, %3, 128;\n\tadd.s64 c2, %3, 256;\n\tadd.s64 c3, %3, 384;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %4, p;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %3, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], a1, b1, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], a1, c1, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], a2, b2, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], a2, c2, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], a3, b3, %4, t;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], a3, c3, %4, t;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(bh_desc), "l"(bl_desc), "r"(idesc), "r"(first)
      : "memory");
}
// descriptor with the start address advanced by `bytes` (no carry out of the 14-bit address field for our tiles)
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
// all previously issued tcgen05.mma of this thread complete -> one arrival on the mbarrier
__device__ __forceinline__ void mma_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ... one arrival on the barrier at the same shared-memory offset of every CTA in cta_mask
__device__ __forceinline__ void mma_commit_multicast(void* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

__device__ __forceinline__ void mma_commit_multicast_warp(void* bar, uint16_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// register re-distribution between warpgroups (all warps of a 128-thread warpgroup must execute the same one)
template <int R>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// byte offset of element (row, k) inside a K-major [rows x 64] fp16 block laid out with 128B swizzle
__host__ __device__ constexpr uint32_t sw128_offset(uint32_t row, uint32_t k) {
  return row * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + ((k & 7u) << 1);
}

}  // namespace umma

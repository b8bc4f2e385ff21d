// Building blocks shared by the tcgen05 kernels (field queries, training forward, reverse sweep).
//
// Stored variables of a sine layer (w = omega of the layer):  u0 = w z0 (bias included), u_i = w z_i,
// v_ij = KAPPA w z_ij;   activations  a0 = sin u0, a_i = cos(u0) u_i, b_ij = cos(u0) v_ij - KAPPA sin(u0) u_i u_j.
// With the weights packed as fp16(w W) the MMA maps stored activations to stored pre-activations directly.
//
// Shared-memory B-operand tile (one sub-tile, 128 columns x 256 input neurons, fp16, 64 KB), MN-major:
// 8-row x 128-byte swizzle atoms, atom(nb = column/64, kg = k/8) at (nb*32 + kg) KB, inside an atom row k%8 holds
// 64 consecutive columns with its 16-byte chunks XOR-swizzled by k%8.  Seen from the weight-gradient GEMM
// (reduction over columns) each 32 KB half nb is a K-major [256 neurons][64 columns] operand image.
#pragma once
#include <cuda_fp16.h>
#include "dudf_common.cuh"
#include "dudf_umma.cuh"

namespace dudf {

constexpr int TC_CHUNK_BYTES = 128 * 64 * 2;   // one weight chunk: 128 neurons x 64 k, fp16
constexpr int TC_STAGES = 5;
constexpr int TC_THREADS = 384;            // warpgroups 0-1: epilogue (216 regs), warpgroup 2: producer + MMA issuer (56 regs)
constexpr int TC_REGS_EPI = 216;
constexpr int TC_REGS_AUX = 56;            // 256 * 216 + 128 * 56 = 62 464 <= 65 536 registers per SM
constexpr int TC_ACT_BYTES = 65536;
constexpr int TC_IMG_BYTES = 32768;            // [256][64] fp16 operand image
constexpr float TC_KAPPA = 0.125f;
constexpr float TC_KAPPA_INV = 8.0f;
constexpr int TC_DIR_FWD = 0, TC_DIR_BWD = 1, TC_DIR_BOTH = 2;

template <int NCH>
struct TcCfg {
  static constexpr int PT = (NCH == 1) ? 128 : (NCH == 4 ? 32 : 12);   // points per sub-tile
  static constexpr int NV = PT * NCH;                                  // columns in use (128, 128, 120)
  static constexpr int GC = (NCH == 10) ? 40 : 32;                     // columns per epilogue step
  static constexpr int NGRP = NV / GC;
  static constexpr int OFF_RING = 2 * TC_ACT_BYTES;
  static constexpr int OFF_WL = OFF_RING + TC_STAGES * TC_CHUNK_BYTES;
  static constexpr int OFF_XS = OFF_WL + 256 * 4;
  static constexpr int OFF_OS = OFF_XS + 2 * 128 * 3 * 4;              // xs sized for the largest sub-tile (a launch may mix orders)
                                                                       // os: [2][256] floats: outputs / seeds
  static constexpr int OFF_BAR = (OFF_OS + 2 * 256 * 4 + 15) / 16 * 16;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
};

__device__ __forceinline__ void tc_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t tc_pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ unsigned char* tc_tile_row(unsigned char* tile, int k) { return tile + (k >> 3) * 1024 + (k & 7) * 128; }
// byte offset (relative to the row) of column chunk c (8 columns each, c = 0..15)
__device__ __forceinline__ uint32_t tc_chunk_off(int c, uint32_t r7) { return (uint32_t)(c >> 3) * 32768u + ((((uint32_t)c & 7u) ^ r7) << 4); }

__device__ __forceinline__ float loss_scale_from(const float* seed_absmax) {
  const float m = seed_absmax ? *seed_absmax : 0.f;
  return (m > 0.f && isfinite(m)) ? exp2f(floorf(log2f(2048.f / m))) : 1.f;
}

// bulk copy shared -> global (TMA engine), grouped completion
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(umma::smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(umma::smem_u32(src_smem)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// drop a 128-byte line from L2 without writing it back (the data is dead: scratch that will be rewritten before its next read)
__device__ __forceinline__ void l2_discard_line(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- producer: weight chunks through the ring.  Phase j of a pair uses image j (forward) or the transposed
// ---- image of layer n_phase-j (reverse sweep).  dir: TC_DIR_FWD, TC_DIR_BWD, or TC_DIR_BOTH (fused training step:
// ---- the n_phase forward phases of a pair followed by its n_phase reverse phases).  REUSE: each of the 8 chunks of a layer is fetched once per sub-tile
// ---- pair (the MMA issuer uses a chunk for both sub-tiles before releasing it); otherwise once per sub-tile.
// CL > 1: the CTAs of a cluster run the same schedule; chunk c is fetched from L2 by CTA (c mod CL) only and
// multicast into the ring slot of every CTA of the cluster (its complete_tx lands on each CTA's full barrier);
// a slot is free again when the MMA issuers of ALL CTAs have released it (empty barriers count CL arrivals).
template <int CL = 1, bool REUSE = false>
__device__ __forceinline__ void tc_producer(const unsigned char* packed, unsigned char* ring, uint64_t* full, uint64_t* empty, int64_t rounds,
                                            int n_phase, int dir, uint32_t cta_rank = 0, int dbg = 0) {
  using namespace umma;
  uint32_t stage = 0, phase = 0, chunk = 0;
  const int n_tot = (dir == TC_DIR_BOTH) ? 2 * n_phase : n_phase;
  for (int64_t r = 0; r < rounds; ++r)
    for (int jt = 0; jt < n_tot; ++jt) {
      const bool backward = (dir == TC_DIR_BWD) || (jt >= n_phase);
      const int j = (jt >= n_phase) ? jt - n_phase : jt;
      const int idx = backward ? (n_phase + (n_phase - 1 - j)) : j;
      const unsigned char* src = packed + (size_t)idx * 8 * TC_CHUNK_BYTES;
      for (int rep = 0; rep < (REUSE ? 1 : 2); ++rep)
        for (int ck = 0; ck < 8; ++ck, ++chunk) {
          mbar_wait_relaxed(&empty[stage], phase ^ 1, 0x100 + stage);
          if ((dbg & 4) && chunk >= (uint32_t)TC_STAGES) {        // pipeline diagnostics: stale weights, no L2 -> SM traffic
            mbar_arrive(&full[stage]);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full[stage], TC_CHUNK_BYTES);
          if constexpr (CL == 1) {
            bulk_g2s(ring + stage * TC_CHUNK_BYTES, src + (size_t)ck * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &full[stage]);
          } else {
            if (chunk % CL == cta_rank)
              bulk_g2s_multicast(ring + stage * TC_CHUNK_BYTES, src + (size_t)ck * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &full[stage],
                                 (uint16_t)((1u << CL) - 1));
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// operand-image copy of a B tile for the weight-gradient GEMM (two 32 KB images), issued by lane 0 of the MMA warp BEFORE
// the MMAs that read the tile so that both proceed concurrently; tc_finish_subtile waits for its shared-memory reads
// before the accumulator is published (the epilogue overwrites the tile after that).
__device__ __forceinline__ void tc_copy_subtile(unsigned char* act, int s, unsigned char* img, int64_t ncb, int64_t cb0, int64_t pair, int layer,
                                                uint64_t img_policy) {
  if (img && (threadIdx.x & 31) == 0) {
    const int64_t cb = cb0 + (pair * 2 + s) * 2;
#pragma unroll
    for (int nb = 0; nb < 2; ++nb) {
      unsigned char* dst = img + ((size_t)layer * ncb + cb + nb) * TC_IMG_BYTES;
      if (img_policy) bulk_s2g_hint(dst, act + s * TC_ACT_BYTES + nb * TC_IMG_BYTES, TC_IMG_BYTES, img_policy);
      else bulk_s2g(dst, act + s * TC_ACT_BYTES + nb * TC_IMG_BYTES, TC_IMG_BYTES);
    }
    bulk_commit();
  }
  __syncwarp();
}
__device__ __forceinline__ void tc_finish_subtile(uint64_t* acc_ready, int s, bool copied) {
  if (copied) {
    if ((threadIdx.x & 31) == 0) bulk_wait_read0();
    __syncwarp();
  }
  umma::mma_commit_warp(&acc_ready[s]);
}

// event log for tools/trace_probe.py: (tag << 56 | aux << 48 | clock); one writer per region
constexpr uint32_t TC_TRACE_REGION = 8192;
__device__ __forceinline__ void tc_trace(unsigned long long* region, uint32_t& n, uint32_t tag, uint32_t aux) {
  if (region && n < TC_TRACE_REGION) {
    if ((threadIdx.x & 31) == 0)
      region[n] = ((unsigned long long)tag << 56) | ((unsigned long long)(aux & 0xff) << 48) | ((unsigned long long)clock64() & 0xffffffffffffull);
    ++n;
  }
}

// ---- MMA issuer: the WHOLE warp runs this role in converged, warp-uniform code and one elected lane issues each
// ---- tcgen05 instruction.  (Issued from a single-lane branch, every tcgen05.mma drags a register -> uniform-register
// ---- election loop and a recomputed descriptor with it: 107-160 clocks per instruction for 64 clocks of tensor work,
// ---- tools/umma_bench.py.  Warp-uniform issue with descriptors advanced by adds reaches 64.1.)
// REUSE = false: per layer  A.h0 A.h1 | B.h0 B.h1  (sub-tile A/B, neuron half h): a sub-tile's accumulator is published as
//   early as possible (training kernels).
// REUSE = true:  per layer  A.h0 B.h0 | A.h1 B.h1: the four chunks of a half stay in the ring for both sub-tiles and are
//   released after the second use, so weights cross L2 -> SM once per 256 columns (query kernel).
// img_f / img_b != null: bulk-copy each finished B tile of a forward / reverse phase (two 32 KB operand images) to
// img[layer][cb0 + 2*subtile + nb].
template <int CL = 1, bool REUSE = false>
__device__ __forceinline__ void tc_mma_role(unsigned char* act, unsigned char* ring, uint64_t* full, uint64_t* empty, uint64_t* act_ready,
                                            uint64_t* acc_ready, uint32_t tmem_base, int64_t rounds, int n_phase, unsigned char* img_f,
                                            unsigned char* img_b, int64_t ncb, int64_t cb0, int dir, int dbg = 0, uint64_t img_policy = 0,
                                            unsigned long long* trace = nullptr) {
  using namespace umma;
  uint32_t tn = 0;
  static_assert(TC_STAGES >= 5, "a neuron half (4 chunks) must fit in the ring with one slot to prefetch into");
  constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, /*A K-major*/ 0, /*B MN-major*/ 1);
  constexpr uint16_t mask = (uint16_t)((1u << CL) - 1);
  const uint64_t a_desc0 = make_desc_sw128(smem_u32(ring), 16, 1024);       // chunk in ring slot 0, k = 0
  const uint64_t b_desc0 = make_desc_sw128(smem_u32(act), 32768, 1024);     // sub-tile 0, k = 0
  const bool skip = (dbg & 2) != 0;                         // pipeline diagnostics (tools/pipe_probe.py): keep the protocol, no MMAs
  uint32_t stage = 0, phase = 0;                            // ring position of the next chunk to be consumed for the first time
  uint32_t act_phase = 0;                                   // bit s = parity of act_ready[s]
  const int n_tot = (dir == TC_DIR_BOTH) ? 2 * n_phase : n_phase;
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t pair = blockIdx.x + r * gridDim.x;
    for (int jt = 0; jt < n_tot; ++jt) {
      const bool backward = (dir == TC_DIR_BWD) || (jt >= n_phase);
      const int j = (jt >= n_phase) ? jt - n_phase : jt;
      const int layer = backward ? (n_phase - j) : j;
      unsigned char* img = (dbg & 16) ? nullptr : (backward ? img_b : img_f);     // dbg bit 4: diagnostics, no operand-image copies
      if constexpr (!REUSE) {
        for (int s = 0; s < 2; ++s) {
          mbar_wait(&act_ready[s], (act_phase >> s) & 1u, 0x200 + s);
          act_phase ^= 1u << s;
          tc_fence_after();
          tc_trace(trace, tn, 1, s);
          tc_copy_subtile(act, s, img, ncb, cb0, pair, layer, img_policy);
          const uint64_t b_desc = desc_advance(b_desc0, s * TC_ACT_BYTES);
          for (int h = 0; h < 2; ++h) {
            const uint32_t d_tmem = tmem_base + s * 256 + h * 128;
            for (int kb = 0; kb < 4; ++kb) {
              mbar_wait(&full[stage], phase, 0x300 + stage);       // bulk-copy completion: already visible to the async proxy
              if (!skip)
                mma_f16_ss_k64_warp(d_tmem, desc_advance(a_desc0, stage * TC_CHUNK_BYTES), desc_advance(b_desc, kb * 8 * 1024), idesc, kb != 0);
              if constexpr (CL == 1) mma_commit_warp(&empty[stage]);
              else mma_commit_multicast_warp(&empty[stage], mask);
              if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
          }
          tc_trace(trace, tn, 2, 2 + s);
          tc_finish_subtile(acc_ready, s, img != nullptr);
        }
      } else {
        for (int h = 0; h < 2; ++h)
          for (int s = 0; s < 2; ++s) {
            if (h == 0) {
              mbar_wait(&act_ready[s], (act_phase >> s) & 1u, 0x200 + s);
              act_phase ^= 1u << s;
              tc_fence_after();
              tc_trace(trace, tn, 1, s);
              tc_copy_subtile(act, s, img, ncb, cb0, pair, layer, img_policy);
            }
            const uint64_t b_desc = desc_advance(b_desc0, s * TC_ACT_BYTES);
            const uint32_t d_tmem = tmem_base + s * 256 + h * 128;
            uint32_t st = stage, ph = phase;
            for (int kb = 0; kb < 4; ++kb) {
              if (s == 0) mbar_wait(&full[st], ph, 0x300 + st);     // bulk-copy completion: already visible to the async proxy
              if (!skip)
                mma_f16_ss_k64_warp(d_tmem, desc_advance(a_desc0, st * TC_CHUNK_BYTES), desc_advance(b_desc, kb * 8 * 1024), idesc, kb != 0);
              if (s == 1) {
                if constexpr (CL == 1) mma_commit_warp(&empty[st]);
                else mma_commit_multicast_warp(&empty[st], mask);
              }
              if (++st == TC_STAGES) { st = 0; ph ^= 1; }
            }
            if (s == 1) { stage = st; phase = ph; }
            tc_trace(trace, tn, 2, h * 2 + s);
            if (h == 1) tc_finish_subtile(acc_ready, s, img != nullptr);
          }
      }
    }
  }
}

template <int GC>
__device__ __forceinline__ void tc_load_group(uint32_t taddr, float* u) {
  uint32_t r[32];
  umma::tmem_ld_x32(taddr, r);
  if constexpr (GC == 40) {
    uint32_t r2[8];
    umma::tmem_ld_x8(taddr + 32, r2);
    umma::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) u[32 + j] = __uint_as_float(r2[j]);
  } else {
    umma::tmem_ld_wait();
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) u[j] = __uint_as_float(r[j]);
}

// split form for software pipelining: issue the TMEM loads of the NEXT column group, work on the current one,
// then tc_ld_take() waits and hands the registers over
template <int GC>
struct TmemRegs {
  uint32_t a[32];
  uint32_t b[GC == 40 ? 8 : 1];
};
template <int GC>
__device__ __forceinline__ void tc_ld_issue(uint32_t taddr, TmemRegs<GC>& r) {
  umma::tmem_ld_x32(taddr, r.a);
  if constexpr (GC == 40) umma::tmem_ld_x8(taddr + 32, r.b);
}
template <int GC>
__device__ __forceinline__ void tc_ld_take(const TmemRegs<GC>& r, float* u) {
  umma::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) u[j] = __uint_as_float(r.a[j]);
  if constexpr (GC == 40) {
#pragma unroll
    for (int j = 0; j < 8; ++j) u[32 + j] = __uint_as_float(r.b[j]);
  }
}

// stored pre-activations u -> stored activations, in place
template <int NCH>
__device__ __forceinline__ void tc_act_point(float* u, float s, float c) {
  // written out channel by channel (xx,xy,xz,yy,yz,zz = 4..9): loops over symmetric index pairs keep the
  // array in local memory (nvcc 12.9 does not scalarise them)
  if constexpr (NCH >= 10) {
    const float ks = TC_KAPPA * s;
    const float kx = ks * u[1], ky = ks * u[2], kz = ks * u[3];
    u[4] = fmaf(c, u[4], -kx * u[1]);
    u[5] = fmaf(c, u[5], -kx * u[2]);
    u[6] = fmaf(c, u[6], -kx * u[3]);
    u[7] = fmaf(c, u[7], -ky * u[2]);
    u[8] = fmaf(c, u[8], -ky * u[3]);
    u[9] = fmaf(c, u[9], -kz * u[3]);
  }
  if constexpr (NCH >= 4) {
    u[1] = c * u[1];
    u[2] = c * u[2];
    u[3] = c * u[3];
  }
  u[0] = s;
}

// the same for accumulators that still carry a power-of-two scale 1 / inv on their derivative channels (split-precision path:
// weights are packed times 64): a_i = (c inv) z_i, b_ij = (c inv) z_ij - (KAPPA s inv^2) z_i z_j.  inv is a power of two, so every
// product equals the one tc_act_point forms from the unscaled values bit for bit; the value channel u[0] is not read.
template <int NCH>
__device__ __forceinline__ void tc_act_point_scaled(float* u, float s, float c, float inv) {
  const float ci = c * inv;
  if constexpr (NCH >= 10) {
    const float ks = (TC_KAPPA * inv * inv) * s;
    const float kx = ks * u[1], ky = ks * u[2], kz = ks * u[3];
    u[4] = fmaf(ci, u[4], -kx * u[1]);
    u[5] = fmaf(ci, u[5], -kx * u[2]);
    u[6] = fmaf(ci, u[6], -kx * u[3]);
    u[7] = fmaf(ci, u[7], -ky * u[2]);
    u[8] = fmaf(ci, u[8], -ky * u[3]);
    u[9] = fmaf(ci, u[9], -kz * u[3]);
  }
  if constexpr (NCH >= 4) {
    u[1] = ci * u[1];
    u[2] = ci * u[2];
    u[3] = ci * u[3];
  }
  u[0] = s;
}

// stored pre-activations of the first layer for GC/NCH consecutive points (pts: xyz triples)
template <int NCH, int GC>
__device__ __forceinline__ void tc_first_layer_group(float* u, const float* pts, float w0, float r0x, float r0y, float r0z, float b0) {
#pragma unroll
  for (int pp = 0; pp < GC / NCH; ++pp) {
    const float* pt = pts + pp * 3;
    float* up = u + pp * NCH;
    up[0] = w0 * fmaf(r0z, pt[2], fmaf(r0y, pt[1], fmaf(r0x, pt[0], b0)));
    if constexpr (NCH >= 4) { up[1] = w0 * r0x; up[2] = w0 * r0y; up[3] = w0 * r0z; }
#pragma unroll
    for (int ch = 4; ch < NCH; ++ch) up[ch] = 0.f;
  }
}

// pack GC fp32 values of one thread into halves and store them as GC/8 16-byte chunks of its tile row
// 16-byte store through a 32-bit shared-memory address (STS.128; a store through the generic pointer costs a 64-bit address
// add per chunk and goes down the generic path)
__device__ __forceinline__ void tc_sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <int GC>
__device__ __forceinline__ void tc_store_group(const float* v, unsigned char* trow, int chunk0, uint32_t r7) {
  const uint32_t base = umma::smem_u32(trow);
#pragma unroll
  for (int c8 = 0; c8 < GC / 8; ++c8)
    tc_sts128(base + tc_chunk_off(chunk0 + c8, r7), tc_pack_h2(v[8 * c8], v[8 * c8 + 1]), tc_pack_h2(v[8 * c8 + 2], v[8 * c8 + 3]),
              tc_pack_h2(v[8 * c8 + 4], v[8 * c8 + 5]), tc_pack_h2(v[8 * c8 + 6], v[8 * c8 + 7]));
}

// sine-jet of GC/NCH points (in place) and store into the B-operand tile.  PRECISE: explicit 2*pi reduction
// (first layer, arguments up to ~50 rad); otherwise MUFU on the raw argument (hidden layers, |u| of a few rad).
template <int NCH, int GC, bool PRECISE>
__device__ __forceinline__ void tc_emit_group(float* u, unsigned char* trow, int chunk0, uint32_t r7) {
#pragma unroll
  for (int pp = 0; pp < GC / NCH; ++pp) {
    float sn, cs;
    if constexpr (PRECISE) sincos_fast(u[pp * NCH], sn, cs);
    else { sn = __sinf(u[pp * NCH]); cs = __cosf(u[pp * NCH]); }
    tc_act_point<NCH>(u + pp * NCH, sn, cs);
  }
  tc_store_group<GC>(u, trow, chunk0, r7);
}

// adjoint of the above: ab = dL/d(stored activations) -> ub = dL/d(stored pre-activations)
template <int NCH>
__device__ __forceinline__ void tc_adj_point(const float* u, const float* ab, float* ub, float s, float c) {
  float u0 = c * ab[0];
  if constexpr (NCH >= 4) {
    const float ux = u[1], uy = u[2], uz = u[3];
    float bx = c * ab[1], by = c * ab[2], bz = c * ab[3];
    u0 = fmaf(-s, fmaf(ab[1], ux, fmaf(ab[2], uy, ab[3] * uz)), u0);
    if constexpr (NCH >= 10) {
      const float ks = TC_KAPPA * s, kc = TC_KAPPA * c;
      const float axx = ab[4], axy = ab[5], axz = ab[6], ayy = ab[7], ayz = ab[8], azz = ab[9];
      float acc2 = axx * fmaf(s, u[4], kc * ux * ux);
      acc2 = fmaf(axy, fmaf(s, u[5], kc * ux * uy), acc2);
      acc2 = fmaf(axz, fmaf(s, u[6], kc * ux * uz), acc2);
      acc2 = fmaf(ayy, fmaf(s, u[7], kc * uy * uy), acc2);
      acc2 = fmaf(ayz, fmaf(s, u[8], kc * uy * uz), acc2);
      acc2 = fmaf(azz, fmaf(s, u[9], kc * uz * uz), acc2);
      u0 -= acc2;
      bx -= ks * fmaf(2.f * axx, ux, fmaf(axy, uy, axz * uz));
      by -= ks * fmaf(axy, ux, fmaf(2.f * ayy, uy, ayz * uz));
      bz -= ks * fmaf(axz, ux, fmaf(ayz, uy, 2.f * azz * uz));
      ub[4] = c * axx; ub[5] = c * axy; ub[6] = c * axz; ub[7] = c * ayy; ub[8] = c * ayz; ub[9] = c * azz;
    }
    ub[1] = bx; ub[2] = by; ub[3] = bz;
  }
  ub[0] = u0;
}

// output layer: os[half*128 + j] = sum_{k in half} wl[k] * tile[k][j]  (256 threads, 2 per column)
template <int NV>
__device__ __forceinline__ void tc_output_dot(const unsigned char* tile, const float* wl_s, float* os, int tid) {
  const int j = tid & 127, half = tid >> 7;
  if (j < NV) {
    const uint32_t cj = (uint32_t)(j & 63) >> 3;
    const unsigned char* base = tile + (j >> 6) * 32768 + (j & 7) * 2 + half * 16 * 1024;
    float sum = 0.f;
#pragma unroll 4
    for (int kg = 0; kg < 16; ++kg) {
#pragma unroll
      for (uint32_t r = 0; r < 8; ++r) {
        const __half hv = *reinterpret_cast<const __half*>(base + kg * 1024 + r * 128 + ((cj ^ r) << 4));
        sum = fmaf(wl_s[half * 128 + kg * 8 + r], __half2float(hv), sum);
      }
    }
    os[half * 128 + j] = sum;
  }
}

// ---- thread-major stash of one column group: chunk j of the group lives at dst + j*1024 floats (+ 4*neuron) -------
// value channels (u0) in fp32 first, then the derivative channels packed as halves (in column order)
template <int NCH, int GC>
struct Stash {
  static constexpr int NP = GC / NCH;                      // points per group
  static constexpr int NF4 = NP / 4;                       // float4 chunks of u0
  static constexpr int ND = GC - NP;                       // derivative values
  static constexpr int NH8 = (ND + 7) / 8;                 // uint4 chunks of halves
  static constexpr int CHUNKS = (NCH == 1) ? GC / 4 : NF4 + NH8;
};

// `hint`: 0 plain stores, 1 st.global.cs (streaming), 2 st.global.wt (write-through)
__device__ __forceinline__ void tt_st16(float* dst, float4 v, int hint) {
  if (hint == 1) __stcs(reinterpret_cast<float4*>(dst), v);
  else if (hint == 2) __stwt(reinterpret_cast<float4*>(dst), v);
  else *reinterpret_cast<float4*>(dst) = v;
}
template <int NCH, int GC>
__device__ __forceinline__ void tt_stash_group(const float* u, float* dst, int hint = 0) {
  using S = Stash<NCH, GC>;
  if constexpr (NCH == 1) {
#pragma unroll
    for (int j4 = 0; j4 < GC / 4; ++j4) tt_st16(dst + j4 * 1024, make_float4(u[j4 * 4], u[j4 * 4 + 1], u[j4 * 4 + 2], u[j4 * 4 + 3]), hint);
  } else {
#pragma unroll
    for (int j4 = 0; j4 < S::NF4; ++j4)
      tt_st16(dst + j4 * 1024, make_float4(u[(j4 * 4) * NCH], u[(j4 * 4 + 1) * NCH], u[(j4 * 4 + 2) * NCH], u[(j4 * 4 + 3) * NCH]), hint);
    float d[S::NH8 * 8];
#pragma unroll
    for (int pp = 0; pp < S::NP; ++pp)
#pragma unroll
      for (int ch = 1; ch < NCH; ++ch) d[pp * (NCH - 1) + ch - 1] = u[pp * NCH + ch];
#pragma unroll
    for (int j = S::ND; j < S::NH8 * 8; ++j) d[j] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < S::NH8; ++c8)
      tt_st16(dst + (S::NF4 + c8) * 1024,
              make_float4(__uint_as_float(tc_pack_h2(d[8 * c8], d[8 * c8 + 1])), __uint_as_float(tc_pack_h2(d[8 * c8 + 2], d[8 * c8 + 3])),
                          __uint_as_float(tc_pack_h2(d[8 * c8 + 4], d[8 * c8 + 5])), __uint_as_float(tc_pack_h2(d[8 * c8 + 6], d[8 * c8 + 7]))),
              hint);
  }
}

// the reverse sweep issues the loads of the NEXT group before it works on the current one (raw 16-byte registers)
template <int NCH, int GC>
__device__ __forceinline__ void tt_stash_load(uint4* raw, const float* src) {
#pragma unroll
  for (int j = 0; j < Stash<NCH, GC>::CHUNKS; ++j) raw[j] = __ldcs(reinterpret_cast<const uint4*>(src + j * 1024));
}

template <int NCH, int GC>
__device__ __forceinline__ void tt_unstash_group(float* u, const uint4* raw) {
  using S = Stash<NCH, GC>;
  if constexpr (NCH == 1) {
#pragma unroll
    for (int j4 = 0; j4 < GC / 4; ++j4) {
      const uint4 t = raw[j4];
      u[j4 * 4] = __uint_as_float(t.x); u[j4 * 4 + 1] = __uint_as_float(t.y);
      u[j4 * 4 + 2] = __uint_as_float(t.z); u[j4 * 4 + 3] = __uint_as_float(t.w);
    }
  } else {
#pragma unroll
    for (int j4 = 0; j4 < S::NF4; ++j4) {
      const uint4 t = raw[j4];
      u[(j4 * 4) * NCH] = __uint_as_float(t.x); u[(j4 * 4 + 1) * NCH] = __uint_as_float(t.y);
      u[(j4 * 4 + 2) * NCH] = __uint_as_float(t.z); u[(j4 * 4 + 3) * NCH] = __uint_as_float(t.w);
    }
    float d[S::NH8 * 8];
#pragma unroll
    for (int c8 = 0; c8 < S::NH8; ++c8) {
      const uint4 t = raw[S::NF4 + c8];
      const __half2* hv = reinterpret_cast<const __half2*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f2 = __half22float2(hv[e]);
        d[8 * c8 + 2 * e] = f2.x;
        d[8 * c8 + 2 * e + 1] = f2.y;
      }
    }
#pragma unroll
    for (int pp = 0; pp < S::NP; ++pp)
#pragma unroll
      for (int ch = 1; ch < NCH; ++ch) u[pp * NCH + ch] = d[pp * (NCH - 1) + ch - 1];
  }
}

// one row segment of a training batch as the kernels see it
struct SegDev {
  const float* x;        // [P][3]
  float* outp;           // forward: [P][NCH] raw channels
  const float* seeds;    // backward: [P][NCH]
  const float* normals;  // fused step: [P][3] ground-truth normals
  const float* dist;     // fused step: [P] ground-truth distances
  int64_t P;
  int64_t npairs;
};

}  // namespace dudf

// Split-precision ("TCX3") forward chain on tcgen05: fp32-grade jets from fp16 tensor-core operands.
//
// Every hidden-layer product  (w W) a  is evaluated as three fp16 MMAs into one fp32 TMEM accumulator,
//     Wh*Ah + Wh*Al + Wl*Ah        with   X = Xh + Xl,  Xh = fp16(X),  Xl = fp16(X - Xh)   (the Wl*Al term is 2^-22 relative),
// which restores the 22+ significand bits the reference's fp32 nn.Linear works with (src/model.py:29-30,116-135); the
// single-pass fp16 chain (dudf_tc.cu) keeps 11.  tools/precision_study.py (operand roundings emulated on the CPU) puts the
// jets of this scheme at 1e-7..5e-7 of fp64 and — with the reverse sweep and the weight-gradient GEMM left single-pass —
// the parameter gradients of loss_s1 / loss_s2 / loss_siren at 5e-4..1e-3, inside north_star's 1e-3 for tensor-core paths.
//
// Same orientation as dudf_tc.cu (D[neuron][column], weights = UMMA A operand, activations = B operand, lane = neuron), but a
// different schedule, because the activation tile now exists twice (hi and lo, 2 x 64 KB for 128 columns): ONE 128-column
// tile per CTA, updated IN PLACE, with the k range of every layer split in two halves.  Per layer the MMA warp issues the groups
//     G(h=0,kh=0) G(1,0) | G(0,1) G(1,1)          (h = output-neuron half / accumulator, kh = input-neuron half = tile rows)
// so that after the first two groups the rows of k-half 0 are dead.  The epilogue of neuron half 0 (all 8 warps: 128 neurons x
// two column halves) starts when G(0,1) retires and overwrites exactly those rows with the next layer's activations while
// G(1,1) still runs; the epilogue of half 1 overlaps the next layer's G(.,0) groups, which only read rows of k-half 0.  The
// tensor pipe stays busy as long as an epilogue half fits under 1 536 clk (one group = 24 MMAs of 64 clk); the four
// accumulators (2 halves x 2 layer parities) fill the 512 TMEM columns.  Weights are scaled by 64 before the split so that
// their lo parts stay out of the fp16 subnormal range; the epilogue folds the 1/64 back in.
//
// The same kernel serves the training forward (TRAIN): it additionally stashes the pre-activations for the reverse sweep and
// lets the MMA warp bulk-copy the hi tile as the operand image of the weight-gradient GEMM — the layouts of dudf_tc_train.cu,
// whose (single-pass) reverse sweep and weight-gradient kernels consume them unchanged.
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"
#include "dudf_umma.cuh"
#include "dudf_tc_common.cuh"

namespace dudf {

using namespace umma;

constexpr int TCX_STAGES = 5;
constexpr float TCX_WSCALE = 64.f;
constexpr float TCX_WSCALE_INV = 1.f / 64.f;
constexpr int TCX_LAYER_CHUNKS = 16;           // 4 groups x 2 k-blocks x (hi, lo) chunks of 128 neurons x 64 k
constexpr int TCX_OFF_LO = TC_ACT_BYTES;       // lo tile behind the hi tile
constexpr int TCX_OFF_RING = 2 * TC_ACT_BYTES;
constexpr int TCX_OFF_WL = TCX_OFF_RING + TCX_STAGES * TC_CHUNK_BYTES;
constexpr int TCX_OFF_XS = TCX_OFF_WL + 256 * 4;
constexpr int TCX_OFF_OS = TCX_OFF_XS + 2 * 128 * 3 * 4;   // two point buffers: the next tile's points are loaded a tile ahead
constexpr int TCX_OFF_BAR = TCX_OFF_OS + 8 * 128 * 4;        // os: [8 = neuron half x lane quarter][128 columns] partial sums
constexpr int TCX_SMEM = TCX_OFF_BAR + 256 + 1024;
static_assert(TCX_SMEM <= 232448, "shared memory budget");

size_t tcx_packed_bytes(int n_lin) { return (size_t)((n_lin > 2) ? (n_lin - 2) : 1) * TCX_LAYER_CHUNKS * TC_CHUNK_BYTES; }

// fp16 hi / lo images of 64 w W in the order the MMA warp consumes them: [layer][kh][h][kb][hi | lo], 16 KB swizzled chunks
__global__ void __launch_bounds__(256) tcx_pack_kernel(NetView net, unsigned char* packed) {
  const int l = blockIdx.y + 1;
  const int n = blockIdx.x;
  const int k = threadIdx.x;
  const float v = TCX_WSCALE * (net.ww * net.W[l][n * 256 + k]);
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int h = n >> 7, r = n & 127, kh = k >> 7, kb = (k >> 6) & 1, kk = k & 63;
  unsigned char* chunk = packed + ((size_t)(l - 1) * TCX_LAYER_CHUNKS + ((kh * 2 + h) * 2 + kb) * 2) * TC_CHUNK_BYTES;
  *reinterpret_cast<__half*>(chunk + sw128_offset(r, kk)) = hi;
  *reinterpret_cast<__half*>(chunk + TC_CHUNK_BYTES + sw128_offset(r, kk)) = lo;
}

int tcx_pack(const NetView& net, void* packed, cudaStream_t st) {
  if (net.n_lin <= 2) return 0;
  tcx_pack_kernel<<<dim3(256, net.n_lin - 2), 256, 0, st>>>(net, (unsigned char*)packed);
  DUDF_LAUNCH_OK();
  return 0;
}

template <int NCH>
struct TcxCfg {
  using B = TcCfg<NCH>;
  static constexpr int PT = B::PT, NV = B::NV, GC = B::GC, NGRP = B::NGRP;
};
// EW epilogue warps (8 or 16): warp w serves TMEM lane quarter w & 3 and the column groups of set w >> 2.  With 16 warps every
// warp owns ONE column group per neuron half (jet-10 tiles have 3 groups: the fourth set idles), 4 warps per scheduler hide the
// TMEM-load / conversion / store latencies of each other, and a neuron half is turned around in ~half the time — the half
// epilogue sits on the critical path of the in-place schedule (tools/tcx_trace.py).
template <int EW> __device__ __forceinline__ void tcx_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory"); }
template <int EW> struct TcxRegs { static constexpr int EPI = (EW == 8) ? TC_REGS_EPI : (EW == 12 ? 152 : 104), AUX = TC_REGS_AUX; };   // 16 warps launch at 128: (512 * 128 - 128 * 56) / 384 = 152   // 20 warps launch at 96: (640 * 96 - 128 * 56) / 512 = 106

__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ---- producer: 16 chunks per layer, in consumption order ----
template <int CL>
__device__ __forceinline__ void tcx_producer(const unsigned char* packed, unsigned char* ring, uint64_t* full, uint64_t* empty, int64_t rounds,
                                             int n_phase, uint32_t cta_rank, int dbg) {
  uint32_t stage = 0, phase = 0, chunk = 0;
  for (int64_t r = 0; r < rounds; ++r)
    for (int j = 0; j < n_phase; ++j) {
      const unsigned char* src = packed + (size_t)j * TCX_LAYER_CHUNKS * TC_CHUNK_BYTES;
      for (int ck = 0; ck < TCX_LAYER_CHUNKS; ++ck, ++chunk) {
        mbar_wait_relaxed(&empty[stage], phase ^ 1, 0x1100 + stage);
        if ((dbg & 4) && chunk >= (uint32_t)TCX_STAGES) {          // diagnostics: stale weights, no L2 -> SM traffic
          mbar_arrive(&full[stage]);
          if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
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
        if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
      }
    }
}

// ---- MMA issuer (whole warp, warp-uniform, one elected lane per instruction; see tc_mma_role) ----
// img != null (training forward): each k-half of the hi tile is bulk-copied to the operand images of the weight-gradient GEMM
// as soon as it is published; its shared-memory reads are complete before the accumulator whose epilogue overwrites those rows
// is handed over.
template <int CL>
__device__ __forceinline__ void tcx_mma_role(unsigned char* act, unsigned char* ring, uint64_t* full, uint64_t* empty, uint64_t* act_ready,
                                             uint64_t* acc_ready, uint32_t tmem_base, int64_t rounds, int n_phase, int64_t ntiles,
                                             unsigned char* img, int64_t ncb, int64_t cb0, uint64_t img_policy, int dbg,
                                             unsigned long long* trace) {
  const bool skip = (dbg & 2) != 0;
  uint32_t tn = 0;
  constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, /*A K-major*/ 0, /*B MN-major*/ 1);
  constexpr uint16_t mask = (uint16_t)((1u << CL) - 1);
  const uint64_t a_desc0 = make_desc_sw128(smem_u32(ring), 16, 1024);
  const uint64_t bh_desc0 = make_desc_sw128(smem_u32(act), 32768, 1024);
  const uint64_t bl_desc0 = desc_advance(bh_desc0, TCX_OFF_LO);
  uint32_t stage = 0, phase = 0, act_phase = 0, jg = 0;
  auto release = [&](uint32_t st) {
    if constexpr (CL == 1) mma_commit_warp(&empty[st]);
    else mma_commit_multicast_warp(&empty[st], mask);
  };
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t tile = blockIdx.x + r * gridDim.x;
    const bool copy = img != nullptr && tile < ntiles && !(dbg & 16384);
    for (int j = 0; j < n_phase; ++j, ++jg) {
      const uint32_t acc = tmem_base + (jg & 1u) * 256;
      for (int kh = 0; kh < 2; ++kh) {
        mbar_wait(&act_ready[kh], (act_phase >> kh) & 1u, 0x1200 + kh);
        act_phase ^= 1u << kh;
        tc_fence_after();
        tc_trace(trace, tn, 1, kh);
        if (copy) {
          if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              unsigned char* dst = img + ((size_t)j * ncb + cb0 + tile * 2 + nb) * TC_IMG_BYTES + kh * (TC_IMG_BYTES / 2);
              const unsigned char* src = act + nb * TC_IMG_BYTES + kh * (TC_IMG_BYTES / 2);
              if (img_policy) bulk_s2g_hint(dst, src, TC_IMG_BYTES / 2, img_policy);
              else bulk_s2g(dst, src, TC_IMG_BYTES / 2);
            }
            bulk_commit();
          }
          __syncwarp();
        }
        for (int h = 0; h < 2; ++h) {
          const uint32_t d_tmem = acc + h * 128;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t koff = (uint32_t)(kh * 2 + kb) * 8192u;
            mbar_wait(&full[stage], phase, 0x1300 + stage);                    // hi chunk: against the hi and the lo tile
            const uint64_t a_hi = desc_advance(a_desc0, stage * TC_CHUNK_BYTES);
            if (!skip) {
              // Wh x (Ah, Al): pairwise per K = 16 step with the weight slice latched in the collector, or as two runs of four steps
              // that re-read it.  Measured (profiles/r3_tcx_ab_ldtm.txt): queries are 2 % (f, grad) to 10 % (value only) FASTER
              // without the reuse, the training forward 0.8 % slower — so each mode takes its own; DUDF_TCX_DBG bit 16 swaps them
              if (((dbg & 16) != 0) == (img != nullptr)) {
                mma_f16_ss_k64_warp(d_tmem, a_hi, desc_advance(bh_desc0, koff), idesc, (kh | kb) != 0);
                mma_f16_ss_k64_warp(d_tmem, a_hi, desc_advance(bl_desc0, koff), idesc, 1);
              } else {
                mma_f16_ss_k64_pair_warp(d_tmem, a_hi, desc_advance(bh_desc0, koff), desc_advance(bl_desc0, koff), idesc, (kh | kb) != 0);
              }
            }
            release(stage);
            if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
            mbar_wait(&full[stage], phase, 0x1300 + stage);                    // lo chunk: against the hi tile
            if (!skip) mma_f16_ss_k64_warp(d_tmem, desc_advance(a_desc0, stage * TC_CHUNK_BYTES), desc_advance(bh_desc0, koff), idesc, 1);
            release(stage);
            if (++stage == TCX_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_trace(trace, tn, 2, kh * 2 + h);
          if (kh == 1) {
            if (copy) {
              if ((threadIdx.x & 31) == 0) {
                if (h == 0) bulk_wait_read1();        // the copy of k-half 0 has left shared memory
                else bulk_wait_read0();
              }
              __syncwarp();
            }
            mma_commit_warp(&acc_ready[h]);
          }
        }
      }
    }
  }
}

// hi / lo split of GC fp32 values of one thread into the two B tiles (saturating: an overflowing activation clamps instead of inf)
__device__ __forceinline__ uint32_t tcx_pack_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <int GC>
__device__ __forceinline__ void tcx_store_group(const float* v, uint32_t row_hi, int chunk0, uint32_t r7) {
#pragma unroll
  for (int c8 = 0; c8 < GC / 8; ++c8) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = v[8 * c8 + 2 * e], b = v[8 * c8 + 2 * e + 1];
      hi[e] = tcx_pack_sat(a, b);
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi[e]));
      lo[e] = tcx_pack_sat(a - hf.x, b - hf.y);
    }
    const uint32_t off = tc_chunk_off(chunk0 + c8, r7);
    tc_sts128(row_hi + off, hi[0], hi[1], hi[2], hi[3]);
    tc_sts128(row_hi + TCX_OFF_LO + off, lo[0], lo[1], lo[2], lo[3]);
  }
}

// column sums over the 32 lanes of a warp: v[0..M-1] of every lane -> v[0] of lane L = sum over lanes of column (L % M)
template <int M>
__device__ __forceinline__ void tcx_colsum(float* v, int lane) {
#pragma unroll
  for (int o = 16; o >= M; o >>= 1) {
#pragma unroll
    for (int i = 0; i < M; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
#pragma unroll
  for (int o = M / 2; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = up ? v[i + o] : v[i];
      const float send = up ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

struct TcxEpi {
  unsigned char* act;
  float* xs;
  float* os;
  const float* wl_s;
  uint64_t* act_ready;
  uint64_t* acc_ready;
  uint32_t tmem_q;        // accumulator base of this warp's lane quarter
  uint32_t acc_phase;     // bit h = parity of acc_ready[h]
  uint32_t jg;            // layer phases completed so far (accumulator parity)
  int q, cw, lane, tid;
  unsigned long long* trace;   // diagnostics (tools/tcx_trace.py): event log of CTA 0 / warp 0, or null
  uint32_t tn;
};

struct TcxTrain {         // training forward: stash + operand images (null for queries)
  float* Ust;
  unsigned char* Aimg;
  int64_t ld, col0;
  int dbg;                // diagnostics (DUDF_TCX_DBG; results are meaningless): 1 no epilogue math, 2 no MMAs, 4 no weight traffic,
                          //   8 no tile stores, 16 swap the collector-reuse choice (training: reuse, queries: none), 32 no accumulator loads,
                          //   256 / 512 / 1024 lane quarter 1 / 0 / 2 idle (see tcx_forward_kernel), 2048 accumulator loads group by group, 8192 no stash stores, 16384 no operand-image copies, 32768 stash into one 128 KB region per CTA (stays in the L2), 65536 plain instead of streaming stash stores
  unsigned long long* trace;
  const float* dirs;      // DIR3 queries: [P][9] = three unit directions (a, b, c) per point
};

// ---- directional third-order jet (mean curvature on tensor cores, src/render_st.py:42-62; SURVEY.md §8 a-M) -------------------
// Ten channels per point for three per-point directions a, b, c (a, b = the two tangent eigen-directions, c = the eigen-normal):
//   0: f   1-3: D_a, D_b, D_c   4-7: D_aa, D_ac, D_bb, D_bc (x KAPPA)   8-9: D_aac, D_bbc (x KAPPA3)
// — everything the chain rule of sin needs to carry D^3 f[a, a, c] and D^3 f[b, b, c] through the layers, and exactly what
// tr(dn/dx) = sum_j T(v_j, v_j, n) / (lambda_2 - lambda_j) consumes: the same 128-column tile as the Hessian jet instead of the
// 20-channel full third-order jet on the CUDA cores.
constexpr float TCX_KAPPA3 = 1.f / 256.f, TCX_KAPPA3_INV = 256.f;
template <int GC>
__device__ __forceinline__ void tcx_first_layer_dir(float* u, const float* pts, float w0, float r0x, float r0y, float r0z, float b0) {
#pragma unroll
  for (int pp = 0; pp < GC / 10; ++pp) {
    const float* q = pts + pp * 12;
    float* up = u + pp * 10;
    up[0] = w0 * fmaf(r0z, q[2], fmaf(r0y, q[1], fmaf(r0x, q[0], b0)));
#pragma unroll
    for (int d = 0; d < 3; ++d) up[1 + d] = w0 * fmaf(r0z, q[5 + 3 * d], fmaf(r0y, q[4 + 3 * d], r0x * q[3 + 3 * d]));
#pragma unroll
    for (int ch = 4; ch < 10; ++ch) up[ch] = 0.f;
  }
}
// stored pre-activations (derivative channels still times 1 / inv) -> stored activations, in place; u[0] is not read
__device__ __forceinline__ void tcx_act_point_dir(float* u, float s, float c, float inv) {
  const float ci = c * inv;
  const float za = inv * u[1], zb = inv * u[2], zc = inv * u[3];
  const float vaa = inv * u[4], vac = inv * u[5], vbb = inv * u[6], vbc = inv * u[7];
  constexpr float K32 = TCX_KAPPA3 / TC_KAPPA;
  const float s32 = K32 * s, c3 = TCX_KAPPA3 * c, ks = TC_KAPPA * s;
  u[8] = fmaf(ci, u[8], -fmaf(s32, fmaf(vaa, zc, 2.f * vac * za), c3 * za * za * zc));
  u[9] = fmaf(ci, u[9], -fmaf(s32, fmaf(vbb, zc, 2.f * vbc * zb), c3 * zb * zb * zc));
  u[4] = fmaf(c, vaa, -ks * za * za);
  u[5] = fmaf(c, vac, -ks * za * zc);
  u[6] = fmaf(c, vbb, -ks * zb * zb);
  u[7] = fmaf(c, vbc, -ks * zb * zc);
  u[1] = c * za;
  u[2] = c * zb;
  u[3] = c * zc;
  u[0] = s;
}

// ---- one 128-column tile, in pieces ------------------------------------------------------------------------------------------
// Per-tile inputs of the epilogue warps.  x == null: grid points (grid_first + p).
struct TcxTileArgs {
  const float* x;
  const float* dirs;      // DIR3 queries
  int64_t P, tile, colt, grid_first;
  int gridN;
  float vs;
  bool valid;
  float* outp;            // TRAIN: raw channels [P][NCH]
};
enum { TCX_FIRST = 0, TCX_HIDDEN = 1, TCX_LAST = 2 };

// the tile's points (+ directions) -> xs; cooperative over the EW epilogue warps, the caller orders a tcx_epi_bar before they are read
template <int NCH, int EW, bool DIR>
__device__ __forceinline__ void tcx_load_points(const TcxEpi& e, const TcxTileArgs& t, float* xs) {
  constexpr int XS = DIR ? 12 : 3;
  using C = TcxCfg<NCH>;
  for (int i = e.tid; i < C::PT; i += EW * 32) {
    const int64_t p = t.tile * C::PT + i;
    float pt[3] = {0.f, 0.f, 0.f};
    if (t.valid && p < t.P) {
      if (t.x) { pt[0] = t.x[p * 3]; pt[1] = t.x[p * 3 + 1]; pt[2] = t.x[p * 3 + 2]; }
      else grid_point(t.grid_first + p, t.gridN, t.vs, pt);
    }
    xs[i * XS] = pt[0]; xs[i * XS + 1] = pt[1]; xs[i * XS + 2] = pt[2];
    if constexpr (DIR) {
#pragma unroll
      for (int k = 0; k < 9; ++k) xs[i * XS + 3 + k] = (t.valid && p < t.P) ? t.dirs[p * 9 + k] : 0.f;
    }
  }
}

// One neuron half h of one layer l of tile t.  KIND: TCX_FIRST (l = 0: from the points in xs, no accumulator), TCX_HIDDEN, TCX_LAST
// (l = L - 1: output layer into e.os).  The caller has waited for the accumulator (l > 0).  FIRST / HIDDEN publish the half to the MMA warp.
template <int NCH, bool TRAIN, int SC, int EW, bool DIR, int KIND>
__device__ __forceinline__ void tcx_half(TcxEpi& e, const NetView& net, const TcxTileArgs& t, const float* xs, int l, int h, float* Ust,
                                         int64_t ld, int dbg) {
  constexpr int XS = DIR ? 12 : 3;
  using C = TcxCfg<NCH>;
  constexpr int GC = C::GC;
  const float w0 = net.w0, ww = net.ww;
  constexpr int NG_PER = (C::NGRP + EW / 4 - 1) / (EW / 4);
  const int g_begin = min(e.cw * NG_PER, C::NGRP), g_end = min(g_begin + NG_PER, C::NGRP);
  const int n = h * 128 + e.q * 32 + e.lane;               // this thread's neuron in this half = its row of the B operand
  const uint32_t r7 = n & 7;
  const uint32_t row_hi = smem_u32(tc_tile_row(e.act, n));
  float r0x = 0.f, r0y = 0.f, r0z = 0.f, b0 = 0.f, bias = 0.f;
  if constexpr (KIND == TCX_FIRST) {
    r0x = net.W[0][n * 3]; r0y = net.W[0][n * 3 + 1]; r0z = net.W[0][n * 3 + 2]; b0 = net.b[0][n];
    // the idle columns of a tile narrower than 128 are zero for its MMAs (a launch may mix jet orders): every thread clears its own row
    if (C::NV < 128 && e.cw == 0) {
      tc_sts128(row_hi + tc_chunk_off(15, r7), 0u, 0u, 0u, 0u);
      tc_sts128(row_hi + TCX_OFF_LO + tc_chunk_off(15, r7), 0u, 0u, 0u, 0u);
    }
  } else {
    bias = ww * net.b[l][n];
  }
  const float wl = (KIND == TCX_LAST) ? e.wl_s[n] : 0.f;
  tc_trace(e.trace, e.tn, 10 + h, l);
  const uint32_t taddr = e.tmem_q + ((e.jg + (uint32_t)l - 1u) & 1u) * 256 + h * 128;
  auto group = [&](int g, const TmemRegs<GC>* pre) {
    float u[GC];
    if constexpr (KIND == TCX_FIRST) {
      if constexpr (DIR) tcx_first_layer_dir<GC>(u, xs + g * (GC / NCH) * XS, w0, r0x, r0y, r0z, b0);
      else tc_first_layer_group<NCH, GC>(u, xs + g * (GC / NCH) * 3, w0, r0x, r0y, r0z, b0);
    } else {
      if (dbg & 32) {            // diagnostics: no accumulator loads
#pragma unroll
        for (int j = 0; j < GC; ++j) u[j] = (float)(j + e.lane);
      } else if (pre) {
        tc_ld_take<GC>(*pre, u);
      } else {
        TmemRegs<GC> tr;
        tc_ld_issue<GC>(taddr + g * GC, tr);
        tc_ld_take<GC>(tr, u);
      }
      if constexpr (TRAIN) {       // the stash holds the unscaled pre-activations of every channel
#pragma unroll
        for (int j = 0; j < GC; ++j) u[j] *= TCX_WSCALE_INV;
#pragma unroll
        for (int pp = 0; pp < GC / NCH; ++pp) u[pp * NCH] += bias;
      } else {                     // queries: only the sine argument is unscaled, the derivative channels keep the 64 (below)
#pragma unroll
        for (int pp = 0; pp < GC / NCH; ++pp) u[pp * NCH] = fmaf(u[pp * NCH], TCX_WSCALE_INV, bias);
      }
    }
    if constexpr (TRAIN && KIND != TCX_FIRST) {      // (the reverse sweep recomputes the first layer from the points: no stash for it)
      if (t.valid && !(dbg & 8192))
        tt_stash_group<NCH, GC>(u, Ust + ((dbg & 32768) ? (size_t)blockIdx.x * 128 : (size_t)l * ld + t.colt) * 256 + n * 4 + (size_t)g * GC * 256,
                                (dbg & 65536) ? 0 : 1);      // streaming stores (st.global.cs): -2.6 % on the launch against plain ones (bit 65536)
    }
    if (!(dbg & 1)) {
#pragma unroll
      for (int pp = 0; pp < GC / NCH; ++pp) {
        float sn, cs;
        if constexpr (SC == 1) sincos_poly(u[pp * NCH], sn, cs);
        else sincos_fast(u[pp * NCH], sn, cs);
        if constexpr (TRAIN) tc_act_point<NCH>(u + pp * NCH, sn, cs);
        else if constexpr (DIR) tcx_act_point_dir(u + pp * NCH, sn, cs, KIND == TCX_FIRST ? 1.f : TCX_WSCALE_INV);
        else tc_act_point_scaled<NCH>(u + pp * NCH, sn, cs, KIND == TCX_FIRST ? 1.f : TCX_WSCALE_INV);
      }
    }
    if constexpr (KIND != TCX_LAST) {
      if (!(dbg & 8)) tcx_store_group<GC>(u, row_hi, g * (GC / 8), r7);
    } else {
      // output layer (256 -> 1 per channel) on the fp32 activations: this warp's 32 neurons, reduced over its lanes
#pragma unroll
      for (int j = 0; j < GC; ++j) u[j] *= wl;
      float* dst = e.os + (h * 4 + e.q) * 128 + g * GC;
      tcx_colsum<32>(u, e.lane);
      dst[e.lane] = u[0];
      if constexpr (GC == 40) {
        tcx_colsum<8>(u + 32, e.lane);
        if (e.lane < 8) dst[32 + e.lane] = u[32];
      }
    }
  };
  // A TMEM load that competes with the accumulating MMAs of the other neuron half takes ~850 clk (tools/tcx_trace.py), three
  // times the ~250 instructions of a group: with two groups per warp, both loads are issued up front so that their latencies
  // overlap instead of adding up (two distinct register sets, no rotation: ptxas keeps them in place).  Half epilogue 2 250 -> 1 830
  // clk, the MMA groups under it 2 020 -> 2 230 (TMEM reads and accumulation share the port), layer period 9 310 -> 8 820 clk:
  // grid queries +4 % (profiles/r3_tcx_trace_ldtm_*.txt, r3_tcx_ab_ldtm.txt).  40-column groups (Hessian jet) LOSE 4-5 % with either
  // form of look-ahead (both loads up front, or the second issued when the first has landed): only one of the two warp sets has a
  // second group, and its earlier loads take TMEM cycles from the MMAs of the other neuron half.
  if (KIND != TCX_FIRST && GC == 32 && NG_PER == 2 && g_end - g_begin == 2 && !(dbg & (32 | 2048))) {
    TmemRegs<GC> t0, t1;
    tmem_ld_x64(taddr + g_begin * GC, t0.a, t1.a);       // one 64-column load (two x32 loads time the same)
    group(g_begin, &t0);
    group(g_begin + 1, &t1);
  } else {
#pragma unroll 1
    for (int g = g_begin; g < g_end; ++g) group(g, nullptr);
  }
  tc_trace(e.trace, e.tn, 12 + h, l);
  if constexpr (KIND != TCX_LAST) {
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (e.lane == 0) mbar_arrive(&e.act_ready[h]);
  }
  tc_trace(e.trace, e.tn, 30 + h, l);
}

__device__ __forceinline__ void tcx_wait_acc(TcxEpi& e, int l, int h) {
  tc_trace(e.trace, e.tn, 40 + h, l);
  mbar_wait(&e.acc_ready[h], (e.acc_phase >> h) & 1u, 0x1400 + h);
  e.acc_phase ^= 1u << h;
  tc_fence_after();
}

// partial sums of the output layer -> channels -> outputs of the tile's points
template <int NCH, bool TRAIN, int EW, bool DIR>
__device__ __forceinline__ void tcx_finish_tile(TcxEpi& e, const NetView& net, const TcxTileArgs& t, const QueryOut& out) {
  using C = TcxCfg<NCH>;
  const int L = net.n_lin - 1;
  tc_trace(e.trace, e.tn, 20, 0);
  tcx_epi_bar<EW>();
  if (e.tid < C::NV) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += e.os[k * 128 + e.tid];
    const int ch = e.tid % NCH;
    if constexpr (DIR) e.os[e.tid] = (ch == 0) ? v + net.b[L][0] : (ch >= 8 ? v * TCX_KAPPA3_INV : (ch >= 4 ? v * TC_KAPPA_INV : v));
    else e.os[e.tid] = (ch == 0) ? v + net.b[L][0] : (ch >= 4 ? v * TC_KAPPA_INV : v);
  }
  tcx_epi_bar<EW>();
  if (e.tid < C::PT) {
    const int64_t p = t.tile * C::PT + e.tid;
    if (t.valid && p < t.P) {
      if constexpr (TRAIN) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) t.outp[p * NCH + ch] = e.os[e.tid * NCH + ch];
      } else {
        finalize_point<NCH>(out, p, e.os + e.tid * NCH);
      }
    }
  }
  tc_trace(e.trace, e.tn, 15, 0);
}

// One launch serves up to two row segments with different jet orders (training: on-surface rows carry the Hessian jet).
// Queries use segment a only (x == null: grid points).
template <int NA, int NB, int CL, bool TRAIN, int SC, int EW, bool DIR = false>
__global__ void __launch_bounds__((EW + 4) * 32, 1)
tcx_forward_kernel(const unsigned char* __restrict__ packed, NetView net, SegDev sa, SegDev sb, int64_t tiles_a, int64_t tiles_b, int gridN,
                   int64_t grid_first, QueryOut out, TcxTrain tr) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + TCX_OFF_RING;
  float* wl_s = (float*)(smem + TCX_OFF_WL);
  uint64_t* bars = (uint64_t*)(smem + TCX_OFF_BAR);
  uint64_t *full = bars, *empty = bars + TCX_STAGES, *act_ready = bars + 2 * TCX_STAGES, *acc_ready = bars + 2 * TCX_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TCX_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t ntiles = tiles_a + tiles_b;
  // every CTA of a cluster runs the same number of rounds (a round past the end works on a fully masked tile)
  const int64_t rounds = (CL > 1) ? (ntiles + gridDim.x - 1) / gridDim.x
                                  : (((int64_t)blockIdx.x < ntiles) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
  if (tid == 0) {
    for (int i = 0; i < TCX_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], EW); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == EW + 1) tmem_alloc<512>(tmem_slot);
  if (tid < 256) wl_s[tid] = net.W[L][tid];
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= EW) {
    setmaxnreg_dec<TcxRegs<EW>::AUX>();
    if (warp == EW) {
      if (lane == 0) tcx_producer<CL>(packed, ring, full, empty, rounds, L - 1, CL > 1 ? cluster_ctarank() : 0u, tr.dbg);
    } else if (warp == EW + 1) {
      tcx_mma_role<CL>(act, ring, full, empty, act_ready, acc_ready, tmem_base, rounds, L - 1, ntiles, TRAIN ? tr.Aimg : nullptr,
                       tr.ld >> 6, tr.col0 >> 6, TRAIN ? l2_policy_evict_first() : 0ull, tr.dbg,
                       (tr.trace && blockIdx.x == 0) ? tr.trace : nullptr);
    }
  } else {
    setmaxnreg_inc<TcxRegs<EW>::EPI>();
    TcxEpi e;
    e.act = act; e.xs = nullptr; e.os = (float*)(smem + TCX_OFF_OS); e.wl_s = wl_s;
    e.act_ready = act_ready; e.acc_ready = acc_ready;
    e.q = warp & 3; e.cw = warp >> 2; e.lane = lane; e.tid = tid;
    e.tmem_q = tmem_base + ((uint32_t)(e.q * 32) << 16);
    e.acc_phase = 0; e.jg = 0;
    e.trace = (tr.trace && blockIdx.x == 0 && warp == 0) ? tr.trace + TC_TRACE_REGION : nullptr;
    e.tn = 0;
    const float vs = gridN > 1 ? 2.0f / (float)(gridN - 1) : 0.f;
    float* xs2[2] = {(float*)(smem + TCX_OFF_XS), (float*)(smem + TCX_OFF_XS) + 128 * 3};
    auto args_of = [&](int64_t rd) {
      TcxTileArgs t;
      const int64_t tile = blockIdx.x + rd * gridDim.x;
      t.valid = tile < ntiles;
      const bool in_a = !t.valid || tile < tiles_a;          // a round past the end works on a fully masked tile of segment a
      const SegDev& sg = in_a ? sa : sb;
      t.x = sg.x; t.P = sg.P; t.outp = sg.outp; t.dirs = tr.dirs;
      t.tile = in_a ? tile : tile - tiles_a;
      t.colt = tr.col0 + tile * 128;
      t.gridN = gridN; t.grid_first = grid_first; t.vs = vs;
      return t;
    };
    auto is_a = [&](int64_t rd) { const int64_t tile = blockIdx.x + rd * gridDim.x; return NB == 0 || tile >= ntiles || tile < tiles_a; };
    // diagnostics: 256 / 512 / 1024 = the epilogue warps of lane quarter 1 / 0 / 2 (the SM partition of the MMA warp / of the
    // producer warp / of an idle warp) do no loads, math or stores — does the MMA stream slow down through its own partition?
    const int dbg = tr.dbg | ((((tr.dbg & 256) && e.q == 1) || ((tr.dbg & 512) && e.q == 0) || ((tr.dbg & 1024) && e.q == 2)) ? 41 : 0);
    constexpr int NBX = NB > 0 ? NB : NA;
    auto load_points = [&](int64_t rd, float* xs) {
      const TcxTileArgs t = args_of(rd);
      if (is_a(rd)) tcx_load_points<NA, EW, DIR>(e, t, xs);
      else tcx_load_points<NBX, EW, false>(e, t, xs);
    };
    auto first_half = [&](int64_t rd, int h) {
      const TcxTileArgs t = args_of(rd);
      if (is_a(rd)) tcx_half<NA, TRAIN, SC, EW, DIR, TCX_FIRST>(e, net, t, xs2[rd & 1], 0, h, tr.Ust, tr.ld, dbg);
      else tcx_half<NBX, TRAIN, SC, EW, false, TCX_FIRST>(e, net, t, xs2[rd & 1], 0, h, tr.Ust, tr.ld, dbg);
    };
    // Tiles are software-pipelined across their boundary: the single in-place tile leaves the tensor pipe idle from the last MMA of a
    // tile until the first layer of the next one is written, and the output layer + finalisation + point loads + first layer used to
    // sit in that gap (7 500 of 69 000 clk per tile, tools/tcx_trace.py).  Now the next tile's points are loaded a tile ahead, and its
    // first layer is written as soon as the last accumulator half of the current tile is complete (then the rows of that k-half are
    // dead) — BEFORE the current tile's output layer is evaluated, which then runs under the next tile's MMAs.
    if (rounds > 0) {
      load_points(0, xs2[0]);
      tcx_epi_bar<EW>();
      first_half(0, 0);
      first_half(0, 1);
    }
    for (int64_t rd = 0; rd < rounds; ++rd) {
      const TcxTileArgs t = args_of(rd);
      const bool a = is_a(rd), more = rd + 1 < rounds;
      tc_trace(e.trace, e.tn, 14, a ? NA : NBX);
      if (more) load_points(rd + 1, xs2[(rd + 1) & 1]);
      for (int l = 1; l < L - 1; ++l)
        for (int h = 0; h < 2; ++h) {
          tcx_wait_acc(e, l, h);
          if (a) tcx_half<NA, TRAIN, SC, EW, DIR, TCX_HIDDEN>(e, net, t, nullptr, l, h, tr.Ust, tr.ld, dbg);
          else tcx_half<NBX, TRAIN, SC, EW, false, TCX_HIDDEN>(e, net, t, nullptr, l, h, tr.Ust, tr.ld, dbg);
        }
      tcx_epi_bar<EW>();                                      // the next tile's points are visible to every warp; the previous tile's outputs have been read
      for (int h = 0; h < 2; ++h) {
        tcx_wait_acc(e, L - 1, h);                            // every MMA that read the rows of k-half h has completed
        if (more) first_half(rd + 1, h);
        if (a) tcx_half<NA, TRAIN, SC, EW, DIR, TCX_LAST>(e, net, t, nullptr, L - 1, h, tr.Ust, tr.ld, dbg);
        else tcx_half<NBX, TRAIN, SC, EW, false, TCX_LAST>(e, net, t, nullptr, L - 1, h, tr.Ust, tr.ld, dbg);
      }
      e.jg += (uint32_t)(L - 1);
      if (a) tcx_finish_tile<NA, TRAIN, EW, DIR>(e, net, t, out);
      else tcx_finish_tile<NBX, TRAIN, EW, false>(e, net, t, out);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();                 // nobody exits while a peer may still multicast into its ring
  if (warp == EW + 1) tmem_dealloc<512>(tmem_base);
}

static int tcx_cluster_size() {
  static int v = -1;            // DUDF_TCX_CLUSTER=1|2 overrides the cluster size (default 2: halves the L2 reads of the weight stream)
  if (v < 0) {
    const char* e = getenv("DUDF_TCX_CLUSTER");
    const int c = e ? atoi(e) : 2;
    v = (c == 1 || c == 2) ? c : 2;
  }
  return v;
}

template <int NA, int NB, int CL, bool TRAIN, int SC, int EW, bool DIR = false>
static int tcx_launch(const void* packed, const NetView& net, const SegDev& a, const SegDev& b, int64_t ta, int64_t tb, int gridN, int64_t first,
                      const QueryOut& out, const TcxTrain& tr, int sms, cudaStream_t st) {
  auto k = tcx_forward_kernel<NA, NB, CL, TRAIN, SC, EW, DIR>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TCX_SMEM));
  const int64_t ntiles = ta + tb;
  int grid = (int)std::min<int64_t>(ntiles, sms);
  if (grid < 1) return 0;
  if (CL > 1) grid = std::max(CL, (std::min<int>(sms, (int)((ntiles + CL - 1) / CL * CL)) / CL) * CL);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((EW + 4) * 32);
  cfg.dynamicSmemBytes = TCX_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DUDF_CUDA_OK(cudaLaunchKernelEx(&cfg, k, (const unsigned char*)packed, net, a, b, ta, tb, gridN, first, out, tr));
  DUDF_LAUNCH_OK();
  return 0;
}

// sine / cosine of the epilogue: 0 (default) = MUFU.SIN / MUFU.COS after an explicit 2 pi Cody-Waite reduction (absolute error 4e-7;
// jets f 1-3e-6, grad / Hessian 0.9-1.3e-5 of max, MeshUDF topology at 128^3 identical to the oracle's — tests/test_gpu_tcx3.py,
// tests/test_mc_topology.py pass in both modes); 1 = FMA-only minimax polynomials (DUDF_TCX_SINCOS=poly, 1e-7, ~190 instructions
// per 8-point group more: 8 % slower on grid queries, 25 % on value-only queries, profiles/r2_tcx_ab_sincos_burst.txt) for the
// same measured jet error — what remains is the truncating fp32 accumulation of the tensor pipe, not the sine
static int tcx_sincos_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DUDF_TCX_SINCOS");
    v = (e && strcmp(e, "poly") == 0) ? 1 : 0;
  }
  return v;
}
static int tcx_dbg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DUDF_TCX_DBG"); v = e ? atoi(e) : 0; }
  return v;
}

// epilogue warps per CTA: 8.  The 16-warp variant (4 warps per scheduler, one column group per warp and neuron half) is kept behind
// -DDUDF_TCX_EW16 + DUDF_TCX_EW=16: parity green, but the half epilogue is instruction-issue bound (~390 instructions per 32-column
// group, the same total on 8 or 16 warps: 2 200 clk either way, tools/tcx_trace.py), so it buys 3 % on grid queries and costs 2 % on
// the training forward (profiles/r2_tcx_ab.txt)
static int tcx_epi_warps() {
#ifdef DUDF_TCX_EW16
  static int v = -1;
  if (v < 0) { const char* e = getenv("DUDF_TCX_EW"); v = (e && atoi(e) == 16) ? 16 : 8; }
  return v;
#else
  return 8;
#endif
}

template <int NA, int NB, bool TRAIN, int SC, int EW>
static int tcx_launch_ew(int cl, const void* packed, const NetView& net, const SegDev& a, const SegDev& b, int64_t ta, int64_t tb, int gridN,
                         int64_t first, const QueryOut& out, const TcxTrain& tr, int sms, cudaStream_t st) {
  if (cl >= 2) return tcx_launch<NA, NB, 2, TRAIN, SC, EW>(packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
  return tcx_launch<NA, NB, 1, TRAIN, SC, EW>(packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
}

template <int NA, int NB, bool TRAIN>
static int tcx_launch_cl(const void* packed, const NetView& net, const SegDev& a, const SegDev& b, int64_t ta, int64_t tb, int gridN, int64_t first,
                         const QueryOut& out, const TcxTrain& tr0, int sms, cudaStream_t st) {
  int cl = tcx_cluster_size();
  while (cl > 1 && ta + tb < 2 * cl) cl >>= 1;
  TcxTrain tr = tr0;
  tr.dbg = tcx_dbg();
  tr.trace = tc_get_trace();
#ifdef DUDF_TCX_EW16
  if (tcx_epi_warps() == 16) {
    if (tcx_sincos_mode() == 0) return tcx_launch_ew<NA, NB, TRAIN, 0, 16>(cl, packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
    return tcx_launch_ew<NA, NB, TRAIN, 1, 16>(cl, packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
  }
#endif
  // (12 epilogue warps — three sets, one 40-column group per warp and neuron half, 152 registers — were tried for the training forward:
  //  parity green, 0.553-0.556 ms against 0.541: like the 16-warp variant, a shorter half epilogue only crowds the TMEM port)
  if (tcx_sincos_mode() == 0) return tcx_launch_ew<NA, NB, TRAIN, 0, 8>(cl, packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
  return tcx_launch_ew<NA, NB, TRAIN, 1, 8>(cl, packed, net, a, b, ta, tb, gridN, first, out, tr, sms, st);
}

int tcx_forward(const void* packed, const NetView& net, int nch, const float* x, int64_t P, int gridN, int64_t grid_first, const QueryOut& out,
                int sms, cudaStream_t st) {
  SegDev a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  a.x = x;
  a.P = P;
  TcxTrain tr;
  memset(&tr, 0, sizeof(tr));
  DUDF_REQUIRE(net.n_lin >= 3, "split tensor-core path: at least two hidden layers");
  switch (nch) {
    case 1: return tcx_launch_cl<1, 0, false>(packed, net, a, b, (P + TcxCfg<1>::PT - 1) / TcxCfg<1>::PT, 0, gridN, grid_first, out, tr, sms, st);
    case 4: return tcx_launch_cl<4, 0, false>(packed, net, a, b, (P + TcxCfg<4>::PT - 1) / TcxCfg<4>::PT, 0, gridN, grid_first, out, tr, sms, st);
    case 10: return tcx_launch_cl<10, 0, false>(packed, net, a, b, (P + TcxCfg<10>::PT - 1) / TcxCfg<10>::PT, 0, gridN, grid_first, out, tr, sms, st);
  }
  DUDF_REQUIRE(false, "split tensor-core path: unsupported channel count %d", nch);
}

// directional third-order jet of P points: packed [P][10] raw channels (see tcx_first_layer_dir)
int tcx_forward_dir3(const void* packed, const NetView& net, const float* x, const float* dirs, int64_t P, float* out_packed, int sms,
                     cudaStream_t st) {
  SegDev a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  a.x = x;
  a.P = P;
  TcxTrain tr;
  memset(&tr, 0, sizeof(tr));
  tr.dbg = tcx_dbg();
  tr.dirs = dirs;
  QueryOut out;
  memset(&out, 0, sizeof(out));
  out.packed = out_packed;
  DUDF_REQUIRE(net.n_lin >= 3, "split tensor-core path: at least two hidden layers");
  const int64_t tiles = (P + TcxCfg<10>::PT - 1) / TcxCfg<10>::PT;
  int cl = tcx_cluster_size();
  while (cl > 1 && tiles < 2 * cl) cl >>= 1;
  if (tcx_sincos_mode() == 0) {
    if (cl >= 2) return tcx_launch<10, 0, 2, false, 0, 8, true>(packed, net, a, b, tiles, 0, 0, 0, out, tr, sms, st);
    return tcx_launch<10, 0, 1, false, 0, 8, true>(packed, net, a, b, tiles, 0, 0, 0, out, tr, sms, st);
  }
  if (cl >= 2) return tcx_launch<10, 0, 2, false, 1, 8, true>(packed, net, a, b, tiles, 0, 0, 0, out, tr, sms, st);
  return tcx_launch<10, 0, 1, false, 1, 8, true>(packed, net, a, b, tiles, 0, 0, 0, out, tr, sms, st);
}

// training forward: same stash / operand-image layout and column allocation (256 columns per sub-tile pair) as tc_train_forward
int tcx_train_forward(const void* packed, const NetView& net, const TcSegment* segs, int nseg, float* Ust, void* Aimg, int64_t ld, int64_t col0,
                      int sms, cudaStream_t st) {
  DUDF_REQUIRE(ld % 64 == 0 && col0 % 128 == 0, "tensor-core stash: ld must be a multiple of 64 and col0 of 128");
  DUDF_REQUIRE(nseg == 1 || nseg == 2, "tensor-core training: 1 or 2 segments per launch");
  DUDF_REQUIRE(net.n_lin >= 3, "split tensor-core path: at least two hidden layers");
  SegDev a, b;
  memset(&a, 0, sizeof(a));
  memset(&b, 0, sizeof(b));
  auto fill = [](SegDev& d, const TcSegment& s) {
    d.x = s.x; d.outp = s.packed; d.P = s.rows;
    const int pp = tc_train_pair_points(s.nch);
    d.npairs = (s.rows + pp - 1) / pp;
  };
  fill(a, segs[0]);
  if (nseg == 2) fill(b, segs[1]);
  const int na = segs[0].nch, nb = nseg == 2 ? segs[1].nch : 0;
  const int64_t ta = 2 * a.npairs, tb = 2 * b.npairs;       // whole pairs: the reverse sweep reads the stash of a padding sub-tile too
  DUDF_REQUIRE(col0 + (ta + tb) * 128 <= ld, "tensor-core stash too small");
  TcxTrain tr{Ust, (unsigned char*)Aimg, ld, col0, 0, nullptr, nullptr};
  QueryOut out;
  memset(&out, 0, sizeof(out));
  if (nb == 0) {
    if (na == 1) return tcx_launch_cl<1, 0, true>(packed, net, a, b, ta, tb, 0, 0, out, tr, sms, st);
    if (na == 4) return tcx_launch_cl<4, 0, true>(packed, net, a, b, ta, tb, 0, 0, out, tr, sms, st);
    if (na == 10) return tcx_launch_cl<10, 0, true>(packed, net, a, b, ta, tb, 0, 0, out, tr, sms, st);
  } else if (na == 10) {
    if (nb == 4) return tcx_launch_cl<10, 4, true>(packed, net, a, b, ta, tb, 0, 0, out, tr, sms, st);
    if (nb == 1) return tcx_launch_cl<10, 1, true>(packed, net, a, b, ta, tb, 0, 0, out, tr, sms, st);
  }
  DUDF_REQUIRE(false, "split tensor-core training: unsupported segment channel counts (%d, %d)", na, nb);
}

}  // namespace dudf

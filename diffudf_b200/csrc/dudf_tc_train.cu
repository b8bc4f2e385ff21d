// Tensor-core training path: forward jets with stash, reverse sweep and weight gradients on tcgen05.
//
// Same orientation, tile layout and pipeline as the query kernel (dudf_tc.cu / dudf_tc_common.cuh).  Additions:
//   * one launch serves up to two row segments with different jet orders (loss_s1: on-surface rows carry the Hessian
//     jet, the others the gradient jet): the MMA / producer roles only see 128-column sub-tiles, the epilogue picks
//     the per-point math by segment, so the persistent grid is balanced over ALL sub-tile pairs of the step;
//   * the forward stashes the stored pre-activations thread-major (every warp store is one contiguous 512-byte
//     line): the sine argument u0 in fp32, the derivative channels in fp16; the MMA thread bulk-copies (TMA engine,
//     smem -> global) each finished B-operand tile as two 32 KB [256 neurons][64 columns] images — exactly the
//     K-major operands of the weight-gradient GEMM;
//   * the reverse sweep runs the chain backwards with the transposed weight images, reads the stash, applies the
//     sine-jet adjoint per thread and emits the pre-activation adjoints as its B operand (and, through the same bulk
//     copies, as operand images); all adjoints carry a power-of-two loss scale S chosen from max|seed| so that fp16
//     keeps its full significand over the observed 1e6 dynamic range of the adjoints;
//   * the weight gradient of every hidden layer is one split-K GEMM over all columns (M = N = 256, K = columns).
#include <cuda_fp16.h>
#include <cstdlib>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"
#include "dudf_umma.cuh"
#include "dudf_tc_common.cuh"
#include "dudf_loss.cuh"

namespace dudf {

using namespace umma;

int tc_train_pair_cols(int nch) { (void)nch; return 256; }
int tc_train_pair_points(int nch) { return nch == 1 ? 2 * TcCfg<1>::PT : nch == 4 ? 2 * TcCfg<4>::PT : 2 * TcCfg<10>::PT; }

// fp32 pair -> packed halves, saturating at the largest finite half (one F2FP.SATFINITE.PACK_AB): adjoints that would
// overflow fp16 clamp instead of becoming inf (the loss scale leaves 29x head-room, so this is a safety net)
__device__ __forceinline__ uint32_t tt_pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <int GC>
__device__ __forceinline__ void tt_store_group_sat(const float* v, unsigned char* trow, int chunk0, uint32_t r7) {
  const uint32_t base = umma::smem_u32(trow);
#pragma unroll
  for (int c8 = 0; c8 < GC / 8; ++c8)
    tc_sts128(base + tc_chunk_off(chunk0 + c8, r7), tt_pack_h2_sat(v[8 * c8], v[8 * c8 + 1]), tt_pack_h2_sat(v[8 * c8 + 2], v[8 * c8 + 3]),
              tt_pack_h2_sat(v[8 * c8 + 4], v[8 * c8 + 5]), tt_pack_h2_sat(v[8 * c8 + 6], v[8 * c8 + 7]));
}

struct EpiCtx {
  unsigned char* act;
  float* xs;
  float* os;
  const float* wl_s;
  uint64_t* act_ready;
  uint64_t* acc_ready;
  uint32_t tmem_lane;
  uint32_t r7;
  int n, tid, lane;
  uint32_t acc_phase;
  unsigned long long* trace = nullptr;   // diagnostics (tools/trace_probe.py fused): event log of CTA 0 / warp 0
  uint32_t tn = 0;
};

// =============================================================================================
// forward: all layers of one sub-tile pair
// =============================================================================================
template <int NCH>
__device__ __forceinline__ void tt_fwd_pair(EpiCtx& e, const NetView& net, const SegDev& sg, int64_t pair, int64_t colp, float* Ust, int64_t ld,
                                            float r0x, float r0y, float r0z, float b0) {
  using C = TcCfg<NCH>;
  const int L = net.n_lin - 1;
  const float w0 = net.w0, ww = net.ww;
  tc_epi_bar();
  for (int i = e.tid; i < 2 * C::PT; i += 256) {
    const int64_t p = pair * 2 * C::PT + i;
    float pt[3] = {0.f, 0.f, 0.f};
    if (p < sg.P) { pt[0] = sg.x[p * 3]; pt[1] = sg.x[p * 3 + 1]; pt[2] = sg.x[p * 3 + 2]; }
    e.xs[i * 3] = pt[0]; e.xs[i * 3 + 1] = pt[1]; e.xs[i * 3 + 2] = pt[2];
  }
  if (C::NV < 128) {       // idle columns of both tiles are zero for this segment's math
    for (int s = 0; s < 2; ++s) *reinterpret_cast<uint4*>(tc_tile_row(e.act + s * TC_ACT_BYTES, e.n) + tc_chunk_off(15, e.r7)) = make_uint4(0, 0, 0, 0);
  }
  tc_epi_bar();
  for (int l = 0; l < L; ++l) {
    const float bias = (l > 0) ? ww * net.b[l][e.n] : 0.f;
    for (int s = 0; s < 2; ++s) {
      unsigned char* trow = tc_tile_row(e.act + s * TC_ACT_BYTES, e.n);
      const int64_t colt = colp + s * 128;                        // first stash column of this sub-tile
      float* ust = Ust + ((size_t)l * ld + colt) * 256 + e.n * 4;
      if (l > 0) {
        mbar_wait(&e.acc_ready[s], (e.acc_phase >> s) & 1u, 0x400 + s);
        e.acc_phase ^= 1u << s;
        tc_fence_after();
      }
      if (l == 0) {
#pragma unroll 1
        for (int g = 0; g < C::NGRP; ++g) {
          float u[C::GC];
          tc_first_layer_group<NCH, C::GC>(u, e.xs + (s * C::PT + g * (C::GC / NCH)) * 3, w0, r0x, r0y, r0z, b0);
          tc_emit_group<NCH, C::GC, true>(u, trow, g * (C::GC / 8), e.r7);      // no stash: the reverse sweep recomputes this layer from the points
        }
      } else {
        TmemRegs<C::GC> nxt;
        tc_ld_issue<C::GC>(e.tmem_lane + s * 256, nxt);
#pragma unroll 1
        for (int g = 0; g < C::NGRP; ++g) {
          float u[C::GC];
          tc_ld_take<C::GC>(nxt, u);
          if (g + 1 < C::NGRP) tc_ld_issue<C::GC>(e.tmem_lane + s * 256 + (g + 1) * C::GC, nxt);
#pragma unroll
          for (int pp = 0; pp < C::GC / NCH; ++pp) u[pp * NCH] += bias;
          tt_stash_group<NCH, C::GC>(u, ust + (size_t)g * C::GC * 256, 1);
          tc_emit_group<NCH, C::GC, true>(u, trow, g * (C::GC / 8), e.r7);
        }
      }
      if (l < L - 1) {
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (e.lane == 0) mbar_arrive(&e.act_ready[s]);
      } else {
        tc_epi_bar();
        tc_output_dot<C::NV>(e.act + s * TC_ACT_BYTES, e.wl_s, e.os + s * 256, e.tid);
        tc_epi_bar();
        if (e.tid < C::NV) {
          const int ch = e.tid % NCH;
          const int64_t p = (pair * 2 + s) * C::PT + e.tid / NCH;
          const float v = e.os[s * 256 + e.tid] + e.os[s * 256 + 128 + e.tid];
          if (p < sg.P) sg.outp[p * NCH + ch] = (ch == 0) ? v + net.b[L][0] : (ch >= 4 ? v * TC_KAPPA_INV : v);
        }
      }
    }
  }
}

template <int NA, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
tt_forward_kernel(const unsigned char* __restrict__ packed, NetView net, SegDev sa, SegDev sb, float* __restrict__ Ust,
                  unsigned char* __restrict__ Aimg, int64_t ld, int64_t col0) {
  using C = TcCfg<NA>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  float* wl_s = (float*)(smem + C::OFF_WL);
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TC_STAGES, *act_ready = bars + 2 * TC_STAGES, *acc_ready = bars + 2 * TC_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t npairs = sa.npairs + sb.npairs;
  const int64_t rounds = ((int64_t)blockIdx.x < npairs) ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // pairs of this CTA
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (tid < 256) wl_s[tid] = net.W[L][tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<TC_REGS_AUX>();
    if (warp == 8) {
      if (lane == 0) tc_producer<1>(packed, ring, full, empty, rounds, L - 1, TC_DIR_FWD);
    } else if (warp == 9) {
      tc_mma_role<1>(act, ring, full, empty, act_ready, acc_ready, tmem_base, rounds, L - 1, Aimg, nullptr, ld >> 6, col0 >> 6, TC_DIR_FWD);
    }
  } else {
    setmaxnreg_inc<TC_REGS_EPI>();
    const int q = warp & 3, h = warp >> 2;
    EpiCtx e;
    e.act = act; e.xs = (float*)(smem + C::OFF_XS); e.os = (float*)(smem + C::OFF_OS); e.wl_s = wl_s;
    e.act_ready = act_ready; e.acc_ready = acc_ready;
    e.n = h * 128 + q * 32 + lane; e.tid = tid; e.lane = lane; e.r7 = e.n & 7;
    e.tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    e.acc_phase = 0;
    const float r0x = net.W[0][e.n * 3], r0y = net.W[0][e.n * 3 + 1], r0z = net.W[0][e.n * 3 + 2], b0 = net.b[0][e.n];
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t colp = col0 + pair * 256;
      if (pair < sa.npairs) {
        tt_fwd_pair<NA>(e, net, sa, pair, colp, Ust, ld, r0x, r0y, r0z, b0);
      } else {
        if constexpr (NB > 0) tt_fwd_pair<NB>(e, net, sb, pair - sa.npairs, colp, Ust, ld, r0x, r0y, r0z, b0);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

// =============================================================================================
// reverse sweep
// =============================================================================================
// one column group: u (stash), ab (adjoints of the activations: seeds x output weights at the top, TMEM below) ->
// adjoints of u, written to the B tile (layers > 0); accumulates the thread's gradient partial sums
// RECOMP0 (with FIRST): the first layer's pre-activations are recomputed from the points instead of read from the stash
struct FirstRow { float w0, rx, ry, rz, b; };
template <int NCH, int GC, bool TOP, bool FIRST, bool RECOMP0 = false, bool NOW = false>
__device__ __forceinline__ void tt_bwd_group(const uint4* raw, TmemRegs<GC>& tr, uint32_t next_taddr, float wl, const float* sdg,
                                             const float* pts, unsigned char* trow, int chunk0, uint32_t r7, float& bsum, float& wlsum,
                                             float (&w0s)[3], const FirstRow* fr = nullptr) {
  float u[GC], ab[GC];
  // NOW (fused kernel): the group's own accumulators are loaded here and used in place — the TMEM load (tens of clocks) hides behind
  // the scratch conversion below; prefetching them a group ahead made ptxas copy all 32 / 40 staging registers out and back
  // (64 moves per 8-point group, profiles/r1b_ncu_fused_hot_sass.txt)
  if constexpr (!TOP && NOW) tc_ld_issue<GC>(next_taddr, tr);
  if constexpr (FIRST && RECOMP0) tc_first_layer_group<NCH, GC>(u, pts, fr->w0, fr->rx, fr->ry, fr->rz, fr->b);
  else tt_unstash_group<NCH, GC>(u, raw);
  if constexpr (TOP) {
#pragma unroll
    for (int j = 0; j < GC; ++j) ab[j] = wl * sdg[j];
  } else {
    tc_ld_take<GC>(tr, ab);
    if constexpr (!NOW) {
      if (next_taddr) tc_ld_issue<GC>(next_taddr, tr);         // next group's accumulators, in flight during the math
    }
  }
#pragma unroll
  for (int pp = 0; pp < GC / NCH; ++pp) {
    float sn, cs, ub[NCH];
    if constexpr (FIRST) sincos_fast(u[pp * NCH], sn, cs);                    // first layer: arguments up to ~50 rad
    else { sn = __sinf(u[pp * NCH]); cs = __cosf(u[pp * NCH]); }               // hidden layers: a few rad, MUFU on the raw argument
    tc_adj_point<NCH>(u + pp * NCH, ab + pp * NCH, ub, sn, cs);
    if constexpr (TOP) {
      tc_act_point<NCH>(u + pp * NCH, sn, cs);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) wlsum = fmaf(sdg[pp * NCH + ch], u[pp * NCH + ch], wlsum);
    }
    bsum += ub[0];
    if constexpr (FIRST) {
      const float* pt = pts + pp * 3;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float t = ub[0] * pt[d];
        if constexpr (NCH >= 4) t += ub[1 + d];
        w0s[d] += t;
      }
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) ab[pp * NCH + ch] = ub[ch];
  }
  if constexpr (!FIRST) tt_store_group_sat<GC>(ab, trow, chunk0, r7);
}

template <int NCH>
__device__ __forceinline__ void tt_bwd_pair(EpiCtx& e, const NetView& net, const GradView& grad, const SegDev& sg, int64_t pair, int64_t colp,
                                            const float* Ust, int64_t ld, float S, float wl, bool prefetch_l2, bool first_pair, int64_t colp_next,
                                            const FirstRow& fr) {
  using C = TcCfg<NCH>;
  const int L = net.n_lin - 1;
  const float invS = 1.0f / S;
  float* sd = e.os;                                               // stored seeds of both sub-tiles [2][256]
  tc_trace(e.trace, e.tn, 14, NCH);
  tc_epi_bar();
  for (int i = e.tid; i < 2 * C::PT; i += 256) {
    const int64_t p = pair * 2 * C::PT + i;
    float pt[3] = {0.f, 0.f, 0.f};
    if (p < sg.P) { pt[0] = sg.x[p * 3]; pt[1] = sg.x[p * 3 + 1]; pt[2] = sg.x[p * 3 + 2]; }
    e.xs[i * 3] = pt[0]; e.xs[i * 3 + 1] = pt[1]; e.xs[i * 3 + 2] = pt[2];
  }
  for (int i = e.tid; i < 2 * C::NV; i += 256) {
    const int s = i / C::NV, j = i % C::NV;
    const int64_t p = (pair * 2 + s) * C::PT + j / NCH;
    const int ch = j % NCH;
    float v = 0.f;
    if (p < sg.P) {
      v = sg.seeds[p * NCH + ch];
      if (ch == 0 && v != 0.f) atomicAdd(&grad.b[L][0], v);
      v *= (ch >= 4) ? S * TC_KAPPA_INV : S;
    }
    sd[s * 256 + j] = v;
  }
  if (C::NV < 128) {
    for (int s = 0; s < 2; ++s) *reinterpret_cast<uint4*>(tc_tile_row(e.act + s * TC_ACT_BYTES, e.n) + tc_chunk_off(15, e.r7)) = make_uint4(0, 0, 0, 0);
  }
  tc_epi_bar();
  constexpr int NRAW = Stash<NCH, C::GC>::CHUNKS;
  // stash of (layer l, sub-tile s): the first column group of the NEXT (layer, sub-tile) is requested while the last group of the current
  // one is being worked on — an HBM read takes 1 500-2 000 clk under this kernel's load, and a request issued only at the boundary was
  // fully exposed (0.372 -> 0.348 ms per launch).  Requesting TWO groups ahead costs more in register copies than it hides (0.362 ms).
  auto stash_of = [&](int l, int s) { return Ust + ((size_t)l * ld + colp + s * 128) * 256 + e.n * 4; };
  // ... and the whole stash of a (layer, sub-tile) is asked into the L2 TWO units ahead (one thread, one bulk prefetch per column
  // group: NRAW contiguous 4 KB rows), so that the register loads above find it there instead of in HBM
  auto prefetch_unit = [&](int l, int s, int64_t col) {
    if (e.tid == 0 && prefetch_l2) {
      const float* base = Ust + ((size_t)l * ld + col + s * 128) * 256;
#pragma unroll
      for (int g = 0; g < C::NGRP; ++g) bulk_prefetch_l2(base + (size_t)g * C::GC * 256, NRAW * 4096);
    }
  };
  if (first_pair) {                 // later pairs: requested by the previous pair of this CTA (below)
    prefetch_unit(L - 1, 0, colp);
    prefetch_unit(L - 1, 1, colp);
  }
  // the first layer's pre-activations are recomputed from the points (RECOMP0): no stash is written or read for layer 0
  uint4 nxt[NRAW];
  if (L > 1) tt_stash_load<NCH, C::GC>(nxt, stash_of(L - 1, 0));
  for (int l = L - 1; l >= 0; --l) {
    const float wl_cur = (l == 0) ? net.w0 : net.ww;
    const bool top = (l == L - 1), first = (l == 0);
    for (int s = 0; s < 2; ++s) {
      unsigned char* trow = tc_tile_row(e.act + s * TC_ACT_BYTES, e.n);
      const float* ust = stash_of(l, s);
      const bool more = (s == 0) ? (l > 0) : (l > 1);           // the next (layer, sub-tile) reads a stash
      const float* ust_next = more ? ((s == 0) ? stash_of(l, 1) : stash_of(l - 1, 0)) : ust;
      tc_trace(e.trace, e.tn, 40 + s, l);
      if (l > 1) prefetch_unit(l - 1, s, colp);
      else if (colp_next >= 0) prefetch_unit(L - 1, s, colp_next);      // the top layer of this CTA's next pair (group geometry of THIS
                                                                        // segment: a pair of the other jet order is covered approximately)
      if (!top) {
        mbar_wait(&e.acc_ready[s], (e.acc_phase >> s) & 1u, 0x400 + s);
        e.acc_phase ^= 1u << s;
        tc_fence_after();
      }
      tc_trace(e.trace, e.tn, 30 + s, l);
      float bsum = 0.f, wlsum = 0.f, w0s[3] = {0.f, 0.f, 0.f};
      TmemRegs<C::GC> tr;
#pragma unroll 1
      for (int g = 0; g < C::NGRP; ++g) {
        uint4 ug[NRAW];
        if (!first) {
#pragma unroll
          for (int j = 0; j < NRAW; ++j) ug[j] = nxt[j];
          if (g + 1 < C::NGRP) tt_stash_load<NCH, C::GC>(nxt, ust + (size_t)(g + 1) * C::GC * 256);
          else if (more) tt_stash_load<NCH, C::GC>(nxt, ust_next);
        }
        // the group's own accumulators are loaded inside the group and used in place (NOW): the TMEM load hides behind the stash
        // conversion, and ptxas no longer copies the 32 / 40 staging registers of a prefetched group out and back
        const uint32_t tcur = e.tmem_lane + s * 256 + g * C::GC;
        const float* sdg = sd + s * 256 + g * C::GC;
        const float* pts = e.xs + (s * C::PT + g * (C::GC / NCH)) * 3;
        const int c0 = g * (C::GC / 8);
        if (top && first) tt_bwd_group<NCH, C::GC, true, true, true, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s, &fr);
        else if (top)     tt_bwd_group<NCH, C::GC, true, false, false, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s);
        else if (first)   tt_bwd_group<NCH, C::GC, false, true, true, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s, &fr);
        else              tt_bwd_group<NCH, C::GC, false, false, false, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s);
      }
      tc_trace(e.trace, e.tn, 32 + s, l);
      atomicAdd(&grad.b[l][e.n], bsum * wl_cur * invS);
      if (top) atomicAdd(&grad.W[L][e.n], wlsum * invS);
      if (first) {
#pragma unroll
        for (int d = 0; d < 3; ++d) atomicAdd(&grad.W[0][e.n * 3 + d], w0s[d] * net.w0 * invS);
      } else {
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (e.lane == 0) mbar_arrive(&e.act_ready[s]);
      }
    }
  }
}

template <int NA, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
tt_backward_kernel(const unsigned char* __restrict__ packed, NetView net, GradView grad, SegDev sa, SegDev sb,
                   const float* __restrict__ seed_absmax, const float* __restrict__ Ust, unsigned char* __restrict__ Zimg, int64_t ld,
                   int64_t col0, unsigned long long* trace, int prefetch_l2) {
  using C = TcCfg<NA>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TC_STAGES, *act_ready = bars + 2 * TC_STAGES, *acc_ready = bars + 2 * TC_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t npairs = sa.npairs + sb.npairs;
  const int64_t rounds = ((int64_t)blockIdx.x < npairs) ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // pairs of this CTA
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<TC_REGS_AUX>();
    if (warp == 8) {
      if (lane == 0) tc_producer<1>(packed, ring, full, empty, rounds, L - 1, TC_DIR_BWD);
    } else if (warp == 9) {
      tc_mma_role<1>(act, ring, full, empty, act_ready, acc_ready, tmem_base, rounds, L - 1, nullptr, Zimg, ld >> 6, col0 >> 6, TC_DIR_BWD, 0,
                     (prefetch_l2 & 2) ? l2_policy_evict_first() : 0ull, (trace && blockIdx.x == 0) ? trace : nullptr);
    }
  } else {
    setmaxnreg_inc<TC_REGS_EPI>();
    const int q = warp & 3, h = warp >> 2;
    EpiCtx e;
    e.trace = (trace && blockIdx.x == 0 && warp == 0) ? trace + TC_TRACE_REGION : nullptr;
    e.act = act; e.xs = (float*)(smem + C::OFF_XS); e.os = (float*)(smem + C::OFF_OS); e.wl_s = nullptr;
    e.act_ready = act_ready; e.acc_ready = acc_ready;
    e.n = h * 128 + q * 32 + lane; e.tid = tid; e.lane = lane; e.r7 = e.n & 7;
    e.tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    e.acc_phase = 0;
    const float S = loss_scale_from(seed_absmax);
    const float wl = net.W[L][e.n];
    FirstRow fr;
    fr.w0 = net.w0; fr.rx = net.W[0][e.n * 3]; fr.ry = net.W[0][e.n * 3 + 1]; fr.rz = net.W[0][e.n * 3 + 2]; fr.b = net.b[0][e.n];
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      const int64_t colp = col0 + pair * 256;
      const bool first_pair = pair == (int64_t)blockIdx.x;
      const int64_t colp_next = (pair + gridDim.x < npairs) ? col0 + (pair + gridDim.x) * 256 : -1;
      if (pair < sa.npairs) {
        tt_bwd_pair<NA>(e, net, grad, sa, pair, colp, Ust, ld, S, wl, (prefetch_l2 & 1) != 0, first_pair, colp_next, fr);
      } else {
        if constexpr (NB > 0) tt_bwd_pair<NB>(e, net, grad, sb, pair - sa.npairs, colp, Ust, ld, S, wl, (prefetch_l2 & 1) != 0, first_pair, colp_next, fr);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

// =============================================================================================
// weight gradients of the hidden layers: gW_l[n][k] += (ww / S) sum_col Zimg_l[n][col] * Aimg_{l-1}[k][col]
// Operand images are [layer][column block of 64][256 neurons][64 columns] fp16, 128-byte swizzled (K-major).
// =============================================================================================
constexpr int TW_STAGES = 3;
constexpr int TW_SMEM = TW_STAGES * 2 * TC_IMG_BYTES + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
tt_wgrad_kernel(GradView grad, const unsigned char* __restrict__ Zimg, const unsigned char* __restrict__ Aimg, int64_t ncb, int splits,
                float ww, const float* __restrict__ seed_absmax, int layer0, int evict_first) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + TW_STAGES * 2 * TC_IMG_BYTES);
  uint64_t *full = bars, *empty = bars + TW_STAGES, *done = bars + 2 * TW_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TW_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = layer0 + blockIdx.x / splits;
  const int split = blockIdx.x % splits;
  const int64_t per = (ncb + splits - 1) / splits;
  const int64_t cb0 = split * per, cb1 = min(ncb, cb0 + per);
  if (tid == 0) {
    for (int i = 0; i < TW_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const unsigned char* zsrc = Zimg + (size_t)l * ncb * TC_IMG_BYTES;
  const unsigned char* asrc = Aimg + (size_t)(l - 1) * ncb * TC_IMG_BYTES;
  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint64_t policy = evict_first ? l2_policy_evict_first() : 0ull;
      // (asking the blocks behind the ring into the L2 with bulk prefetches makes this HBM-bound kernel 30-45 % SLOWER: measured, removed)
      for (int64_t cb = cb0; cb < cb1; ++cb) {
        mbar_wait(&empty[stage], phase ^ 1, 0x500 + stage);
        mbar_arrive_expect_tx(&full[stage], 2 * TC_IMG_BYTES);
        if (policy) {             // the images are read exactly once: keep them from displacing what the next step's kernels find in the L2
          bulk_g2s_hint(smem + stage * 2 * TC_IMG_BYTES, zsrc + (size_t)cb * TC_IMG_BYTES, TC_IMG_BYTES, &full[stage], policy);
          bulk_g2s_hint(smem + stage * 2 * TC_IMG_BYTES + TC_IMG_BYTES, asrc + (size_t)cb * TC_IMG_BYTES, TC_IMG_BYTES, &full[stage], policy);
        } else {
          bulk_g2s(smem + stage * 2 * TC_IMG_BYTES, zsrc + (size_t)cb * TC_IMG_BYTES, TC_IMG_BYTES, &full[stage]);
          bulk_g2s(smem + stage * 2 * TC_IMG_BYTES + TC_IMG_BYTES, asrc + (size_t)cb * TC_IMG_BYTES, TC_IMG_BYTES, &full[stage]);
        }
        if (++stage == TW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // whole warp, warp-uniform code, one elected lane issues (see tc_mma_role)
    constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, 0, 0);
    const uint64_t desc0 = make_desc_sw128(smem_u32(smem), 16, 1024);
    uint32_t stage = 0, phase = 0;
    for (int64_t cb = cb0; cb < cb1; ++cb) {
      mbar_wait(&full[stage], phase, 0x600 + stage);
      tc_fence_after();
      const uint64_t z_desc = desc_advance(desc0, stage * 2 * TC_IMG_BYTES);
      const uint64_t a_desc = desc_advance(z_desc, TC_IMG_BYTES);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          mma_f16_ss_warp(tmem_base + h * 256, desc_advance(z_desc, h * 16384 + k4 * 32), desc_advance(a_desc, k4 * 32), idesc,
                          (cb > cb0) || (k4 != 0));
      mma_commit_warp(&empty[stage]);
      if (++stage == TW_STAGES) { stage = 0; phase ^= 1; }
    }
    mma_commit_warp(done);
  } else if (cb1 > cb0) {
    // warps 0-3: drain the two 128 x 256 accumulators into the gradient with vector reductions
    mbar_wait(done, 0, 0x700);
    tc_fence_after();
    const float factor = ww / loss_scale_from(seed_absmax);
    float* dst = grad.W[l];
    for (int h = 0; h < 2; ++h) {
      const int n = h * 128 + warp * 32 + lane;
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + h * 256 + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float* p = dst + (size_t)n * 256 + c0 + j;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(r[j]) * factor),
                       "f"(__uint_as_float(r[j + 1]) * factor), "f"(__uint_as_float(r[j + 2]) * factor),
                       "f"(__uint_as_float(r[j + 3]) * factor)
                       : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem_base);
}

// =============================================================================================
// launchers
// =============================================================================================
static SegDev make_seg(const TcSegment& s) {
  SegDev d;
  d.x = s.x; d.outp = s.packed; d.seeds = s.seeds; d.normals = s.normals; d.dist = s.dist; d.P = s.rows;
  const int pp = tc_train_pair_points(s.nch);
  d.npairs = (s.rows + pp - 1) / pp;
  return d;
}

template <int NA, int NB>
static int tt_launch_fwd(const void* packed, const NetView& net, const SegDev& a, const SegDev& b, float* Ust, void* Aimg, int64_t ld,
                         int64_t col0, int sms, cudaStream_t st) {
  auto k = tt_forward_kernel<NA, NB>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<NA>::SMEM));
  const int grid = (int)std::min<int64_t>(a.npairs + b.npairs, sms);
  if (grid < 1) return 0;
  k<<<grid, TC_THREADS, TcCfg<NA>::SMEM, st>>>((const unsigned char*)packed, net, a, b, Ust, (unsigned char*)Aimg, ld, col0);
  DUDF_LAUNCH_OK();
  return 0;
}

template <int NA, int NB>
static int tt_launch_bwd(const void* packed, const NetView& net, const GradView& grad, const SegDev& a, const SegDev& b,
                         const float* seed_absmax, const float* Ust, void* Zimg, int64_t ld, int64_t col0, int sms, cudaStream_t st) {
  auto k = tt_backward_kernel<NA, NB>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<NA>::SMEM));
  const int grid = (int)std::min<int64_t>(a.npairs + b.npairs, sms);
  if (grid < 1) return 0;
  static const int prefetch_l2 = [] { const char* e = getenv("DUDF_BWD_PREFETCH_L2"); return e ? atoi(e) : 3; }();   // A/B switch: bit 0 stash prefetch, bit 1 evict-first adjoint images
  k<<<grid, TC_THREADS, TcCfg<NA>::SMEM, st>>>((const unsigned char*)packed, net, grad, a, b, seed_absmax, Ust, (unsigned char*)Zimg, ld, col0,
                                               tc_get_trace(), prefetch_l2);
  DUDF_LAUNCH_OK();
  return 0;
}

// segments must occupy consecutive column ranges of the stash starting at col0 (256 columns per pair)
int tc_train_forward(const void* packed, const NetView& net, const TcSegment* segs, int nseg, float* Ust, void* Aimg, int64_t ld, int64_t col0,
                     int sms, cudaStream_t st) {
  DUDF_REQUIRE(ld % 64 == 0 && col0 % 128 == 0, "tensor-core stash: ld must be a multiple of 64 and col0 of 128");
  DUDF_REQUIRE(nseg == 1 || nseg == 2, "tensor-core training: 1 or 2 segments per launch");
  SegDev a = make_seg(segs[0]), b;
  memset(&b, 0, sizeof(b));
  const int na = segs[0].nch, nb = nseg == 2 ? segs[1].nch : 0;
  if (nseg == 2) b = make_seg(segs[1]);
  if (nb == 0) {
    if (na == 1) return tt_launch_fwd<1, 0>(packed, net, a, b, Ust, Aimg, ld, col0, sms, st);
    if (na == 4) return tt_launch_fwd<4, 0>(packed, net, a, b, Ust, Aimg, ld, col0, sms, st);
    if (na == 10) return tt_launch_fwd<10, 0>(packed, net, a, b, Ust, Aimg, ld, col0, sms, st);
  } else if (na == 10) {
    if (nb == 4) return tt_launch_fwd<10, 4>(packed, net, a, b, Ust, Aimg, ld, col0, sms, st);
    if (nb == 1) return tt_launch_fwd<10, 1>(packed, net, a, b, Ust, Aimg, ld, col0, sms, st);
  }
  DUDF_REQUIRE(false, "tensor-core training: unsupported segment channel counts (%d, %d)", na, nb);
}

int tc_train_backward(const void* packed, const NetView& net, const GradView& grad, const TcSegment* segs, int nseg, const float* seed_absmax,
                      const float* Ust, void* Zimg, int64_t ld, int64_t col0, int sms, cudaStream_t st) {
  DUDF_REQUIRE(seed_absmax != nullptr, "tensor-core reverse sweep needs the seed magnitude (loss scale)");
  DUDF_REQUIRE(ld % 64 == 0 && col0 % 128 == 0, "tensor-core stash: ld must be a multiple of 64 and col0 of 128");
  DUDF_REQUIRE(nseg == 1 || nseg == 2, "tensor-core training: 1 or 2 segments per launch");
  SegDev a = make_seg(segs[0]), b;
  memset(&b, 0, sizeof(b));
  const int na = segs[0].nch, nb = nseg == 2 ? segs[1].nch : 0;
  if (nseg == 2) b = make_seg(segs[1]);
  if (nb == 0) {
    if (na == 1) return tt_launch_bwd<1, 0>(packed, net, grad, a, b, seed_absmax, Ust, Zimg, ld, col0, sms, st);
    if (na == 4) return tt_launch_bwd<4, 0>(packed, net, grad, a, b, seed_absmax, Ust, Zimg, ld, col0, sms, st);
    if (na == 10) return tt_launch_bwd<10, 0>(packed, net, grad, a, b, seed_absmax, Ust, Zimg, ld, col0, sms, st);
  } else if (na == 10) {
    if (nb == 4) return tt_launch_bwd<10, 4>(packed, net, grad, a, b, seed_absmax, Ust, Zimg, ld, col0, sms, st);
    if (nb == 1) return tt_launch_bwd<10, 1>(packed, net, grad, a, b, seed_absmax, Ust, Zimg, ld, col0, sms, st);
  }
  DUDF_REQUIRE(false, "tensor-core training: unsupported segment channel counts (%d, %d)", na, nb);
}

// layers [layer_lo, layer_hi) of the hidden 256 x 256 matrices (1 .. L-1); the whole grid is spent on that range, so a caller can
// launch the layers in groups and hand each finished group to the gradient all-reduce while the next one runs
int tc_train_wgrad(const NetView& net, const GradView& grad, const void* Zimg, const void* Aimg, int64_t ld, const float* seed_absmax,
                   int sms, cudaStream_t st, int layer_lo, int layer_hi) {
  const int L = net.n_lin - 1;
  if (L < 2 || ld <= 0) return 0;
  if (layer_lo <= 0) layer_lo = 1;
  if (layer_hi <= 0 || layer_hi > L) layer_hi = L;
  if (layer_hi <= layer_lo) return 0;
  DUDF_REQUIRE(ld % 64 == 0, "tensor-core stash: ld must be a multiple of 64");
  DUDF_REQUIRE(seed_absmax != nullptr, "tensor-core weight gradient needs the seed magnitude (loss scale)");
  const int64_t ncb = ld / 64;
  const int nl = layer_hi - layer_lo;
  int splits = std::max(1, sms / nl);
  splits = (int)std::min<int64_t>(splits, ncb);
  static const int wgrad_evict_first = [] { const char* e = getenv("DUDF_WGRAD_EVICT_FIRST"); return e ? atoi(e) : 1; }();   // A/B switch
  DUDF_CUDA_OK(cudaFuncSetAttribute(tt_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
  tt_wgrad_kernel<<<nl * splits, 192, TW_SMEM, st>>>(grad, (const unsigned char*)Zimg, (const unsigned char*)Aimg, ncb, splits, net.ww,
                                                    seed_absmax, layer_lo, wgrad_evict_first);
  DUDF_LAUNCH_OK();
  return 0;
}

// =============================================================================================
// fused step (loss_s1 / loss_siren): forward, loss epilogue and reverse sweep of a sub-tile pair inside one CTA.
// The pre-activation stash becomes a per-CTA scratch of 2 sub-tiles x n_hidden layers that is written, read back and
// overwritten by the same CTA: it lives in L2 (consumed lines are discarded, not written back), so the HBM traffic
// of a step is the operand images of the weight-gradient GEMM only.  The first layer is recomputed in the reverse
// sweep.  The loss scale of the adjoints comes from the PREVIOUS step's max|seed| (the seeds of a step are not known
// before its forward); this step's maximum is collected for the next one.
// =============================================================================================
struct FusedDev {
  LossRowCfg loss;
  double* terms;             // [4], accumulated
  const float* amax_prev;    // max|stored seed| of the previous step -> loss scale
  float* amax_next;          // this step's (atomicMax)
  int flags;                 // bit 0: discard consumed scratch lines from L2
  unsigned long long* trace; // diagnostics: event log of CTA 0 (MMA warp, epilogue warp 0) or null
};

template <int NCH>
__device__ __forceinline__ void tt_fused_pair(EpiCtx& e, const NetView& net, const GradView& grad, const SegDev& sg, const FusedDev& fd,
                                              int64_t pair, float* Ucta, float S, float wl, const FirstRow& fr, double* acc_terms,
                                              float* acc_misc) {
  using C = TcCfg<NCH>;
  constexpr int GC = C::GC;
  constexpr int NRAW = Stash<NCH, GC>::CHUNKS;
  const int L = net.n_lin - 1;
  const float ww = net.ww;
  const float invS = 1.0f / S;
  float* sd = e.os;                                               // outputs, then stored seeds, of both sub-tiles [2][256]
  const bool nostash = ((fd.flags >> 8) & 32) != 0;               // diagnostics: no scratch traffic (results are meaningless)
  tc_trace(e.trace, e.tn, 14, NCH);
  tc_epi_bar();
  for (int i = e.tid; i < 2 * C::PT; i += 256) {
    const int64_t p = pair * 2 * C::PT + i;
    float pt[3] = {0.f, 0.f, 0.f};
    if (p < sg.P) { pt[0] = sg.x[p * 3]; pt[1] = sg.x[p * 3 + 1]; pt[2] = sg.x[p * 3 + 2]; }
    e.xs[i * 3] = pt[0]; e.xs[i * 3 + 1] = pt[1]; e.xs[i * 3 + 2] = pt[2];
  }
  if (C::NV < 128) {
    for (int s = 0; s < 2; ++s) *reinterpret_cast<uint4*>(tc_tile_row(e.act + s * TC_ACT_BYTES, e.n) + tc_chunk_off(15, e.r7)) = make_uint4(0, 0, 0, 0);
  }
  tc_epi_bar();
  // ---------------- forward ----------------
  for (int l = 0; l < L; ++l) {
    const float bias = (l > 0) ? ww * net.b[l][e.n] : 0.f;
    for (int s = 0; s < 2; ++s) {
      unsigned char* trow = tc_tile_row(e.act + s * TC_ACT_BYTES, e.n);
      float* ust = Ucta + ((size_t)l * 256 + s * 128) * 256 + e.n * 4;
      tc_trace(e.trace, e.tn, 42 + s, l);
      if (l > 0) {
        mbar_wait(&e.acc_ready[s], (e.acc_phase >> s) & 1u, 0x400 + s);
        e.acc_phase ^= 1u << s;
        tc_fence_after();
      }
      tc_trace(e.trace, e.tn, 10 + s, l);
      if (l == 0) {
#pragma unroll 1
        for (int g = 0; g < C::NGRP; ++g) {
          float u[GC];
          tc_first_layer_group<NCH, GC>(u, e.xs + (s * C::PT + g * (GC / NCH)) * 3, fr.w0, fr.rx, fr.ry, fr.rz, fr.b);
          tc_emit_group<NCH, GC, true>(u, trow, g * (GC / 8), e.r7);
        }
      } else {
        TmemRegs<GC> nxt;
#pragma unroll 1
        for (int g = 0; g < C::NGRP; ++g) {
          float u[GC];
          tc_ld_issue<GC>(e.tmem_lane + s * 256 + g * GC, nxt);      // loaded and used in place (see tt_bwd_group)
          tc_ld_take<GC>(nxt, u);
#pragma unroll
          for (int pp = 0; pp < GC / NCH; ++pp) u[pp * NCH] += bias;
          if (!nostash) tt_stash_group<NCH, GC>(u, ust + (size_t)g * GC * 256);
          tc_emit_group<NCH, GC, true>(u, trow, g * (GC / 8), e.r7);
        }
      }
      tc_trace(e.trace, e.tn, 12 + s, l);
      if (l < L - 1) {
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (e.lane == 0) mbar_arrive(&e.act_ready[s]);
      }
    }
  }
  // ---------------- output layer + loss: one thread per point of the pair ----------------
  tc_epi_bar();
  tc_trace(e.trace, e.tn, 20, 0);
  for (int s = 0; s < 2; ++s) tc_output_dot<C::NV>(e.act + s * TC_ACT_BYTES, e.wl_s, e.os + s * 256, e.tid);
  tc_epi_bar();
  tc_trace(e.trace, e.tn, 21, 0);
  if ((e.tid & ~31) < 2 * C::PT) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    float bl = 0.f, amax = 0.f;
    if (e.tid < 2 * C::PT) {
      const int s = e.tid / C::PT, i = e.tid % C::PT;
      const int64_t p = pair * 2 * C::PT + e.tid;
      float* col = sd + s * 256 + i * NCH;
      float v[10], sv[10];
#pragma unroll
      for (int ch = 0; ch < 10; ++ch) {
        v[ch] = 0.f;
        if (ch < NCH) {
          const float r = col[ch] + col[128 + ch];
          v[ch] = (ch == 0) ? r + net.b[L][0] : (ch >= 4 ? r * TC_KAPPA_INV : r);
        }
      }
#pragma unroll
      for (int ch = 0; ch < 10; ++ch) sv[ch] = 0.f;
      if (p < sg.P) {
        loss_row(fd.loss, NCH, v, sg.dist[p], sg.normals + p * 3, t, sv);
        if (sg.outp) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) sg.outp[p * NCH + ch] = v[ch];
        }
      }
      bl = sv[0];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const float st = sv[ch] * (ch >= 4 ? TC_KAPPA_INV : 1.f);
        amax = fmaxf(amax, fabsf(st));
        col[ch] = st * S;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
      bl += __shfl_xor_sync(0xffffffffu, bl, o);
      amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    }
    if (e.lane == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t[k] != 0.0) atomicAdd(&acc_terms[k], t[k]);
      if (bl != 0.f) atomicAdd(&acc_misc[0], bl);
      if (amax > 0.f && isfinite(amax)) atomicMax(reinterpret_cast<int*>(&acc_misc[1]), __float_as_int(amax));
    }
  }
  tc_epi_bar();
  tc_trace(e.trace, e.tn, 22, 0);
  // ---------------- reverse sweep: layers L-1 .. 1 read the scratch, layer 0 is recomputed from the points ----------------
  for (int l = L - 1; l >= 1; --l) {
    const bool top = (l == L - 1);
    for (int s = 0; s < 2; ++s) {
      unsigned char* trow = tc_tile_row(e.act + s * TC_ACT_BYTES, e.n);
      const float* ust = Ucta + ((size_t)l * 256 + s * 128) * 256 + e.n * 4;
      uint4 nxt[NRAW];
      if (!nostash) tt_stash_load<NCH, GC>(nxt, ust);
      else {
#pragma unroll
        for (int j = 0; j < NRAW; ++j) nxt[j] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
      }
      tc_trace(e.trace, e.tn, 40 + s, l);
      if (!top) {
        mbar_wait(&e.acc_ready[s], (e.acc_phase >> s) & 1u, 0x400 + s);
        e.acc_phase ^= 1u << s;
        tc_fence_after();
      }
      tc_trace(e.trace, e.tn, 30 + s, l);
      float bsum = 0.f, wlsum = 0.f, w0s[3] = {0.f, 0.f, 0.f};
      TmemRegs<GC> tr;
#pragma unroll 1
      for (int g = 0; g < C::NGRP; ++g) {
        uint4 ug[NRAW];
#pragma unroll
        for (int j = 0; j < NRAW; ++j) ug[j] = nxt[j];
        if (g + 1 < C::NGRP && !nostash) tt_stash_load<NCH, GC>(nxt, ust + (size_t)(g + 1) * GC * 256);
        const uint32_t tcur = e.tmem_lane + s * 256 + g * GC;
        const float* sdg = sd + s * 256 + g * GC;
        const float* pts = e.xs + (s * C::PT + g * (GC / NCH)) * 3;
        const int c0 = g * (GC / 8);
        if (top) tt_bwd_group<NCH, GC, true, false, false, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s);
        else     tt_bwd_group<NCH, GC, false, false, false, true>(ug, tr, tcur, wl, sdg, pts, trow, c0, e.r7, bsum, wlsum, w0s);
        if ((fd.flags & 1) && (e.lane & 7) == 0) {
#pragma unroll
          for (int j = 0; j < NRAW; ++j) l2_discard_line(ust + (size_t)g * GC * 256 + j * 1024);
        }
      }
      tc_trace(e.trace, e.tn, 32 + s, l);
      atomicAdd(&grad.b[l][e.n], bsum * ww * invS);
      if (top) atomicAdd(&grad.W[L][e.n], wlsum * invS);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (e.lane == 0) mbar_arrive(&e.act_ready[s]);
    }
  }
  for (int s = 0; s < 2; ++s) {
    tc_trace(e.trace, e.tn, 40 + s, 0);
    mbar_wait(&e.acc_ready[s], (e.acc_phase >> s) & 1u, 0x400 + s);
    e.acc_phase ^= 1u << s;
    tc_fence_after();
    tc_trace(e.trace, e.tn, 30 + s, 0);
    float bsum = 0.f, wlsum = 0.f, w0s[3] = {0.f, 0.f, 0.f};
    TmemRegs<GC> tr;
#pragma unroll 1
    for (int g = 0; g < C::NGRP; ++g) {
      const uint32_t tcur = e.tmem_lane + s * 256 + g * GC;
      const float* pts = e.xs + (s * C::PT + g * (GC / NCH)) * 3;
      tt_bwd_group<NCH, GC, false, true, true, true>(nullptr, tr, tcur, wl, nullptr, pts, nullptr, 0, e.r7, bsum, wlsum, w0s, &fr);
    }
    tc_trace(e.trace, e.tn, 32 + s, 0);
    atomicAdd(&grad.b[0][e.n], bsum * net.w0 * invS);
#pragma unroll
    for (int d = 0; d < 3; ++d) atomicAdd(&grad.W[0][e.n * 3 + d], w0s[d] * net.w0 * invS);
  }
  tc_trace(e.trace, e.tn, 15, NCH);
}

template <int NA, int NB>
__global__ void __launch_bounds__(TC_THREADS, 1)
tt_fused_kernel(const unsigned char* __restrict__ packed, NetView net, GradView grad, SegDev sa, SegDev sb, FusedDev fd,
                float* __restrict__ scratch, unsigned char* __restrict__ Aimg, unsigned char* __restrict__ Zimg, int64_t ncb) {
  using C = TcCfg<NA>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  float* wl_s = (float*)(smem + C::OFF_WL);
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TC_STAGES, *act_ready = bars + 2 * TC_STAGES, *acc_ready = bars + 2 * TC_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 4);
  double* acc_terms = (double*)(smem + C::OFF_BAR + 128);          // [4] loss-term shares of this CTA
  float* acc_misc = (float*)(smem + C::OFF_BAR + 160);             // [0] output-bias gradient, [1] max|stored seed|
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t npairs = sa.npairs + sb.npairs;
  const int64_t rounds = ((int64_t)blockIdx.x < npairs) ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
    for (int k = 0; k < 4; ++k) acc_terms[k] = 0.0;
    acc_misc[0] = 0.f; acc_misc[1] = 0.f;
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (tid < 256) wl_s[tid] = net.W[L][tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (((fd.flags >> 8) & 128) && (blockIdx.x & 1)) {        // diagnostics: odd CTAs start half a pair late (are the store-heavy
    const long long t0 = clock64();                          // forward sweeps of all CTAs at the same time the problem?  No:
    while (clock64() - t0 < 120000) {}                       // the per-layer forward time does not change)
  }

  if (warp >= 8) {
    setmaxnreg_dec<TC_REGS_AUX>();
    if (warp == 8) {
      if (lane == 0) tc_producer<1>(packed, ring, full, empty, rounds, L - 1, TC_DIR_BOTH, 0, fd.flags >> 8);
    } else if (warp == 9) {
      tc_mma_role<1>(act, ring, full, empty, act_ready, acc_ready, tmem_base, rounds, L - 1, Aimg, Zimg, ncb, 0, TC_DIR_BOTH, fd.flags >> 8,
                       (fd.flags & 2) ? l2_policy_evict_first() : 0ull, (fd.trace && blockIdx.x == 0) ? fd.trace : nullptr);
    }
  } else {
    setmaxnreg_inc<TC_REGS_EPI>();
    const int q = warp & 3, h = warp >> 2;
    EpiCtx e;
    e.act = act; e.xs = (float*)(smem + C::OFF_XS); e.os = (float*)(smem + C::OFF_OS); e.wl_s = wl_s;
    e.act_ready = act_ready; e.acc_ready = acc_ready;
    e.n = h * 128 + q * 32 + lane; e.tid = tid; e.lane = lane; e.r7 = e.n & 7;
    e.tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    e.acc_phase = 0;
    e.trace = (fd.trace && blockIdx.x == 0 && warp == 0) ? fd.trace + TC_TRACE_REGION : nullptr;
    const float S = loss_scale_from(fd.amax_prev);
    const float wl = net.W[L][e.n];
    FirstRow fr;
    fr.w0 = net.w0; fr.rx = net.W[0][e.n * 3]; fr.ry = net.W[0][e.n * 3 + 1]; fr.rz = net.W[0][e.n * 3 + 2]; fr.b = net.b[0][e.n];
    float* Ucta = scratch + (size_t)blockIdx.x * L * 256 * 256;
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      if (pair < sa.npairs) {
        tt_fused_pair<NA>(e, net, grad, sa, fd, pair, Ucta, S, wl, fr, acc_terms, acc_misc);
      } else {
        if constexpr (NB > 0) tt_fused_pair<NB>(e, net, grad, sb, fd, pair - sa.npairs, Ucta, S, wl, fr, acc_terms, acc_misc);
      }
    }
    tc_epi_bar();
    if (tid < 4 && acc_terms[tid] != 0.0) atomicAdd(&fd.terms[tid], acc_terms[tid]);
    if (tid == 4 && acc_misc[0] != 0.f) atomicAdd(&grad.b[L][0], acc_misc[0]);
    if (tid == 5 && acc_misc[1] > 0.f) atomicMax(reinterpret_cast<int*>(fd.amax_next), __float_as_int(acc_misc[1]));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

size_t tc_fused_scratch_bytes(const NetView& net, int sms) { return (size_t)sms * (net.n_lin - 1) * 256 * 256 * sizeof(float); }

template <int NA, int NB>
static int tt_launch_fused(const void* packed, const NetView& net, const GradView& grad, const SegDev& a, const SegDev& b, const FusedDev& fd,
                           float* scratch, void* Aimg, void* Zimg, int64_t ld, int sms, cudaStream_t st) {
  auto k = tt_fused_kernel<NA, NB>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<NA>::SMEM));
  const int grid = (int)std::min<int64_t>(a.npairs + b.npairs, sms);
  if (grid < 1) return 0;
  k<<<grid, TC_THREADS, TcCfg<NA>::SMEM, st>>>((const unsigned char*)packed, net, grad, a, b, fd, scratch, (unsigned char*)Aimg,
                                               (unsigned char*)Zimg, ld >> 6);
  DUDF_LAUNCH_OK();
  return 0;
}

// One fused launch over 1 or 2 row segments (segment columns are consecutive from column 0 of the operand images).
int tc_train_fused(const void* packed, const NetView& net, const GradView& grad, const TcSegment* segs, int nseg, const TcFusedLoss& fl,
                   float* scratch, void* Aimg, void* Zimg, int64_t ld, int sms, cudaStream_t st) {
  DUDF_REQUIRE(ld % 64 == 0, "tensor-core stash: ld must be a multiple of 64");
  DUDF_REQUIRE(nseg == 1 || nseg == 2, "tensor-core training: 1 or 2 segments per launch");
  DUDF_REQUIRE(fl.mode == DUDF_LOSS_S1 || fl.mode == DUDF_LOSS_SIREN, "fused step: loss_s1 or loss_siren (loss_s2 needs batch statistics first)");
  DUDF_REQUIRE(fl.terms && fl.amax_prev && fl.amax_next && scratch, "fused step: null argument");
  SegDev a = make_seg(segs[0]), b;
  memset(&b, 0, sizeof(b));
  const int na = segs[0].nch, nb = nseg == 2 ? segs[1].nch : 0;
  if (nseg == 2) b = make_seg(segs[1]);
  DUDF_REQUIRE((a.npairs + b.npairs) * 256 <= ld, "fused step: operand images too small");
  FusedDev fd;
  fd.loss.mode = fl.mode; fd.loss.alpha = fl.alpha; fd.loss.invP = 1.0f / (float)fl.P_global;
  for (int k = 0; k < 4; ++k) { fd.loss.w[k] = fl.w[k]; fd.loss.up[k] = 1.f; }
  fd.terms = fl.terms; fd.amax_prev = fl.amax_prev; fd.amax_next = fl.amax_next; fd.flags = fl.flags;
  fd.trace = tc_get_trace();
  if (nb == 0) {
    if (na == 4) return tt_launch_fused<4, 0>(packed, net, grad, a, b, fd, scratch, Aimg, Zimg, ld, sms, st);
    if (na == 10) return tt_launch_fused<10, 0>(packed, net, grad, a, b, fd, scratch, Aimg, Zimg, ld, sms, st);
  } else if (na == 10 && nb == 4) {
    return tt_launch_fused<10, 4>(packed, net, grad, a, b, fd, scratch, Aimg, Zimg, ld, sms, st);
  }
  DUDF_REQUIRE(false, "fused step: unsupported segment channel counts (%d, %d)", na, nb);
}

}  // namespace dudf

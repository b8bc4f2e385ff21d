// Tensor-core training path: forward jets with stash, reverse sweep and weight gradients on tcgen05.
//
// Same orientation and pipeline as the query kernel (dudf_tc.cu): lane = neuron, column = one jet
// channel of one point, weights are the UMMA A operand streamed through a bulk-copy ring, the
// activations (forward) / pre-activation adjoints (backward) are the B operand written by the
// epilogue warps.  Differences:
//   * stored variables: u0 = w z0, u_i = w z_i, v_ij = KAPPA w z_ij and a0 = sin u0, a_i = cos(u0) u_i,
//     b_ij = KAPPA a_ij (KAPPA = 1/8 keeps second-order channels inside fp16 range);
//   * the forward stashes the stored pre-activations in fp32, thread-major ([layer][column group][4-vector][neuron])
//     so that every warp store is one contiguous 512-byte line, and the stored activations as fp16: the MMA
//     thread bulk-copies (TMA engine, smem -> global) the finished B-operand tile into per-k-block planes
//     [layer][kblock][column][64 neurons], which are exactly the MN-major operands of the weight gradient;
//   * the backward runs the chain in reverse with the transposed weight images, reads the stash,
//     applies the sine-jet adjoint per thread and emits the adjoints both as the next B operand and
//     as fp16 images; all adjoints carry a power-of-two loss scale S chosen from max|seed| so that
//     fp16 keeps its full significand over the observed 1e6 dynamic range;
//   * the weight gradient of every hidden layer is one split-K GEMM over all columns
//     (M = N = 256, K = columns) whose operands are those images.
#include <cuda_fp16.h>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"
#include "dudf_umma.cuh"

namespace dudf {

using namespace umma;

constexpr int TT_CHUNK = 128 * 64 * 2;
constexpr int TT_STAGES = 5;
constexpr int TT_THREADS = 320;
constexpr float TT_KAPPA = 0.125f;
constexpr float TT_KAPPA_INV = 8.0f;
constexpr int TT_IMG = 256 * 128;             // bytes of one 256 x 64 fp16 image

template <int NCH>
struct TtCfg {
  static constexpr int PT = (NCH == 1) ? 128 : (NCH == 4 ? 32 : 8);
  static constexpr int N = PT * NCH;
  static constexpr int GC = (NCH == 10) ? 40 : 32;
  static constexpr int KB_BYTES = N * 128;
  static constexpr int ACT_BYTES = 4 * KB_BYTES;
  static constexpr int OFF_RING = 2 * ACT_BYTES;
  static constexpr int OFF_WL = OFF_RING + TT_STAGES * TT_CHUNK;
  static constexpr int OFF_XS = OFF_WL + 256 * 4;
  static constexpr int OFF_OS = OFF_XS + 2 * PT * 3 * 4;              // forward: outputs; backward: seeds
  static constexpr int OFF_BAR = (OFF_OS + 2 * N * 4 + 15) / 16 * 16;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
};

int tc_train_pair_cols(int nch) { return nch == 1 ? 2 * TtCfg<1>::N : nch == 4 ? 2 * TtCfg<4>::N : 2 * TtCfg<10>::N; }
int tc_train_pair_points(int nch) { return nch == 1 ? 2 * TtCfg<1>::PT : nch == 4 ? 2 * TtCfg<4>::PT : 2 * TtCfg<10>::PT; }

__device__ __forceinline__ float loss_scale_from(const float* seed_absmax) {
  const float m = seed_absmax ? *seed_absmax : 0.f;
  return (m > 0.f && isfinite(m)) ? exp2f(floorf(log2f(2048.f / m))) : 1.f;
}

__device__ __forceinline__ void tt_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -60000.f), 60000.f); }

// bulk copy shared -> global (TMA engine), grouped completion
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- roles shared by the forward and backward chain kernels ------------------------------------
// image index of MMA phase j (0 .. n_phase-1): forward uses layer j+1, backward layer L-1-j (transposed images)
__device__ __forceinline__ const unsigned char* tt_image(const unsigned char* packed, int n_phase, int j, bool backward) {
  const int idx = backward ? (n_phase + (n_phase - 1 - j)) : j;
  return packed + (size_t)idx * 8 * TT_CHUNK;
}

__device__ __forceinline__ void tt_producer(const unsigned char* packed, unsigned char* ring, uint64_t* full, uint64_t* empty,
                                            int64_t npairs, int n_phase, bool backward) {
  uint32_t stage = 0, phase = 0;
  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
    for (int j = 0; j < n_phase; ++j) {
      const unsigned char* src = tt_image(packed, n_phase, j, backward);
      for (int s = 0; s < 2; ++s)
        for (int ck = 0; ck < 8; ++ck) {
          mbar_wait(&empty[stage], phase ^ 1, 0x100 + stage);
          mbar_arrive_expect_tx(&full[stage], TT_CHUNK);
          bulk_g2s(ring + stage * TT_CHUNK, src + (size_t)ck * TT_CHUNK, TT_CHUNK, &full[stage]);
          if (++stage == TT_STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

// `img` (may be null): plane array [layer][4][ld][128 B]; phase j's finished tile of sub-tile s is copied to the rows
// [col0 + (pair*2+s)*N, +N) of the 4 planes of layer img_layer(j) (forward: j, backward: L-1-j)
template <int NCH>
__device__ __forceinline__ void tt_mma(unsigned char* act, unsigned char* ring, uint64_t* full, uint64_t* empty, uint64_t* act_ready,
                                       uint64_t* acc_ready, uint32_t tmem_base, int64_t npairs, int n_phase, unsigned char* img,
                                       int64_t ld, int64_t col0, bool backward) {
  using C = TtCfg<NCH>;
  constexpr uint32_t idesc = make_idesc_f16(128, C::N, 0, 0, 0);
  uint32_t stage = 0, phase = 0;
  uint32_t act_phase[2] = {0, 0};
  for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x)
    for (int j = 0; j < n_phase; ++j)
      for (int s = 0; s < 2; ++s) {
        mbar_wait(&act_ready[s], act_phase[s], 0x200 + s);
        act_phase[s] ^= 1;
        tc_fence_after();
        const uint32_t act_s = smem_u32(act + s * C::ACT_BYTES);
        for (int h = 0; h < 2; ++h) {
          const uint32_t d_tmem = tmem_base + s * 256 + h * 128;
          for (int kb = 0; kb < 4; ++kb) {
            mbar_wait(&full[stage], phase, 0x300 + stage);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + stage * TT_CHUNK);
            const uint32_t b_addr = act_s + kb * C::KB_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              mma_f16_ss(d_tmem, make_desc_sw128(a_addr + k4 * 32, 16, 1024), make_desc_sw128(b_addr + k4 * 32, 16, 1024), idesc,
                         (kb | k4) != 0);
            mma_commit(&empty[stage]);
            if (++stage == TT_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if (img) {
          const int layer = backward ? (n_phase - j) : j;
          const int64_t row0 = col0 + (pair * 2 + s) * C::N;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            bulk_s2g(img + (((size_t)layer * 4 + kb) * ld + row0) * 128, act + s * C::ACT_BYTES + kb * C::KB_BYTES, C::KB_BYTES);
          bulk_commit();
          bulk_wait_read0();          // the tile may be overwritten once acc_ready is published
        }
        mma_commit(&acc_ready[s]);
      }
}

// ---- per-point math in stored variables ---------------------------------------------------------
template <int NCH>
__device__ __forceinline__ void act_point(const float* u, float* a, float s, float c) {
  a[0] = s;
  if constexpr (NCH >= 4) {
#pragma unroll
    for (int i = 0; i < 3; ++i) a[1 + i] = c * u[1 + i];
  }
  if constexpr (NCH >= 10) {
    const float ks = TT_KAPPA * s;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) a[4 + sym2(i, j)] = fmaf(c, u[4 + sym2(i, j)], -ks * u[1 + i] * u[1 + j]);
  }
}

template <int NCH>
__device__ __forceinline__ void adj_point(const float* u, const float* ab, float* ub, float s, float c) {
  float u0 = c * ab[0];
  if constexpr (NCH >= 4) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      acc = fmaf(ab[1 + i], u[1 + i], acc);
      ub[1 + i] = c * ab[1 + i];
    }
    u0 = fmaf(-s, acc, u0);
    if constexpr (NCH >= 10) {
      const float ks = TT_KAPPA * s, kc = TT_KAPPA * c;
      float acc2 = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j) {
          const int q = 4 + sym2(i, j);
          acc2 = fmaf(ab[q], fmaf(s, u[q], kc * u[1 + i] * u[1 + j]), acc2);
          ub[q] = c * ab[q];
          if (i == j) {
            ub[1 + i] = fmaf(-2.f * ks * ab[q], u[1 + i], ub[1 + i]);
          } else {
            ub[1 + i] = fmaf(-ks * ab[q], u[1 + j], ub[1 + i]);
            ub[1 + j] = fmaf(-ks * ab[q], u[1 + i], ub[1 + j]);
          }
        }
      u0 -= acc2;
    }
  }
  ub[0] = u0;
}

// write GC values of one thread (its neuron n) into the smem B-operand tile (2-byte scattered, K-major rows)
template <int GC>
__device__ __forceinline__ void emit_group(const float* v, unsigned char* tile_g, const uint32_t* sw) {
#pragma unroll
  for (int j = 0; j < GC; ++j) *reinterpret_cast<__half*>(tile_g + j * 128 + sw[j & 7]) = __float2half_rn(v[j]);
}

// =============================================================================================
// forward with stash
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(TT_THREADS, 1)
tt_forward_kernel(const unsigned char* __restrict__ packed, NetView net, const float* __restrict__ x, int64_t P, float* __restrict__ outp,
                  float* __restrict__ Ust, unsigned char* __restrict__ Aimg, int64_t ld, int64_t col0) {
  using C = TtCfg<NCH>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  float* wl_s = (float*)(smem + C::OFF_WL);
  float* xs = (float*)(smem + C::OFF_XS);
  float* os = (float*)(smem + C::OFF_OS);
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TT_STAGES, *act_ready = bars + 2 * TT_STAGES, *acc_ready = bars + 2 * TT_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TT_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  const int64_t ncb = ld >> 6;
  if (tid == 0) {
    for (int i = 0; i < TT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (tid < 256) wl_s[tid] = net.W[L][tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) tt_producer(packed, ring, full, empty, npairs, L - 1, false);
  } else if (warp == 9) {
    if (lane == 0) tt_mma<NCH>(act, ring, full, empty, act_ready, acc_ready, tmem_base, npairs, L - 1, Aimg, ld, col0, false);
  } else {
    const int q = warp & 3, h = warp >> 2;
    const int n = h * 128 + q * 32 + lane;
    const uint32_t chunk = (n & 63) >> 3;
    uint32_t sw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sw[j] = ((chunk ^ j) << 4);
    const uint32_t tile_off = (n >> 6) * C::KB_BYTES + (n & 7) * 2;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    const float w0 = net.w0, ww = net.ww;
    const float r0x = net.W[0][n * 3], r0y = net.W[0][n * 3 + 1], r0z = net.W[0][n * 3 + 2], b0 = net.b[0][n];
    const float bL = net.b[L][0];
    uint32_t acc_phase[2] = {0, 0};
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      tt_epi_bar();
      for (int i = tid; i < 2 * C::PT; i += 256) {
        const int64_t p = pair * 2 * C::PT + i;
        float pt[3] = {0.f, 0.f, 0.f};
        if (p < P) { pt[0] = x[p * 3]; pt[1] = x[p * 3 + 1]; pt[2] = x[p * 3 + 2]; }
        xs[i * 3] = pt[0]; xs[i * 3 + 1] = pt[1]; xs[i * 3 + 2] = pt[2];
      }
      tt_epi_bar();
      for (int l = 0; l < L; ++l) {
        const float bias = (l > 0) ? ww * net.b[l][n] : 0.f;
        for (int s = 0; s < 2; ++s) {
          unsigned char* tile = act + s * C::ACT_BYTES + tile_off;
          const int64_t colt = col0 + (pair * 2 + s) * C::N;           // first stash column of this sub-tile
          float* ust = Ust + ((size_t)l * ld + colt) * 256 + n * 4;     // thread-major: [column group][4-vector][neuron]
          if (l > 0) {
            mbar_wait(&acc_ready[s], acc_phase[s], 0x400 + s);
            acc_phase[s] ^= 1;
            tc_fence_after();
          }
#pragma unroll 1
          for (int g = 0; g < C::N / C::GC; ++g) {
            float u[C::GC], a[C::GC];
            if (l == 0) {
#pragma unroll
              for (int pp = 0; pp < C::GC / NCH; ++pp) {
                const float* pt = xs + (s * C::PT + g * (C::GC / NCH) + pp) * 3;
                float* up = u + pp * NCH;
                up[0] = w0 * fmaf(r0z, pt[2], fmaf(r0y, pt[1], fmaf(r0x, pt[0], b0)));
                if constexpr (NCH >= 4) { up[1] = w0 * r0x; up[2] = w0 * r0y; up[3] = w0 * r0z; }
#pragma unroll
                for (int ch = 4; ch < NCH; ++ch) up[ch] = 0.f;
              }
            } else {
              uint32_t r[C::GC];
              const uint32_t taddr = tmem_lane + s * 256 + g * C::GC;
              tmem_ld_x32(taddr, r);
              if constexpr (C::GC == 40) tmem_ld_x8(taddr + 32, r + 32);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < C::GC; ++j) u[j] = __uint_as_float(r[j]);
#pragma unroll
              for (int pp = 0; pp < C::GC / NCH; ++pp) u[pp * NCH] += bias;
            }
#pragma unroll
            for (int pp = 0; pp < C::GC / NCH; ++pp) {
              float sn, cs;
              sincos_fast(u[pp * NCH], sn, cs);
              act_point<NCH>(u + pp * NCH, a + pp * NCH, sn, cs);
            }
#pragma unroll
            for (int j4 = 0; j4 < C::GC / 4; ++j4)
              *reinterpret_cast<float4*>(ust + (size_t)g * C::GC * 256 + j4 * 1024) =
                  make_float4(u[j4 * 4], u[j4 * 4 + 1], u[j4 * 4 + 2], u[j4 * 4 + 3]);
            emit_group<C::GC>(a, tile + g * C::GC * 128, sw);
          }
          if (l < L - 1) {
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[s]);
          } else {
            tt_epi_bar();
            if (tid < C::N) {
              const unsigned char* rowp = act + s * C::ACT_BYTES + tid * 128;
              float sum = 0.f;
#pragma unroll
              for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                  const uint4 v = *reinterpret_cast<const uint4*>(rowp + kb * C::KB_BYTES + ((c8 ^ (tid & 7)) << 4));
                  const __half2* hv = reinterpret_cast<const __half2*>(&v);
                  const float* wv = wl_s + kb * 64 + c8 * 8;
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f2 = __half22float2(hv[e]);
                    sum = fmaf(wv[2 * e], f2.x, sum);
                    sum = fmaf(wv[2 * e + 1], f2.y, sum);
                  }
                }
              const int ch = tid % NCH;
              os[s * C::N + tid] = (ch == 0) ? sum + bL : (ch >= 4 ? sum * TT_KAPPA_INV : sum);
            }
            tt_epi_bar();
            if (tid < C::PT) {
              const int64_t p = (pair * 2 + s) * C::PT + tid;
              if (p < P) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) outp[p * NCH + ch] = os[s * C::N + tid * NCH + ch];
              }
            }
          }
        }
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
template <int NCH>
__global__ void __launch_bounds__(TT_THREADS, 1)
tt_backward_kernel(const unsigned char* __restrict__ packed, NetView net, GradView grad, const float* __restrict__ x, int64_t P,
                   const float* __restrict__ seeds, const float* __restrict__ seed_absmax, const float* __restrict__ Ust,
                   unsigned char* __restrict__ Zimg, int64_t ld, int64_t col0) {
  using C = TtCfg<NCH>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  float* xs = (float*)(smem + C::OFF_XS);
  float* sd = (float*)(smem + C::OFF_OS);                 // stored seeds of both sub-tiles [2][N]
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TT_STAGES, *act_ready = bars + 2 * TT_STAGES, *acc_ready = bars + 2 * TT_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TT_STAGES + 4);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  const int64_t ncb = ld >> 6;
  if (tid == 0) {
    for (int i = 0; i < TT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) tt_producer(packed, ring, full, empty, npairs, L - 1, true);
  } else if (warp == 9) {
    if (lane == 0) tt_mma<NCH>(act, ring, full, empty, act_ready, acc_ready, tmem_base, npairs, L - 1, Zimg, ld, col0, true);
  } else {
    const int q = warp & 3, h = warp >> 2;
    const int n = h * 128 + q * 32 + lane;
    const uint32_t chunk = (n & 63) >> 3;
    uint32_t sw[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sw[j] = ((chunk ^ j) << 4);
    const uint32_t tile_off = (n >> 6) * C::KB_BYTES + (n & 7) * 2;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    const float S = loss_scale_from(seed_absmax);
    const float invS = 1.0f / S;
    const float wl = net.W[L][n];
    const float w0 = net.w0, ww = net.ww;
    uint32_t acc_phase[2] = {0, 0};
    for (int64_t pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
      tt_epi_bar();
      for (int i = tid; i < 2 * C::PT; i += 256) {
        const int64_t p = pair * 2 * C::PT + i;
        float pt[3] = {0.f, 0.f, 0.f};
        if (p < P) { pt[0] = x[p * 3]; pt[1] = x[p * 3 + 1]; pt[2] = x[p * 3 + 2]; }
        xs[i * 3] = pt[0]; xs[i * 3 + 1] = pt[1]; xs[i * 3 + 2] = pt[2];
      }
      for (int i = tid; i < 2 * C::N; i += 256) {
        const int64_t p = pair * 2 * C::PT + i / NCH;
        const int ch = i % NCH;
        float v = 0.f;
        if (p < P) {
          v = seeds[p * NCH + ch];
          if (ch == 0 && v != 0.f) atomicAdd(&grad.b[L][0], v);
          v *= (ch >= 4) ? S * TT_KAPPA_INV : S;
        }
        sd[i] = v;
      }
      tt_epi_bar();
      for (int l = L - 1; l >= 0; --l) {
        const float wl_cur = (l == 0) ? w0 : ww;
        for (int s = 0; s < 2; ++s) {
          unsigned char* tile = act + s * C::ACT_BYTES + tile_off;
          const int64_t colt = col0 + (pair * 2 + s) * C::N;
          const float* ust = Ust + ((size_t)l * ld + colt) * 256 + n * 4;
          if (l < L - 1) {
            mbar_wait(&acc_ready[s], acc_phase[s], 0x400 + s);
            acc_phase[s] ^= 1;
            tc_fence_after();
          }
          float bsum = 0.f, wlsum = 0.f, w0s[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
          for (int g = 0; g < C::N / C::GC; ++g) {
            float u[C::GC], ab[C::GC], ub[C::GC];
#pragma unroll
            for (int j4 = 0; j4 < C::GC / 4; ++j4) {
              const float4 t = *reinterpret_cast<const float4*>(ust + (size_t)g * C::GC * 256 + j4 * 1024);
              u[j4 * 4] = t.x; u[j4 * 4 + 1] = t.y; u[j4 * 4 + 2] = t.z; u[j4 * 4 + 3] = t.w;
            }
            if (l == L - 1) {
#pragma unroll
              for (int j = 0; j < C::GC; ++j) ab[j] = wl * sd[s * C::N + g * C::GC + j];
            } else {
              uint32_t r[C::GC];
              const uint32_t taddr = tmem_lane + s * 256 + g * C::GC;
              tmem_ld_x32(taddr, r);
              if constexpr (C::GC == 40) tmem_ld_x8(taddr + 32, r + 32);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < C::GC; ++j) ab[j] = __uint_as_float(r[j]);
            }
#pragma unroll
            for (int pp = 0; pp < C::GC / NCH; ++pp) {
              float sn, cs;
              sincos_fast(u[pp * NCH], sn, cs);
              if (l == L - 1) {
                float a[NCH];
                act_point<NCH>(u + pp * NCH, a, sn, cs);
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) wlsum = fmaf(sd[s * C::N + g * C::GC + pp * NCH + ch], a[ch], wlsum);
              }
              adj_point<NCH>(u + pp * NCH, ab + pp * NCH, ub + pp * NCH, sn, cs);
              bsum += ub[pp * NCH];
              if (l == 0) {
                const float* pt = xs + (s * C::PT + g * (C::GC / NCH) + pp) * 3;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  float t = ub[pp * NCH] * pt[d];
                  if constexpr (NCH >= 4) t += ub[pp * NCH + 1 + d];
                  w0s[d] += t;
                }
              }
            }
            if (l > 0) {
#pragma unroll
              for (int j = 0; j < C::GC; ++j) ub[j] = clamp_h(ub[j]);
              emit_group<C::GC>(ub, tile + g * C::GC * 128, sw);
            }
          }
          atomicAdd(&grad.b[l][n], bsum * wl_cur * invS);
          if (l == L - 1) atomicAdd(&grad.W[L][n], wlsum * invS);
          if (l == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) atomicAdd(&grad.W[0][n * 3 + d], w0s[d] * w0 * invS);
          }
          if (l > 0) {
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[s]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

// =============================================================================================
// weight gradients of the hidden layers: gW_l[n][k] += (ww / S) sum_col Zimg_l[n][col] * Aimg_{l-1}[k][col]
// =============================================================================================
constexpr int TW_STAGES = 3;
constexpr int TW_SMEM = TW_STAGES * 2 * TT_IMG + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
tt_wgrad_kernel(GradView grad, const unsigned char* __restrict__ Zimg, const unsigned char* __restrict__ Aimg, int64_t ncb, int splits,
                float ww, const float* __restrict__ seed_absmax) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + TW_STAGES * 2 * TT_IMG);
  uint64_t *full = bars, *empty = bars + TW_STAGES, *done = bars + 2 * TW_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TW_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = 1 + blockIdx.x / splits;
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
  const unsigned char* zsrc = Zimg + (size_t)l * ncb * TT_IMG;
  const unsigned char* asrc = Aimg + (size_t)(l - 1) * ncb * TT_IMG;
  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t cb = cb0; cb < cb1; ++cb) {
        mbar_wait(&empty[stage], phase ^ 1, 0x500 + stage);
        mbar_arrive_expect_tx(&full[stage], 2 * TT_IMG);
        bulk_g2s(smem + stage * 2 * TT_IMG, zsrc + (size_t)cb * TT_IMG, TT_IMG, &full[stage]);
        bulk_g2s(smem + stage * 2 * TT_IMG + TT_IMG, asrc + (size_t)cb * TT_IMG, TT_IMG, &full[stage]);
        if (++stage == TW_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, 0, 0);
      uint32_t stage = 0, phase = 0;
      for (int64_t cb = cb0; cb < cb1; ++cb) {
        mbar_wait(&full[stage], phase, 0x600 + stage);
        tc_fence_after();
        const uint32_t z_addr = smem_u32(smem + stage * 2 * TT_IMG);
        const uint32_t a_addr = z_addr + TT_IMG;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            mma_f16_ss(tmem_base + h * 256, make_desc_sw128(z_addr + h * 16384 + k4 * 32, 16, 1024),
                       make_desc_sw128(a_addr + k4 * 32, 16, 1024), idesc, (cb > cb0) || (k4 != 0));
        mma_commit(&empty[stage]);
        if (++stage == TW_STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(done);
    }
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

// ---- v2: operands are the per-k-block planes [layer][4][ld rows = columns][64 neurons] written by the chain
// kernels' bulk stores, i.e. MN-major tiles (neuron contiguous, reduction index = row).  One stage = 64 rows of all
// 4 planes of both operands (8 bulk copies of 8 KB).
constexpr int TW2_STAGES = 3;
constexpr int TW2_ROWS = 64;
constexpr int TW2_PLANE = TW2_ROWS * 128;                        // 8 KB
constexpr int TW2_STAGE_BYTES = 8 * TW2_PLANE;                   // Z planes 0..3 | A planes 0..3
constexpr int TW2_SMEM = TW2_STAGES * TW2_STAGE_BYTES + 1024 + 256;

__global__ void __launch_bounds__(192, 1)
tt_wgrad2_kernel(GradView grad, const unsigned char* __restrict__ Zimg, const unsigned char* __restrict__ Aimg, int64_t ld, int splits,
                 float ww, const float* __restrict__ seed_absmax) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + TW2_STAGES * TW2_STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + TW2_STAGES, *done = bars + 2 * TW2_STAGES;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TW2_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int l = 1 + blockIdx.x / splits;
  const int split = blockIdx.x % splits;
  const int64_t nrb = ld / TW2_ROWS;                              // row blocks
  const int64_t per = (nrb + splits - 1) / splits;
  const int64_t rb0 = split * per, rb1 = min(nrb, rb0 + per);
  if (tid == 0) {
    for (int i = 0; i < TW2_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 4) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb0; rb < rb1; ++rb) {
        mbar_wait(&empty[stage], phase ^ 1, 0x500 + stage);
        mbar_arrive_expect_tx(&full[stage], TW2_STAGE_BYTES);
        unsigned char* dst = smem + stage * TW2_STAGE_BYTES;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          bulk_g2s(dst + kb * TW2_PLANE, Zimg + (((size_t)l * 4 + kb) * ld + rb * TW2_ROWS) * 128, TW2_PLANE, &full[stage]);
          bulk_g2s(dst + (4 + kb) * TW2_PLANE, Aimg + (((size_t)(l - 1) * 4 + kb) * ld + rb * TW2_ROWS) * 128, TW2_PLANE, &full[stage]);
        }
        if (++stage == TW2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, 256, 0, /*A MN-major*/ 1, /*B MN-major*/ 1);
      uint32_t stage = 0, phase = 0;
      for (int64_t rb = rb0; rb < rb1; ++rb) {
        mbar_wait(&full[stage], phase, 0x600 + stage);
        tc_fence_after();
        const uint32_t z_addr = smem_u32(smem + stage * TW2_STAGE_BYTES);
        const uint32_t a_addr = z_addr + 4 * TW2_PLANE;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int ks = 0; ks < TW2_ROWS / 16; ++ks)
            mma_f16_ss(tmem_base + h * 256, make_desc_sw128(z_addr + 2 * h * TW2_PLANE + ks * 2048, TW2_PLANE, 1024),
                       make_desc_sw128(a_addr + ks * 2048, TW2_PLANE, 1024), idesc, (rb > rb0) || (ks != 0));
        mma_commit(&empty[stage]);
        if (++stage == TW2_STAGES) { stage = 0; phase ^= 1; }
      }
      mma_commit(done);
    }
  } else if (rb1 > rb0) {
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
template <int NCH>
static int tt_launch_fwd(const void* packed, const NetView& net, const float* x, int64_t P, float* outp, float* Ust, void* Aimg, int64_t ld,
                         int64_t col0, int sms, cudaStream_t st) {
  using C = TtCfg<NCH>;
  auto k = tt_forward_kernel<NCH>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  const int grid = (int)std::min<int64_t>(npairs, sms);
  if (grid < 1) return 0;
  k<<<grid, TT_THREADS, C::SMEM, st>>>((const unsigned char*)packed, net, x, P, outp, Ust, (unsigned char*)Aimg, ld, col0);
  DUDF_LAUNCH_OK();
  return 0;
}

int tc_train_forward(const void* packed, const NetView& net, int nch, const float* x, int64_t P, float* outp, float* Ust, void* Aimg,
                     int64_t ld, int64_t col0, int sms, cudaStream_t st) {
  DUDF_REQUIRE(ld % 64 == 0 && col0 % 8 == 0, "tensor-core stash: ld must be a multiple of 64 and col0 of 8");
  switch (nch) {
    case 1: return tt_launch_fwd<1>(packed, net, x, P, outp, Ust, Aimg, ld, col0, sms, st);
    case 4: return tt_launch_fwd<4>(packed, net, x, P, outp, Ust, Aimg, ld, col0, sms, st);
    case 10: return tt_launch_fwd<10>(packed, net, x, P, outp, Ust, Aimg, ld, col0, sms, st);
  }
  DUDF_REQUIRE(false, "tensor-core training: unsupported channel count %d", nch);
}

template <int NCH>
static int tt_launch_bwd(const void* packed, const NetView& net, const GradView& grad, const float* x, int64_t P, const float* seeds,
                         const float* seed_absmax, const float* Ust, void* Zimg, int64_t ld, int64_t col0, int sms, cudaStream_t st) {
  using C = TtCfg<NCH>;
  auto k = tt_backward_kernel<NCH>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  const int grid = (int)std::min<int64_t>(npairs, sms);
  if (grid < 1) return 0;
  k<<<grid, TT_THREADS, C::SMEM, st>>>((const unsigned char*)packed, net, grad, x, P, seeds, seed_absmax, Ust, (unsigned char*)Zimg, ld, col0);
  DUDF_LAUNCH_OK();
  return 0;
}

int tc_train_backward(const void* packed, const NetView& net, const GradView& grad, int nch, const float* x, int64_t P, const float* seeds,
                      const float* seed_absmax, const float* Ust, void* Zimg, int64_t ld, int64_t col0, int sms, cudaStream_t st) {
  DUDF_REQUIRE(seed_absmax != nullptr, "tensor-core reverse sweep needs the seed magnitude (loss scale)");
  switch (nch) {
    case 1: return tt_launch_bwd<1>(packed, net, grad, x, P, seeds, seed_absmax, Ust, Zimg, ld, col0, sms, st);
    case 4: return tt_launch_bwd<4>(packed, net, grad, x, P, seeds, seed_absmax, Ust, Zimg, ld, col0, sms, st);
    case 10: return tt_launch_bwd<10>(packed, net, grad, x, P, seeds, seed_absmax, Ust, Zimg, ld, col0, sms, st);
  }
  DUDF_REQUIRE(false, "tensor-core training: unsupported channel count %d", nch);
}

int tc_train_wgrad(const NetView& net, const GradView& grad, const void* Zimg, const void* Aimg, int64_t ld, const float* seed_absmax,
                   int sms, cudaStream_t st) {
  const int L = net.n_lin - 1;
  if (L < 2 || ld <= 0) return 0;
  DUDF_REQUIRE(ld % 64 == 0, "tensor-core stash: ld must be a multiple of 64");
  DUDF_REQUIRE(seed_absmax != nullptr, "tensor-core weight gradient needs the seed magnitude (loss scale)");
  const int64_t nrb = ld / TW2_ROWS;
  int splits = std::max(1, sms / (L - 1));
  splits = (int)std::min<int64_t>(splits, nrb);
  DUDF_CUDA_OK(cudaFuncSetAttribute(tt_wgrad2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW2_SMEM));
  tt_wgrad2_kernel<<<(L - 1) * splits, 192, TW2_SMEM, st>>>(grad, (const unsigned char*)Zimg, (const unsigned char*)Aimg, ld, splits, net.ww,
                                                           seed_absmax);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

// fp32 CUDA-core ("precise") path of the DUDF hot path: forward jets of order 0..3 through the
// SIREN MLP, the reverse sweep, the weight-gradient contraction, the loss epilogues and Adam.
// Everything here is plain fp32 FFMA so that results are fp32-grade (tolerance 1e-5 against
// the fp64 oracle); the tcgen05 path (dudf_tc.cu) is the fast one and shares the data layouts.
//
// Reference behaviour restated (no code shared): src/model.py:116-135 (forward),
// src/diff_operators.py:187-212 (gradient / hessian), src/loss_functions.py:82-155 (losses),
// train.py:204-222 (step).  Math: SURVEY.md §8 a-M.
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"

namespace dudf {

template <int NCH>
struct Cfg {
  static constexpr int TN = (NCH == 20) ? 4 : 8;                     // neurons per thread
  static constexpr int PPT = (NCH == 1) ? 8 : (NCH == 4 ? 2 : 1);    // points per thread
  static constexpr int CPT = PPT * NCH;                              // columns per thread
  static constexpr int NG = 256 / TN;                                // threads along neurons
  static constexpr int CG = 256 / NG;                                // threads along columns
  static constexpr int NC = CG * CPT;                                // columns per tile
  static constexpr int PT = CG * PPT;                                // points per tile
  static constexpr int S = NC + 4;                                   // smem row stride (floats)
  static constexpr int KC = 16;                                      // reduction chunk
  static constexpr size_t smem_bytes = sizeof(float) * (256 * S + 2 * KC * 256 + PT * 3 + NC + 16);
};

// ---- the sine "activation" on a jet: z[NCH] (pre-activations incl. bias) -> a[NCH] ----------
template <int NCH>
__device__ __forceinline__ void sine_jet(const float* z, float* a, float w, float s, float c) {
  a[0] = s;
  if constexpr (NCH >= 4) {
    const float wc = w * c;
#pragma unroll
    for (int i = 0; i < 3; ++i) a[1 + i] = wc * z[1 + i];
    if constexpr (NCH >= 10) {
      // written out channel by channel (xx,xy,xz,yy,yz,zz = 4..9; xxx,xxy,xxz,xyy,xyz,xzz,yyy,yyz,yzz,zzz = 10..19): loops over
      // symmetric index tuples keep the accumulator tile in local memory (nvcc 12.9 does not scalarise them: 400-byte stack frame)
      const float w2s = w * w * s;
      const float zx = z[1], zy = z[2], zz = z[3];
      a[4] = wc * z[4] - w2s * zx * zx;
      a[5] = wc * z[5] - w2s * zx * zy;
      a[6] = wc * z[6] - w2s * zx * zz;
      a[7] = wc * z[7] - w2s * zy * zy;
      a[8] = wc * z[8] - w2s * zy * zz;
      a[9] = wc * z[9] - w2s * zz * zz;
      if constexpr (NCH >= 20) {
        const float w3c = w * w * w * c;
        const float xx = z[4], xy = z[5], xz = z[6], yy = z[7], yz = z[8], z2 = z[9];
        a[10] = wc * z[10] - w2s * (xx * zx + xx * zx + xx * zx) - w3c * zx * zx * zx;
        a[11] = wc * z[11] - w2s * (xx * zy + xy * zx + xy * zx) - w3c * zx * zx * zy;
        a[12] = wc * z[12] - w2s * (xx * zz + xz * zx + xz * zx) - w3c * zx * zx * zz;
        a[13] = wc * z[13] - w2s * (xy * zy + xy * zy + yy * zx) - w3c * zx * zy * zy;
        a[14] = wc * z[14] - w2s * (xy * zz + xz * zy + yz * zx) - w3c * zx * zy * zz;
        a[15] = wc * z[15] - w2s * (xz * zz + xz * zz + z2 * zx) - w3c * zx * zz * zz;
        a[16] = wc * z[16] - w2s * (yy * zy + yy * zy + yy * zy) - w3c * zy * zy * zy;
        a[17] = wc * z[17] - w2s * (yy * zz + yz * zy + yz * zy) - w3c * zy * zy * zz;
        a[18] = wc * z[18] - w2s * (yz * zz + yz * zz + z2 * zy) - w3c * zy * zz * zz;
        a[19] = wc * z[19] - w2s * (z2 * zz + z2 * zz + z2 * zz) - w3c * zz * zz * zz;
      }
    }
  }
}

// ---- its adjoint: ab[NCH] = dL/da (stored-variable convention) -> zb[NCH] = dL/dz -------------
template <int NCH>
__device__ __forceinline__ void sine_jet_adjoint(const float* z, const float* ab, float* zb, float w, float s, float c) {
  const float wc = w * c;
  float z0 = wc * ab[0];
  if constexpr (NCH >= 4) {
    const float w2s = w * w * s;
    float acc1 = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      acc1 += ab[1 + i] * z[1 + i];
      zb[1 + i] = wc * ab[1 + i];
    }
    z0 -= w2s * acc1;
    if constexpr (NCH >= 10) {
      // channel by channel (xx,xy,xz,yy,yz,zz = 4..9), same operation order as the index loops this replaces (they kept the tile in
      // local memory, see sine_jet)
      const float w3c = w * w * w * c;
      const float zx = z[1], zy = z[2], zz = z[3];
      float acc2 = 0.f;
      acc2 += ab[4] * (w2s * z[4] + w3c * zx * zx);  zb[4] = wc * ab[4];  zb[1] -= 2.f * w2s * ab[4] * zx;
      acc2 += ab[5] * (w2s * z[5] + w3c * zx * zy);  zb[5] = wc * ab[5];  zb[1] -= w2s * ab[5] * zy;  zb[2] -= w2s * ab[5] * zx;
      acc2 += ab[6] * (w2s * z[6] + w3c * zx * zz);  zb[6] = wc * ab[6];  zb[1] -= w2s * ab[6] * zz;  zb[3] -= w2s * ab[6] * zx;
      acc2 += ab[7] * (w2s * z[7] + w3c * zy * zy);  zb[7] = wc * ab[7];  zb[2] -= 2.f * w2s * ab[7] * zy;
      acc2 += ab[8] * (w2s * z[8] + w3c * zy * zz);  zb[8] = wc * ab[8];  zb[2] -= w2s * ab[8] * zz;  zb[3] -= w2s * ab[8] * zy;
      acc2 += ab[9] * (w2s * z[9] + w3c * zz * zz);  zb[9] = wc * ab[9];  zb[3] -= 2.f * w2s * ab[9] * zz;
      z0 -= acc2;
    }
  }
  zb[0] = z0;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[TN][CPT] = sum_k Wsrc[k][n_i] * As[k][col_j]   (Wsrc: [256][256] fp32 in global, row = reduction index)
template <int NCH>
__device__ __forceinline__ void simt_gemm(const float* __restrict__ Wsrc, const float* As, float* Ws,
                                          float (&acc)[Cfg<NCH>::TN][Cfg<NCH>::CPT], int ng, int cg) {
  using C = Cfg<NCH>;
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < C::TN; ++i)
#pragma unroll
    for (int j = 0; j < C::CPT; ++j) acc[i][j] = 0.f;
  // prefetch chunk 0
#pragma unroll
  for (int r = 0; r < 4; ++r) cp_async16(Ws + (r * 256 + tid) * 4, Wsrc + (r * 256 + tid) * 4);
  cp_async_commit();
  for (int kc = 0; kc < 256 / C::KC; ++kc) {
    float* cur = Ws + (kc & 1) * C::KC * 256;
    if (kc + 1 < 256 / C::KC) {
      float* nxt = Ws + ((kc + 1) & 1) * C::KC * 256;
      const float* src = Wsrc + (size_t)(kc + 1) * C::KC * 256;
#pragma unroll
      for (int r = 0; r < 4; ++r) cp_async16(nxt + (r * 256 + tid) * 4, src + (r * 256 + tid) * 4);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < C::KC; ++kk) {
      float w[C::TN];
#pragma unroll
      for (int i = 0; i < C::TN; ++i) w[i] = cur[kk * 256 + ng + i * C::NG];
      float a[C::CPT];
      const float* arow = As + (kc * C::KC + kk) * C::S + cg * C::CPT;
      if constexpr (C::CPT % 4 == 0) {
#pragma unroll
        for (int j = 0; j < C::CPT / 4; ++j) {
          float4 v = *reinterpret_cast<const float4*>(arow + 4 * j);
          a[4 * j] = v.x; a[4 * j + 1] = v.y; a[4 * j + 2] = v.z; a[4 * j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < C::CPT / 2; ++j) {
          float2 v = *reinterpret_cast<const float2*>(arow + 2 * j);
          a[2 * j] = v.x; a[2 * j + 1] = v.y;
        }
      }
#pragma unroll
      for (int i = 0; i < C::TN; ++i)
#pragma unroll
        for (int j = 0; j < C::CPT; ++j) acc[i][j] = fmaf(w[i], a[j], acc[i][j]);
    }
    __syncthreads();
  }
}

template <int NCH>
__device__ __forceinline__ void store_cols(float* dst, const float* v) {
  using C = Cfg<NCH>;
  if constexpr (C::CPT % 4 == 0) {
#pragma unroll
    for (int j = 0; j < C::CPT / 4; ++j)
      *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < C::CPT / 2; ++j) *reinterpret_cast<float2*>(dst + 2 * j) = make_float2(v[2 * j], v[2 * j + 1]);
  }
}
template <int NCH>
__device__ __forceinline__ void load_cols(const float* src, float* v) {
  using C = Cfg<NCH>;
  if constexpr (C::CPT % 4 == 0) {
#pragma unroll
    for (int j = 0; j < C::CPT / 4; ++j) {
      float4 t = *reinterpret_cast<const float4*>(src + 4 * j);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < C::CPT / 2; ++j) {
      float2 t = *reinterpret_cast<const float2*>(src + 2 * j);
      v[2 * j] = t.x; v[2 * j + 1] = t.y;
    }
  }
}

// =============================================================================================
// forward
// =============================================================================================
template <int NCH, bool STASH>
__global__ void __launch_bounds__(256) simt_forward_kernel(NetView net, const float* __restrict__ x, int64_t P,
                                                           int gridN, int64_t grid_first, QueryOut out,
                                                           float* __restrict__ Zst, float* __restrict__ Ast,
                                                           int64_t ctot, int64_t col0) {
  using C = Cfg<NCH>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + 256 * C::S;
  float* xs = Ws + 2 * C::KC * 256;
  float* os = xs + C::PT * 3;
  const int tid = threadIdx.x;
  const int ng = tid % C::NG, cg = tid / C::NG;
  const int L = net.n_lin - 1;                       // sine layers
  const int64_t ntiles = (P + C::PT - 1) / C::PT;
  const float vs = gridN > 1 ? 2.0f / (float)(gridN - 1) : 0.f;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * C::PT;
    __syncthreads();
    if (tid < C::PT) {
      int64_t p = p0 + tid;
      float pt[3] = {0.f, 0.f, 0.f};
      if (p < P) {
        if (x) { pt[0] = x[p * 3]; pt[1] = x[p * 3 + 1]; pt[2] = x[p * 3 + 2]; }
        else grid_point(grid_first + p, gridN, vs, pt);
      }
      xs[tid * 3] = pt[0]; xs[tid * 3 + 1] = pt[1]; xs[tid * 3 + 2] = pt[2];
    }
    __syncthreads();
    float acc[C::TN][C::CPT];
    // ---- layer 0 (K = 3) in registers ----
#pragma unroll
    for (int i = 0; i < C::TN; ++i) {
      const int n = ng + i * C::NG;
      const float w0x = net.W[0][n * 3], w0y = net.W[0][n * 3 + 1], w0z = net.W[0][n * 3 + 2], b0 = net.b[0][n];
#pragma unroll
      for (int pp = 0; pp < C::PPT; ++pp) {
        const float* pt = xs + (cg * C::PPT + pp) * 3;
        float* z = &acc[i][pp * NCH];
        z[0] = fmaf(w0z, pt[2], fmaf(w0y, pt[1], fmaf(w0x, pt[0], b0)));
        if constexpr (NCH >= 4) { z[1] = w0x; z[2] = w0y; z[3] = w0z; }
#pragma unroll
        for (int c = 4; c < NCH; ++c) z[c] = 0.f;
      }
    }
    for (int l = 0; l < L; ++l) {
      const float w = (l == 0) ? net.w0 : net.ww;
      if (l > 0) {
        simt_gemm<NCH>(net.Wt[l], As, Ws, acc, ng, cg);     // ends with __syncthreads: As is free
      }
#pragma unroll
      for (int i = 0; i < C::TN; ++i) {
        const int n = ng + i * C::NG;
        const float bias = (l > 0) ? net.b[l][n] : 0.f;
        float a[C::CPT];
#pragma unroll
        for (int pp = 0; pp < C::PPT; ++pp) {
          float* z = &acc[i][pp * NCH];
          z[0] += bias;
          float s, c;
          sincos_precise(w * z[0], s, c);
          sine_jet<NCH>(z, a + pp * NCH, w, s, c);
        }
        if constexpr (STASH) {
          const size_t off = ((size_t)l * 256 + n) * ctot + col0 + tile * C::NC + cg * C::CPT;
          store_cols<NCH>(Zst + off, &acc[i][0]);
          store_cols<NCH>(Ast + off, a);
        }
        store_cols<NCH>(As + n * C::S + cg * C::CPT, a);
      }
      __syncthreads();
    }
    // ---- output layer: one dot product of length 256 per column ----
    {
      const int warp = tid >> 5, lane = tid & 31;
      const float* Wl = net.W[L];
      for (int j = warp; j < C::NC; j += 8) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) sum = fmaf(Wl[lane + 32 * r], As[(lane + 32 * r) * C::S + j], sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) os[j] = sum + ((j % NCH == 0) ? net.b[L][0] : 0.f);
      }
    }
    __syncthreads();
    if (tid < C::PT && p0 + tid < P) finalize_point<NCH>(out, p0 + tid, os + tid * NCH);
  }
}

// =============================================================================================
// reverse sweep (data gradient chain); weight gradients of the 256x256 layers are contracted by
// simt_wgrad_kernel from the stashes written here and by the forward.
// =============================================================================================
template <int NCH>
__global__ void __launch_bounds__(256) simt_backward_kernel(NetView net, GradView grad, const float* __restrict__ x,
                                                            int64_t P, const float* __restrict__ seeds,
                                                            const float* __restrict__ Zst, float* __restrict__ Zbst,
                                                            int64_t ctot, int64_t col0) {
  using C = Cfg<NCH>;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Ws = As + 256 * C::S;
  float* xs = Ws + 2 * C::KC * 256;
  float* sd = xs + C::PT * 3;                        // seeds of the tile [NC]
  __shared__ float red_b[DUDF_MAX_LAYERS][256];      // bias gradients accumulated over this CTA's tiles
  __shared__ float red_wl[256];                      // output-layer weight gradient
  __shared__ float red_w0[256][3];                   // first-layer weight gradient
  __shared__ float red_bl;
  const int tid = threadIdx.x;
  const int ng = tid % C::NG, cg = tid / C::NG;
  const int L = net.n_lin - 1;
  const int64_t ntiles = (P + C::PT - 1) / C::PT;
  for (int l = 0; l < L; ++l) red_b[l][tid] = 0.f;
  red_wl[tid] = 0.f;
  red_w0[tid][0] = red_w0[tid][1] = red_w0[tid][2] = 0.f;
  if (tid == 0) red_bl = 0.f;

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * C::PT;
    __syncthreads();
    if (tid < C::PT) {
      int64_t p = p0 + tid;
      float pt[3] = {0.f, 0.f, 0.f};
      if (p < P) { pt[0] = x[p * 3]; pt[1] = x[p * 3 + 1]; pt[2] = x[p * 3 + 2]; }
      xs[tid * 3] = pt[0]; xs[tid * 3 + 1] = pt[1]; xs[tid * 3 + 2] = pt[2];
    }
    if (tid < C::NC) {
      int64_t p = p0 + tid / NCH;
      float v = (p < P) ? seeds[p0 * NCH + tid] : 0.f;
      sd[tid] = v;
      if (tid % NCH == 0) atomicAdd(&red_bl, v);
    }
    __syncthreads();
    float acc[C::TN][C::CPT];
    {
      const float* Wl = net.W[L];
#pragma unroll
      for (int i = 0; i < C::TN; ++i) {
        const float wl = Wl[ng + i * C::NG];
#pragma unroll
        for (int j = 0; j < C::CPT; ++j) acc[i][j] = wl * sd[cg * C::CPT + j];
      }
    }
    for (int l = L - 1; l >= 0; --l) {
      const float w = (l == 0) ? net.w0 : net.ww;
#pragma unroll
      for (int i = 0; i < C::TN; ++i) {
        const int n = ng + i * C::NG;
        const size_t off = ((size_t)l * 256 + n) * ctot + col0 + tile * C::NC + cg * C::CPT;
        float z[C::CPT], zb[C::CPT];
        load_cols<NCH>(Zst + off, z);
        float bsum = 0.f, wlsum = 0.f, w0s[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int pp = 0; pp < C::PPT; ++pp) {
          float s, c;
          sincos_precise(w * z[pp * NCH], s, c);
          if (l == L - 1) {
            float a[NCH];
            sine_jet<NCH>(z + pp * NCH, a, w, s, c);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) wlsum = fmaf(sd[cg * C::CPT + pp * NCH + ch], a[ch], wlsum);
          }
          sine_jet_adjoint<NCH>(z + pp * NCH, &acc[i][pp * NCH], zb + pp * NCH, w, s, c);
          bsum += zb[pp * NCH];
          if (l == 0) {
            const float* pt = xs + (cg * C::PPT + pp) * 3;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              float t = zb[pp * NCH] * pt[d];
              if constexpr (NCH >= 4) t += zb[pp * NCH + 1 + d];
              w0s[d] += t;
            }
          }
        }
        atomicAdd(&red_b[l][n], bsum);
        if (l == L - 1) atomicAdd(&red_wl[n], wlsum);
        if (l == 0) {
#pragma unroll
          for (int d = 0; d < 3; ++d) atomicAdd(&red_w0[n][d], w0s[d]);
        } else {
          store_cols<NCH>(Zbst + off, zb);
          store_cols<NCH>(As + n * C::S + cg * C::CPT, zb);
        }
      }
      if (l > 0) {
        __syncthreads();
        simt_gemm<NCH>(net.W[l], As, Ws, acc, ng, cg);      // abar_{l-1}[k] = sum_n W_l[n][k] zbar_l[n]
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < L; ++l) atomicAdd(&grad.b[l][tid], red_b[l][tid]);
  atomicAdd(&grad.W[L][tid], red_wl[tid]);
#pragma unroll
  for (int d = 0; d < 3; ++d) atomicAdd(&grad.W[0][tid * 3 + d], red_w0[tid][d]);
  if (tid == 0) atomicAdd(&grad.b[L][0], red_bl);
}

// =============================================================================================
// weight gradient of the hidden 256x256 layers: Wbar_l[n][k] += sum_col Zbar_l[n][col] A_{l-1}[k][col]
// grid = (4 output tiles of 128x128, L-1 layers, splits over columns)
// =============================================================================================
__global__ void __launch_bounds__(256) simt_wgrad_kernel(GradView grad, const float* __restrict__ Zbst,
                                                         const float* __restrict__ Ast, int64_t ctot, int64_t ncols) {
  constexpr int KC = 16, SS = 128 + 4;
  __shared__ __align__(16) float Zs[KC][SS];
  __shared__ __align__(16) float Bs[KC][SS];
  const int l = blockIdx.y + 1;
  const int n0 = (blockIdx.x >> 1) * 128, k0 = (blockIdx.x & 1) * 128;
  const int tid = threadIdx.x, tn = tid & 15, tk = tid >> 4;
  const float* Zsrc = Zbst + ((size_t)l * 256 + n0) * ctot;
  const float* Asrc = Ast + ((size_t)(l - 1) * 256 + k0) * ctot;
  const int64_t per = ((ncols + gridDim.z - 1) / gridDim.z + KC - 1) / KC * KC;
  const int64_t c_begin = (int64_t)blockIdx.z * per;
  const int64_t c_end = min(ncols, c_begin + per);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 1, lc = (tid & 1) * 8;       // loader: row lr, 8 columns starting at lc
  for (int64_t c = c_begin; c < c_end; c += KC) {
    float zv[8], av[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      int64_t col = c + lc + q;
      bool ok = col < c_end;
      zv[q] = ok ? Zsrc[(size_t)lr * ctot + col] : 0.f;
      av[q] = ok ? Asrc[(size_t)lr * ctot + col] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) { Zs[lc + q][lr] = zv[q]; Bs[lc + q][lr] = av[q]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      float zr[8], ar[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { zr[i] = Zs[kk][tn + 16 * i]; ar[i] = Bs[kk][tk + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(zr[i], ar[j], acc[i][j]);
    }
  }
  float* dst = grad.W[l];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dst[(size_t)(n0 + tn + 16 * i) * 256 + k0 + tk + 16 * j], acc[i][j]);
}

// =============================================================================================
// launchers
// =============================================================================================
template <int NCH, bool STASH>
static int launch_forward_t(const NetView& net, const float* x, int64_t P, int gridN, int64_t grid_first,
                            const QueryOut& out, float* Zst, float* Ast, int64_t ctot, int64_t col0, int sms,
                            cudaStream_t st) {
  using C = Cfg<NCH>;
  auto k = simt_forward_kernel<NCH, STASH>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
  int64_t ntiles = (P + C::PT - 1) / C::PT;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)sms * 2);
  if (grid < 1) return 0;
  k<<<grid, 256, C::smem_bytes, st>>>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0);
  DUDF_LAUNCH_OK();
  return 0;
}

int simt_forward(const NetView& net, int nch, const float* x, int64_t P, int gridN, int64_t grid_first,
                 const QueryOut& out, float* Zst, float* Ast, int64_t ctot, int64_t col0, int sms, cudaStream_t st) {
  const bool stash = Zst != nullptr;
  switch (nch) {
    case 1: return stash ? launch_forward_t<1, true>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st)
                         : launch_forward_t<1, false>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st);
    case 4: return stash ? launch_forward_t<4, true>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st)
                         : launch_forward_t<4, false>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st);
    case 10: return stash ? launch_forward_t<10, true>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st)
                          : launch_forward_t<10, false>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st);
    case 20: DUDF_REQUIRE(!stash, "third-order jets are query-only");
             return launch_forward_t<20, false>(net, x, P, gridN, grid_first, out, Zst, Ast, ctot, col0, sms, st);
  }
  DUDF_REQUIRE(false, "unsupported channel count %d", nch);
}

template <int NCH>
static int launch_backward_t(const NetView& net, const GradView& grad, const float* x, int64_t P, const float* seeds,
                             const float* Zst, float* Zbst, int64_t ctot, int64_t col0, int sms, cudaStream_t st) {
  using C = Cfg<NCH>;
  auto k = simt_backward_kernel<NCH>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
  int64_t ntiles = (P + C::PT - 1) / C::PT;
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)sms);
  if (grid < 1) return 0;
  k<<<grid, 256, C::smem_bytes, st>>>(net, grad, x, P, seeds, Zst, Zbst, ctot, col0);
  DUDF_LAUNCH_OK();
  return 0;
}

int simt_backward(const NetView& net, const GradView& grad, int nch, const float* x, int64_t P, const float* seeds,
                  const float* Zst, float* Zbst, int64_t ctot, int64_t col0, int sms, cudaStream_t st) {
  switch (nch) {
    case 1: return launch_backward_t<1>(net, grad, x, P, seeds, Zst, Zbst, ctot, col0, sms, st);
    case 4: return launch_backward_t<4>(net, grad, x, P, seeds, Zst, Zbst, ctot, col0, sms, st);
    case 10: return launch_backward_t<10>(net, grad, x, P, seeds, Zst, Zbst, ctot, col0, sms, st);
  }
  DUDF_REQUIRE(false, "unsupported channel count %d for the reverse sweep", nch);
}

int simt_wgrad(const NetView& net, const GradView& grad, const float* Zbst, const float* Ast, int64_t ctot,
               int64_t ncols, int sms, cudaStream_t st) {
  const int L = net.n_lin - 1;
  if (L < 2 || ncols <= 0) return 0;
  int splits = std::max(1, std::min(64, (int)((int64_t)sms * 4 / (4 * (L - 1)))));
  splits = (int)std::min<int64_t>(splits, (ncols + 255) / 256);
  dim3 grid(4, L - 1, splits);
  simt_wgrad_kernel<<<grid, 256, 0, st>>>(grad, Zbst, Ast, ctot, ncols);
  DUDF_LAUNCH_OK();
  return 0;
}

int simt_tile_points(int nch) { return nch == 1 ? Cfg<1>::PT : nch == 4 ? Cfg<4>::PT : nch == 10 ? Cfg<10>::PT : Cfg<20>::PT; }
int simt_tile_cols(int nch) { return nch == 1 ? Cfg<1>::NC : nch == 4 ? Cfg<4>::NC : nch == 10 ? Cfg<10>::NC : Cfg<20>::NC; }

}  // namespace dudf

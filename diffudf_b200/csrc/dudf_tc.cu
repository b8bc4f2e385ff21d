// Tensor-core ("TC16") field queries: the fused SIREN chain on tcgen05.mma with TMEM accumulators.
//
// Orientation.  For every hidden layer the kernel computes  D[neuron][column] = (w W)[neuron][k] * A[k][column]
// where a column is one jet channel of one point.  The weights are the UMMA "A" operand (K-major, M = 128
// neurons per instruction, two halves per layer), the activations the "B" operand, stored MN-major
// (column-contiguous rows of one input neuron each) in 128-byte-swizzled shared memory.  The accumulator
// lives in TMEM with lane = neuron, so one epilogue thread owns ALL channels of ALL points of its neuron:
// the sine-jet (which mixes the channels of a point) is thread-local, and the thread writes its row of the
// next layer's B operand with 16-byte stores of eight packed halves.  Activations never leave the SM.
//
// A sub-tile is 128 columns: 128 points (value only), 32 points x 4 channels, or 12 points x 10 channels
// (+ 8 idle columns).  Second-order channels are carried scaled by KAPPA = 1/8 (fp16 range head-room).
//
// Pipeline per CTA (persistent, one CTA per SM, 384 threads = 3 warpgroups; setmaxnreg moves registers to the epilogue):
//   warps 0-7  epilogue: TMEM -> registers -> sin/cos jet -> fp16 -> swizzled smem (+ first and last layer)
//   warp  8    producer: streams 16 KB weight chunks (128 neurons x 64 k) global/L2 -> smem ring with
//              cp.async.bulk (TMA engine) completing on mbarriers
//   warp  9    MMA issuer: the whole warp runs warp-uniform code, one elected lane issues tcgen05.mma; tcgen05.commit
//              frees ring slots / publishes accumulators
// Two sub-tiles are in flight so that the MMAs of one overlap the epilogue of the other.
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"
#include "dudf_umma.cuh"
#include "dudf_tc_common.cuh"

namespace dudf {

using namespace umma;

size_t tc_packed_bytes(int n_lin) { return (size_t)((n_lin > 2) ? (n_lin - 2) : 1) * 2 * 8 * TC_CHUNK_BYTES; }

// ---- weight packing: fp16(ww * W) in ready-to-copy swizzled chunks [layer][half][kblock]; the images of
// ---- the transposed matrices (A operand of the reverse sweep) follow the forward ones ----
__global__ void __launch_bounds__(256) tc_pack_kernel(NetView net, unsigned char* packed) {
  const int l = blockIdx.y + 1;                       // linear layer index 1 .. n_lin-2
  const int n = blockIdx.x;                           // output neuron
  const int k = threadIdx.x;                          // input neuron
  const __half v = __float2half_rn(net.ww * net.W[l][n * 256 + k]);
  {
    const int h = n >> 7, r = n & 127, kb = k >> 6, kk = k & 63;
    unsigned char* chunk = packed + ((size_t)(l - 1) * 8 + h * 4 + kb) * TC_CHUNK_BYTES;
    *reinterpret_cast<__half*>(chunk + sw128_offset(r, kk)) = v;
  }
  {
    const int h = k >> 7, r = k & 127, kb = n >> 6, kk = n & 63;
    unsigned char* chunk = packed + ((size_t)(net.n_lin - 2 + l - 1) * 8 + h * 4 + kb) * TC_CHUNK_BYTES;
    *reinterpret_cast<__half*>(chunk + sw128_offset(r, kk)) = v;
  }
}

int tc_pack(const NetView& net, void* packed, cudaStream_t st) {
  if (net.n_lin <= 2) return 0;
  tc_pack_kernel<<<dim3(256, net.n_lin - 2), 256, 0, st>>>(net, (unsigned char*)packed);
  DUDF_LAUNCH_OK();
  return 0;
}

// CL = thread-block cluster size: the CTAs of a cluster share every weight chunk fetched from L2 (multicast), which
// divides the L2 -> SM weight traffic (the measured limiter of this kernel at CL = 1, tools/pipe_probe.py) by CL.
template <int NCH, int CL, bool RU>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_forward_kernel(const unsigned char* __restrict__ packed, NetView net, const float* __restrict__ x, int64_t P, int gridN,
                  int64_t grid_first, QueryOut out) {
  using C = TcCfg<NCH>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* act = smem;
  unsigned char* ring = smem + C::OFF_RING;
  float* wl_s = (float*)(smem + C::OFF_WL);
  float* xs = (float*)(smem + C::OFF_XS);
  float* os = (float*)(smem + C::OFF_OS);
  uint64_t* bars = (uint64_t*)(smem + C::OFF_BAR);
  uint64_t *full = bars, *empty = bars + TC_STAGES, *act_ready = bars + 2 * TC_STAGES, *acc_ready = bars + 2 * TC_STAGES + 2;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TC_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = net.n_lin - 1;            // sine layers
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  // every CTA of a cluster runs the same number of rounds (a round past the end works on fully masked points)
  const int64_t rounds = (CL > 1) ? (npairs + gridDim.x - 1) / gridDim.x
                                  : (((int64_t)blockIdx.x < npairs) ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (tid == 0) {
    for (int i = 0; i < TC_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&act_ready[s], 8); mbar_init(&acc_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<512>(tmem_slot);
  if (tid < 256) wl_s[tid] = net.W[L][tid];
  if (C::NV < 128 && tid < 256) {         // idle columns of both tiles stay zero for the whole kernel
    for (int s = 0; s < 2; ++s) *reinterpret_cast<uint4*>(tc_tile_row(act + s * TC_ACT_BYTES, tid) + tc_chunk_off(15, tid & 7)) = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();                 // peers' barriers are initialised before anything is multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    setmaxnreg_dec<TC_REGS_AUX>();                          // warpgroup 2 hands its registers to the epilogue warpgroups
    if (warp == 8) {
      if (lane == 0) tc_producer<CL, RU>(packed, ring, full, empty, rounds, L - 1, TC_DIR_FWD, CL > 1 ? cluster_ctarank() : 0u, CL > 1 ? 0 : (int)(out.flags >> 8));
    } else if (warp == 9) {
      tc_mma_role<CL, RU>(act, ring, full, empty, act_ready, acc_ready, tmem_base, rounds, L - 1, nullptr, nullptr, 0, 0, TC_DIR_FWD, out.flags >> 8, 0,
                          (out.trace && blockIdx.x == 0) ? out.trace : nullptr);
    }
  } else {
    setmaxnreg_inc<TC_REGS_EPI>();
    // ===================== epilogue warps (256 threads) =====================
    const int q = warp & 3, h = warp >> 2;
    const int n = h * 128 + q * 32 + lane;                  // this thread's neuron = its row of the B operand
    const uint32_t r7 = n & 7;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * 128;
    const float w0 = net.w0, ww = net.ww;
    const float r0x = net.W[0][n * 3], r0y = net.W[0][n * 3 + 1], r0z = net.W[0][n * 3 + 2], b0 = net.b[0][n];
    const float bL = net.b[L][0];
    const float vs = gridN > 1 ? 2.0f / (float)(gridN - 1) : 0.f;
    uint32_t acc_phase = 0;                                 // bit s = parity of acc_ready[s]
    const bool inplace = ((out.flags >> 8) & 64) == 0;      // diagnostics bit 6: prefetch the accumulators a group ahead
    unsigned long long* etrace = (out.trace && blockIdx.x == 0 && warp == 0) ? out.trace + TC_TRACE_REGION : nullptr;
    uint32_t en = 0;

    for (int64_t rd = 0; rd < rounds; ++rd) {
      const int64_t pair = blockIdx.x + rd * gridDim.x;
      // ---- coordinates of both sub-tiles ----
      tc_trace(etrace, en, 14, 0);
      tc_epi_bar();
      for (int i = tid; i < 2 * C::PT; i += 256) {
        const int64_t p = pair * 2 * C::PT + i;
        float pt[3] = {0.f, 0.f, 0.f};
        if (p < P) {
          if (x) { pt[0] = x[p * 3]; pt[1] = x[p * 3 + 1]; pt[2] = x[p * 3 + 2]; }
          else grid_point(grid_first + p, gridN, vs, pt);
        }
        xs[i * 3] = pt[0]; xs[i * 3 + 1] = pt[1]; xs[i * 3 + 2] = pt[2];
      }
      tc_epi_bar();

      for (int l = 0; l < L; ++l) {
        const float bias = (l > 0) ? ww * net.b[l][n] : 0.f;
        for (int s = 0; s < 2; ++s) {
          unsigned char* trow = tc_tile_row(act + s * TC_ACT_BYTES, n);
          if (l > 0) {
            mbar_wait(&acc_ready[s], (acc_phase >> s) & 1u, 0x400 + s);
            acc_phase ^= 1u << s;
            tc_fence_after();
          }
          tc_trace(etrace, en, 10 + s, l);
          if (l == 0) {
            // first layer in fp32 on CUDA cores: u0 = w0 (W0 x + b0); derivative channels are w0 W0[:, i]
#pragma unroll 1
            for (int g = 0; g < (((out.flags >> 8) & 8) ? 0 : C::NGRP); ++g) {      // flags bit 11: diagnostics, skip
              float u[C::GC];
              tc_first_layer_group<NCH, C::GC>(u, xs + (s * C::PT + g * (C::GC / NCH)) * 3, w0, r0x, r0y, r0z, b0);
              tc_emit_group<NCH, C::GC, true>(u, trow, g * (C::GC / 8), r7);
            }
          } else if (!((out.flags >> 8) & 1)) {       // flags bit 8: pipeline diagnostics — skip the epilogue math
            TmemRegs<C::GC> nxt;
            if (inplace) {
              // accumulators loaded and used in place: a TMEM load takes tens of clocks, while prefetching a group ahead makes
              // ptxas copy the 32 / 40 staging registers of every group (DUDF_TC_INPLACE=0 restores the prefetch)
#pragma unroll 1
              for (int g = 0; g < C::NGRP; ++g) {
                float u[C::GC];
                tc_ld_issue<C::GC>(tmem_lane + s * 256 + g * C::GC, nxt);
                tc_ld_take<C::GC>(nxt, u);
#pragma unroll
                for (int pp = 0; pp < C::GC / NCH; ++pp) u[pp * NCH] += bias;
                tc_emit_group<NCH, C::GC, false>(u, trow, g * (C::GC / 8), r7);
              }
            } else {
              tc_ld_issue<C::GC>(tmem_lane + s * 256, nxt);
#pragma unroll 1
              for (int g = 0; g < C::NGRP; ++g) {
                float u[C::GC];
                tc_ld_take<C::GC>(nxt, u);
                if (g + 1 < C::NGRP) tc_ld_issue<C::GC>(tmem_lane + s * 256 + (g + 1) * C::GC, nxt);   // in flight during the math below
#pragma unroll
                for (int pp = 0; pp < C::GC / NCH; ++pp) u[pp * NCH] += bias;
                tc_emit_group<NCH, C::GC, false>(u, trow, g * (C::GC / 8), r7);
              }
            }
          }
          tc_trace(etrace, en, 12 + s, l);
          if (l < L - 1) {
            // publish the activation tile to the MMA issuer
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[s]);
          } else if (!((out.flags >> 8) & 8)) {
            // ---- output layer (256 -> 1 per channel) from the fp16 tile, then per-point finalisation ----
            tc_epi_bar();
            tc_output_dot<C::NV>(act + s * TC_ACT_BYTES, wl_s, os + s * 256, tid);
            tc_epi_bar();
            if (tid < C::NV) {
              const int ch = tid % NCH;
              const float v = os[s * 256 + tid] + os[s * 256 + 128 + tid];
              os[s * 256 + tid] = (ch == 0) ? v + bL : (ch >= 4 ? v * TC_KAPPA_INV : v);
            }
            tc_epi_bar();
            if (tid < C::PT) {
              const int64_t p = (pair * 2 + s) * C::PT + tid;
              if (p < P) finalize_point<NCH>(out, p, os + s * 256 + tid * NCH);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();                 // nobody exits while a peer may still multicast into its ring
  if (warp == 9) tmem_dealloc<512>(tmem_base);
}

static unsigned long long* g_tc_trace = nullptr;
void tc_set_trace(unsigned long long* buf) { g_tc_trace = buf; }
unsigned long long* tc_get_trace() { return g_tc_trace; }

static int g_tc_cluster = -1;      // DUDF_TC_CLUSTER=1|2|4 overrides the cluster size of the query kernel (default 2)
static int tc_cluster_size() {
  if (g_tc_cluster < 0) {
    const char* e = getenv("DUDF_TC_CLUSTER");
    const int v = e ? atoi(e) : 2;
    g_tc_cluster = (v == 1 || v == 2 || v == 4) ? v : 2;
  }
  return g_tc_cluster;
}

template <int NCH, int CL, bool RU = true>
static int tc_launch(const void* packed, const NetView& net, const float* x, int64_t P, int gridN, int64_t first,
                     const QueryOut& out, int sms, cudaStream_t st) {
  using C = TcCfg<NCH>;
  auto k = tc_forward_kernel<NCH, CL, RU>;
  DUDF_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  int grid = (int)std::min<int64_t>(npairs, sms);
  if (grid < 1) return 0;
  if (CL > 1) grid = std::max(CL, (std::min<int>(sms, (int)((npairs + CL - 1) / CL * CL)) / CL) * CL);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  QueryOut o2 = out;
  o2.trace = g_tc_trace;
  DUDF_CUDA_OK(cudaLaunchKernelEx(&cfg, k, (const unsigned char*)packed, net, x, P, gridN, first, o2));
  DUDF_LAUNCH_OK();
  return 0;
}

template <int NCH>
static int tc_launch_cl(const void* packed, const NetView& net, const float* x, int64_t P, int gridN, int64_t first, const QueryOut& out,
                        int sms, cudaStream_t st) {
  using C = TcCfg<NCH>;
  const int64_t npairs = (P + 2 * C::PT - 1) / (2 * C::PT);
  int cl = tc_cluster_size();
  while (cl > 1 && npairs < 2 * cl) cl >>= 1;              // tiny queries: no point in pairing CTAs
  static const bool no_reuse = getenv("DUDF_TC_REUSE") && atoi(getenv("DUDF_TC_REUSE")) == 0;   // ping-pong order instead of chunk reuse
  if (no_reuse) {
    if (cl == 4) return tc_launch<NCH, 4, false>(packed, net, x, P, gridN, first, out, sms, st);
    if (cl == 2) return tc_launch<NCH, 2, false>(packed, net, x, P, gridN, first, out, sms, st);
    return tc_launch<NCH, 1, false>(packed, net, x, P, gridN, first, out, sms, st);
  }
  if (cl == 4) return tc_launch<NCH, 4>(packed, net, x, P, gridN, first, out, sms, st);
  if (cl == 2) return tc_launch<NCH, 2>(packed, net, x, P, gridN, first, out, sms, st);
  return tc_launch<NCH, 1>(packed, net, x, P, gridN, first, out, sms, st);
}

int tc_forward(const void* packed, const NetView& net, int nch, const float* x, int64_t P, int gridN, int64_t grid_first,
               const QueryOut& out, int sms, cudaStream_t st) {
  switch (nch) {
    case 1: return tc_launch_cl<1>(packed, net, x, P, gridN, grid_first, out, sms, st);
    case 4: return tc_launch_cl<4>(packed, net, x, P, gridN, grid_first, out, sms, st);
    case 10: return tc_launch_cl<10>(packed, net, x, P, gridN, grid_first, out, sms, st);
  }
  DUDF_REQUIRE(false, "tensor-core path: unsupported channel count %d", nch);
}

// =============================================================================================
// bring-up self-test: D[128][N] = A[128][256] * B[N][256]^T with host-built operand images
//   variant 0: K-major A and B, N = 128        variant 1: K-major, N = 80
//   variant 2: K-major A, MN-major B, N = 128  variant 3: MN-major A and B (wgrad-like), N = 128
// =============================================================================================
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const unsigned char* __restrict__ Aimg, const unsigned char* __restrict__ Bimg, int a_bytes, int b_bytes,
                   int N, int variant, float* __restrict__ D) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* As = smem;
  unsigned char* Bs = smem + 65536;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  for (int i = tid * 16; i < a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(As + i) = *reinterpret_cast<const uint4*>(Aimg + i);
  for (int i = tid * 16; i < b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(Bs + i) = *reinterpret_cast<const uint4*>(Bimg + i);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const int a_mn = (variant == 3), b_mn = (variant >= 2);
    const uint32_t idesc = make_idesc_f16(128, N, 0, a_mn, b_mn);
    const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
    for (int ks = 0; ks < 16; ++ks) {       // 16 steps of K = 16
      uint64_t ad, bd;
      if (!a_mn) ad = make_desc_sw128(a0 + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
      else       ad = make_desc_sw128(a0 + ks * 2048, 32768, 1024);          // [mblock(2)][kgroup(32)] atoms of 1 KB
      if (!b_mn) bd = make_desc_sw128(b0 + (ks >> 2) * (N * 128) + (ks & 3) * 32, 16, 1024);
      else       bd = make_desc_sw128(b0 + ks * 2048, 32768, 1024);
      mma_f16_ss(tmem_base, ad, bd, idesc, ks != 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0, 0x900);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    tmem_ld_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem_base);
}

int tc_selftest(int variant, float* max_err, cudaStream_t st) {
  DUDF_REQUIRE(variant >= 0 && variant <= 3, "selftest variant %d", variant);
  const int M = 128, K = 256, N = (variant == 1) ? 80 : 128;
  std::vector<float> A(M * K), B(N * K);
  uint32_t seed = 12345u + variant;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 9) & 0x7FFF) / 16384.0f - 1.0f; };
  for (auto& v : A) v = __half2float(__float2half_rn(rnd()));
  for (auto& v : B) v = __half2float(__float2half_rn(rnd()));
  std::vector<unsigned char> Ai(65536, 0), Bi(65536, 0);
  const bool a_mn = (variant == 3), b_mn = (variant >= 2);
  auto put = [](std::vector<unsigned char>& img, size_t off, float v) {
    __half hv = __float2half_rn(v);
    memcpy(&img[off], &hv, 2);
  };
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      size_t off;
      if (!a_mn) off = (size_t)(k >> 6) * 16384 + sw128_offset(m, k & 63);
      else       off = ((size_t)(m >> 6) * 32 + (k >> 3)) * 1024 + sw128_offset(k & 7, m & 63);
      put(Ai, off, A[m * K + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      size_t off;
      if (!b_mn) off = (size_t)(k >> 6) * (N * 128) + sw128_offset(n, k & 63);
      else       off = ((size_t)(n >> 6) * 32 + (k >> 3)) * 1024 + sw128_offset(k & 7, n & 63);
      put(Bi, off, B[n * K + k]);
    }
  unsigned char *dA = nullptr, *dB = nullptr;
  float* dD = nullptr;
  DUDF_CUDA_OK(cudaMalloc(&dA, 65536));
  DUDF_CUDA_OK(cudaMalloc(&dB, 65536));
  DUDF_CUDA_OK(cudaMalloc(&dD, M * N * sizeof(float)));
  DUDF_CUDA_OK(cudaMemcpy(dA, Ai.data(), 65536, cudaMemcpyHostToDevice));
  DUDF_CUDA_OK(cudaMemcpy(dB, Bi.data(), 65536, cudaMemcpyHostToDevice));
  DUDF_CUDA_OK(cudaMemset(dD, 0, M * N * sizeof(float)));
  const int smem = 2 * 65536 + 1024;
  DUDF_CUDA_OK(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_selftest_kernel<<<1, 128, smem, st>>>(dA, dB, 65536, 65536, N, variant, dD);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<float> D(M * N);
  DUDF_CUDA_OK(cudaMemcpy(D.data(), dD, M * N * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  double worst = 0.0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0.0;
      for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[n * K + k];
      const double e = fabs(ref - (double)D[m * N + n]);
      if (e > worst) worst = e;
    }
  *max_err = (float)worst;
  return 0;
}

// =============================================================================================
// tensor-pipe micro-benchmark (tools/umma_bench.py): one thread per CTA issues `iters` x 16 back-to-back
// tcgen05.mma (K = 16 each) on resident shared-memory operands, alternating between two accumulators, and reports
// clocks per instruction.  Results are meaningless numerically; only the issue/operand-fetch rate matters.
//   variant 0: A K-major, B MN-major, N = 128 (the chain kernels)   1: A, B K-major, N = 128
//   variant 2: A K-major, B MN-major, N = 256                        3: A, B K-major, N = 256
//   variant 4: A in TMEM, B K-major, N = 256                         5: A in TMEM, B K-major, N = 128
//   variant 6 / 7 / 8: as 0 with a tcgen05.commit (to a second mbarrier) after every 4 / 8 / 16 instructions
//   variant 9: as 0, consecutive instructions alternate between two accumulators; 10: between four
//   variant 11: as 0, issued from warp-uniform code (elect per instruction) with descriptors advanced by adds
// =============================================================================================
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) tc_mma_bench_kernel(int variant, int iters, float* __restrict__ clk_per_mma) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int commit_every = variant == 6 ? 4 : variant == 7 ? 8 : variant == 8 ? 16 : 0;
  const int alt = variant == 9 ? 1 : variant == 10 ? 3 : 0;
  const bool variant11 = (variant == 11);
  if (variant >= 6) variant = 0;
  for (int i = tid * 16; i < 196608; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (variant11 && warp == 0) {
    constexpr uint32_t idesc = make_idesc_f16(128, 128, 0, 0, 1);
    const uint64_t ad0 = make_desc_sw128(smem_u32(smem), 16, 1024), bd0 = make_desc_sw128(smem_u32(smem + 65536), 32768, 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (uint32_t)(it & 1) * 128;
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        mma_f16_ss_warp(d, desc_advance(ad0, (ks >> 2) * 16384 + (ks & 3) * 32), desc_advance(bd0, ks * 2048), idesc, ks != 0);
    }
    mma_commit_warp(&bar);
    mbar_wait(&bar, 0, 0xA00);
    const long long t1 = clock64();
    if (tid == 0) clk_per_mma[blockIdx.x] = (float)(t1 - t0) / (float)(iters * 16);
  } else if (!variant11 && tid == 0) {
    const int N = (variant == 2 || variant == 3 || variant == 4) ? 256 : 128;
    const bool b_mn = (variant == 0 || variant == 2), a_tmem = (variant >= 4);
    const uint32_t idesc = make_idesc_f16(128, N, 0, 0, b_mn ? 1 : 0);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem_base + (a_tmem ? 0 : (uint32_t)(it & 1) * (N == 256 ? 256 : 128));
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const uint64_t bd = b_mn ? make_desc_sw128(b0 + ks * 2048, 32768, 1024)
                                 : make_desc_sw128(b0 + (ks >> 2) * (N * 128) + (ks & 3) * 32, 16, 1024);
        if (a_tmem) mma_f16_ts(d, tmem_base + 256 + ks * 8, bd, idesc, ks != 0);
        else mma_f16_ss(alt ? tmem_base + (uint32_t)(ks & alt) * 128 : d, make_desc_sw128(a0 + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), bd, idesc,
                        alt ? (it | (ks > alt)) != 0 : ks != 0);
        if (commit_every && ((ks + 1) % commit_every) == 0) mma_commit(&bar2);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0, 0xA00);
    const long long t1 = clock64();
    clk_per_mma[blockIdx.x] = (float)(t1 - t0) / (float)(iters * 16);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base);
}

int tc_mma_bench(int variant, int ctas, int iters, float* clk_host, cudaStream_t st) {
  DUDF_REQUIRE(variant >= 0 && variant <= 11 && ctas >= 1 && ctas <= 1024 && iters >= 1, "umma bench: bad arguments");
  float* d = nullptr;
  DUDF_CUDA_OK(cudaMalloc(&d, ctas * sizeof(float)));
  const int smem = 196608 + 1024;
  DUDF_CUDA_OK(cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc_mma_bench_kernel<<<ctas, 128, smem, st>>>(variant, iters, d);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaStreamSynchronize(st));
  std::vector<float> h(ctas);
  DUDF_CUDA_OK(cudaMemcpy(h.data(), d, ctas * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d);
  double sum = 0.0;
  for (float v : h) sum += v;
  *clk_host = (float)(sum / ctas);
  return 0;
}

}  // namespace dudf

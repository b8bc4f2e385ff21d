// Per-row epilogues of the DUDF hot path: loss terms and their adjoint seeds, Adam, eigen-normals,
// curvature, the extract_fields fallback and small utilities.  All HBM-bound, one thread per row.
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_loss.cuh"

namespace dudf {

template <int N>
__device__ __forceinline__ void block_reduce_add(double (&v)[N], double* dst) {
  __shared__ double sh[N][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double x = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += sh[threadIdx.x][w];
    if (x != 0.0) atomicAdd(&dst[threadIdx.x], x);
  }
}

// ---------------------------------------------------------------------------------------------
// loss_s1 / loss_siren terms + seeds (src/loss_functions.py:123-155, :82-104); loss_s2 seeds (:106-121)
// seeds use the stored-variable convention: off-diagonal Hessian channels carry Hbar_ij + Hbar_ji.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_seed_kernel(LossArgs a) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  float smax = 0.f;
  if (p < a.P) {
    const int nch = a.nch;
    const float* v = a.packed + p * nch;
    const float f = v[0];
    const float d = a.dist[p];
    const bool on = (d == 0.f);
    float sd[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) sd[c] = 0.f;
    if (a.mode != DUDF_LOSS_S2) {
      LossRowCfg cfg;
      cfg.mode = a.mode; cfg.alpha = a.alpha; cfg.invP = 1.0f / (float)a.P_global;
#pragma unroll
      for (int k = 0; k < 4; ++k) { cfg.w[k] = a.w[k]; cfg.up[k] = a.upstream ? a.upstream[k] : 1.f; }
      float vv[10];
#pragma unroll
      for (int c = 0; c < 10; ++c) vv[c] = (c < nch) ? v[c] : 0.f;
      loss_row(cfg, nch, vv, d, a.normals ? a.normals + p * 3 : nullptr, t, sd);
    } else {  // S2: seeds from the global statistics
      if (on && a.s2_stats) {
        const double n = a.s2_stats[0], s1 = a.s2_stats[1], s2 = a.s2_stats[2];
        const double mean = s1 / n;
        const double var = (s2 - n * mean * mean) / (n - 1.0);
        const double sdv = sqrt(var);
        sd[0] = (float)((a.upstream ? a.upstream[0] : 1.f) * a.w[0] * sgnd(mean) / n + (a.upstream ? a.upstream[1] : 1.f) * a.w[1] * ((double)f - mean) / ((n - 1.0) * sdv));
      }
    }
    if (a.seeds) {
      float* dst = a.seeds + p * nch;
      for (int c = 0; c < nch; ++c) dst[c] = sd[c];
    }
    if (a.seed_absmax) {
      // magnitude of the stored seeds of the tensor-core reverse sweep (second-order channels carry 1/KAPPA = 8)
      for (int c = 0; c < nch; ++c) smax = fmaxf(smax, fabsf(sd[c]) * (c >= 4 ? 8.f : 1.f));
    }
  }
  if (a.seed_absmax) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) smax = fmaxf(smax, __shfl_xor_sync(0xffffffffu, smax, o));
    if ((threadIdx.x & 31) == 0 && smax > 0.f && isfinite(smax)) atomicMax(reinterpret_cast<int*>(a.seed_absmax), __float_as_int(smax));
  }
  if (a.terms && a.mode != DUDF_LOSS_S2) block_reduce_add<4>(t, a.terms);
}

int loss_seeds(const LossArgs& a, cudaStream_t st) {
  if (a.P <= 0) return 0;
  // one row per thread with a serial fp64 eigen-solve on the on-surface rows: the launch is as long as one thread's chain, so the rows
  // are cut into blocks of 64 (a 9 990-row segment: 157 blocks over the 148 SMs instead of 40)
  constexpr int LOSS_BLOCK = 64;
  const int64_t blocks = (a.P + LOSS_BLOCK - 1) / LOSS_BLOCK;
  loss_seed_kernel<<<(unsigned)blocks, LOSS_BLOCK, 0, st>>>(a);
  DUDF_LAUNCH_OK();
  return 0;
}

__global__ void __launch_bounds__(256) s2_stats_kernel(const float* packed, const float* dist, int64_t P, double* stats) {
  double t[3] = {0.0, 0.0, 0.0};
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    if (dist[p] == 0.f) {
      const double f = packed[p];
      t[0] += 1.0; t[1] += f; t[2] += f * f;
    }
  }
  block_reduce_add<3>(t, stats);
}

int loss_s2_stats(const float* packed, const float* dist, int64_t P, double* stats, cudaStream_t st) {
  if (P <= 0) return 0;
  const int blocks = (int)std::min<int64_t>((P + 255) / 256, 1024);
  s2_stats_kernel<<<blocks, 256, 0, st>>>(packed, dist, P, stats);
  DUDF_LAUNCH_OK();
  return 0;
}

__global__ void s2_finish_kernel(const double* stats, float w0, float w1, double* terms) {
  const double n = stats[0], mean = stats[1] / n;
  const double var = (stats[2] - n * mean * mean) / (n - 1.0);
  terms[0] = fabs(mean) * w0;
  terms[1] = sqrt(var) * w1;
  terms[2] = 0.0;
  terms[3] = 0.0;
}
int s2_finish(const double* stats, float w0, float w1, double* terms, cudaStream_t st) {
  s2_finish_kernel<<<1, 1, 0, st>>>(stats, w0, w1, terms);
  DUDF_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults used by train.py:334-337)
// ---------------------------------------------------------------------------------------------
// unsafe / skipped (optional): the fused tensor-core step scales its fp16 adjoints with the PREVIOUS step's seed magnitude; when
// this step's seeds outgrew that scale (scale_guard_kernel below, summed over the ranks by the gradient all-reduce) the gradient
// may hold saturated adjoints, so the update is skipped on every rank — what torch.cuda.amp.GradScaler does on overflow — and
// counted.  The next step's scale comes from this step's magnitude, so at most one step is lost per jump.
// one element of torch.optim.Adam.step (no weight decay, no amsgrad), every operation pinned (no contraction left to the compiler) so
// that the plain and the peer-reducing kernel give bit-identical parameters for the same summed gradient
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, float step_size, float bc2_sqrt, float b1, float b2,
                                            float eps) {
  const float mi = __fmaf_rn(b1, m, __fmul_rn(1.f - b1, g));                       // exp_avg.lerp_(grad, 1-beta1)
  const float vi = __fmaf_rn(b2, v, __fmul_rn(__fmul_rn(1.f - b2, g), g));         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
  m = mi;
  v = vi;
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);
  p = __fmaf_rn(-step_size, __fdiv_rn(mi, denom), p);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float step_size, float bc2_sqrt,
                                                   float b1, float b2, float eps, const float* __restrict__ unsafe,
                                                   long long* __restrict__ skipped) {
  if (unsafe && *unsafe != 0.f) {
    if (skipped && blockIdx.x == 0 && threadIdx.x == 0) *skipped += 1;
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    adam_update(p[i], m[i], v[i], g[i], step_size, bc2_sqrt, b1, b2, eps);
  }
}

int adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int64_t t,
              cudaStream_t st, const float* unsafe, long long* skipped) {
  if (n <= 0) return 0;
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  adam_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, (float)((double)lr / bc1), (float)sqrt(bc2), b1, b2, eps, unsafe, skipped);
  DUDF_LAUNCH_OK();
  return 0;
}

// The same update with its step-dependent scalars read from DEVICE memory, for a training step replayed as a CUDA graph (kernel
// arguments are frozen at capture): state[0] = learning rate (written by the host when the schedule changes it), the 64-bit step count
// behind it (state + 2, 8-byte aligned) and a block ticket (state + 4).  Every block forms lr / (1 - b1^t) and sqrt(1 - b2^t) for
// t = count + 1 in double, like adam_step does on the host; the last block to finish advances the count.
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, int64_t n, float* __restrict__ state, float b1, float b2,
                                                       float eps) {
  __shared__ float sh[2];
  long long* count = reinterpret_cast<long long*>(state + 2);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(state + 4);
  if (threadIdx.x == 0) {
    const double t = (double)(*count + 1);
    sh[0] = (float)((double)state[0] / (1.0 - pow((double)b1, t)));
    sh[1] = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float step_size = sh[0], bc2_sqrt = sh[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    adam_update(p[i], m[i], v[i], g[i], step_size, bc2_sqrt, b1, b2, eps);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {          // every block has read the count before the last one finishes
      *count += 1;
      *ticket = 0u;
    }
  }
}

int adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float b1, float b2, float eps, cudaStream_t st) {
  if (n <= 0) return 0;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  adam_dev_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n, state, b1, b2, eps);
  DUDF_LAUNCH_OK();
  return 0;
}

// Data-parallel Adam with the gradient all-reduce FUSED IN: every rank owns a gradient buffer in peer-mapped (symmetric) memory;
// after a cross-rank barrier each rank's kernel reads element i of ALL ranks' buffers over NVLink (fixed rank order 0..W-1, so
// every rank forms the bit-identical sum and the replicated parameters stay in lock-step), and applies the update — one pass
// over peer memory instead of an NCCL all-reduce (2 latency-bound ring / tree phases for 1.85 MB) followed by an Adam launch.
// The element behind the gradient (index n) is the "loss scale outgrown" flag of the fused step, summed the same way.
struct PeerGrads { const float* g[16]; int world; };
__global__ void __launch_bounds__(256) adam_peers_kernel(float* __restrict__ p, PeerGrads pg, float* __restrict__ m, float* __restrict__ v,
                                                         int64_t n, float step_size, float bc2_sqrt, float b1, float b2, float eps,
                                                         int guarded, long long* __restrict__ skipped, float* __restrict__ g_sum_out) {
  if (guarded) {
    float flag = 0.f;
    for (int r = 0; r < pg.world; ++r) flag += pg.g[r][n];
    if (flag != 0.f) {
      if (skipped && blockIdx.x == 0 && threadIdx.x == 0) *skipped += 1;
      return;
    }
  }
  const int64_t n4 = n >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4 + 1; i4 += (int64_t)gridDim.x * blockDim.x) {
    float gi[4] = {0.f, 0.f, 0.f, 0.f};
    const int64_t i0 = i4 * 4;
    const int cnt = (i4 < n4) ? 4 : (int)(n - i0);
    if (cnt <= 0) break;
    if (cnt == 4) {
      for (int r = 0; r < pg.world; ++r) {
        const float4 t = __ldcv(reinterpret_cast<const float4*>(pg.g[r] + i0));       // peer memory: never a stale cached line
        gi[0] += t.x; gi[1] += t.y; gi[2] += t.z; gi[3] += t.w;
      }
    } else {
      for (int r = 0; r < pg.world; ++r)
        for (int k = 0; k < cnt; ++k) gi[k] += __ldcv(pg.g[r] + i0 + k);
    }
    for (int k = 0; k < cnt; ++k) {
      const int64_t i = i0 + k;
      adam_update(p[i], m[i], v[i], gi[k], step_size, bc2_sqrt, b1, b2, eps);
      if (g_sum_out) g_sum_out[i] = gi[k];
    }
  }
}

int adam_step_peers(float* p, const float* const* peer_grads, int world, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                    int64_t t, int guarded, long long* skipped, float* g_sum_out, cudaStream_t st) {
  if (n <= 0) return 0;
  DUDF_REQUIRE(world >= 1 && world <= 16, "dudf_adam_step_peers: 1 .. 16 ranks (got %d)", world);
  PeerGrads pg;
  pg.world = world;
  for (int r = 0; r < 16; ++r) pg.g[r] = r < world ? peer_grads[r] : nullptr;
  for (int r = 0; r < world; ++r) DUDF_REQUIRE(pg.g[r] != nullptr && ((uintptr_t)pg.g[r] & 15) == 0, "dudf_adam_step_peers: peer buffer %d null or not 16-byte aligned", r);
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  const int blocks = (int)std::min<int64_t>((n / 4 + 256) / 256, 148 * 4);
  adam_peers_kernel<<<blocks, 256, 0, st>>>(p, pg, m, v, n, (float)((double)lr / bc1), (float)sqrt(bc2), b1, b2, eps, guarded, skipped, g_sum_out);
  DUDF_LAUNCH_OK();
  return 0;
}

// flag += 1 when the seeds of this step, under the loss scale derived from the previous step's magnitude, exceed `limit`
// (nominal range of S * max|seed| is (1024, 2048]; fp16 saturates at 65504)
__global__ void scale_guard_kernel(const float* __restrict__ amax_prev, const float* __restrict__ amax_next, float limit,
                                   float* __restrict__ flag) {
  const float mp = *amax_prev, mn = *amax_next;
  const float S = (mp > 0.f && isfinite(mp)) ? exp2f(floorf(log2f(2048.f / mp))) : 1.f;      // loss_scale_from()
  if (!(S * mn <= limit)) *flag += 1.f;                                                        // NaN / inf trip it too
}
int scale_guard(const float* amax_prev, const float* amax_next, float limit, float* flag, cudaStream_t st) {
  scale_guard_kernel<<<1, 1, 0, st>>>(amax_prev, amax_next, limit, flag);
  DUDF_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
__global__ void transpose256_kernel(const float* __restrict__ W, float* __restrict__ Wt) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) tile[r][threadIdx.x] = W[(by + r) * 256 + bx + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) Wt[(bx + r) * 256 + by + threadIdx.x] = tile[threadIdx.x][r];
}
int transpose256(const float* W, float* Wt, cudaStream_t st) {
  transpose256_kernel<<<dim3(8, 8), dim3(32, 8), 0, st>>>(W, Wt);
  DUDF_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------------------------------------
// eigen-normal / principal directions of a batch of Hessians
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void canon_sign(double V[3][3], int k) {
  int im = 0;
  double am = fabs(V[0][k]);
  for (int i = 1; i < 3; ++i)
    if (fabs(V[i][k]) > am) { am = fabs(V[i][k]); im = i; }
  if (V[im][k] < 0.0)
    for (int i = 0; i < 3; ++i) V[i][k] = -V[i][k];
}

__global__ void __launch_bounds__(256) eig_normals_kernel(const float* __restrict__ Hm, const float* __restrict__ ref,
                                                          int ref_mode, int64_t P, float* n, float* dirs, float* lamo) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double H[3][3], lam[3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = (double)Hm[p * 9 + i * 3 + j];
  eigh3<double>(H, lam, V);
  for (int k = 0; k < 3; ++k) canon_sign(V, k);
  if (ref_mode != 0 && ref) {
    const double dt = V[0][2] * ref[p * 3] + V[1][2] * ref[p * 3 + 1] + V[2][2] * ref[p * 3 + 2];
    const bool flip = (ref_mode == 1) ? (dt < 0.0) : (dt > 0.0);
    if (flip)
      for (int i = 0; i < 3; ++i) V[i][2] = -V[i][2];
  }
  for (int i = 0; i < 3; ++i) n[p * 3 + i] = (float)V[i][2];
  if (dirs)
    for (int i = 0; i < 3; ++i) { dirs[p * 6 + i * 2] = (float)V[i][0]; dirs[p * 6 + i * 2 + 1] = (float)V[i][1]; }
  if (lamo)
    for (int k = 0; k < 3; ++k) lamo[p * 3 + k] = (float)lam[k];
}

int eig_normals(const float* H, const float* ref_dir, int ref_mode, int64_t P, float* n, float* dirs, float* lam,
                cudaStream_t st) {
  if (P <= 0) return 0;
  eig_normals_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(H, ref_dir, ref_mode, P, n, dirs, lam);
  DUDF_LAUNCH_OK();
  return 0;
}

// mean / gaussian curvature from H and the symmetric third derivatives (src/render_st.py:42-55):
// dn_i/dx_k = sum_{j<2} v_j[i] (v_j^T T[:,:,k] n) / (lam_2 - lam_j)      (SURVEY.md §8 a-M)
__global__ void __launch_bounds__(256) curvature_kernel(const float* __restrict__ Hm, const float* __restrict__ Tm, int64_t P,
                                                        float* n, float* meanc, float* gauss, float* Jout) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double H[3][3], lam[3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) H[i][j] = (double)Hm[p * 9 + i * 3 + j];
  eigh3<double>(H, lam, V);
  for (int k = 0; k < 3; ++k) canon_sign(V, k);
  double T[3][3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      for (int c = 0; c < 3; ++c) T[a][b][c] = (double)Tm[p * 10 + sym3(a, b, c)];
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int j = 0; j < 2; ++j) {
    const double inv = 1.0 / (lam[2] - lam[j]);
    for (int k = 0; k < 3; ++k) {
      double c = 0.0;
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) c += V[a][j] * T[a][b][k] * V[b][2];
      c *= inv;
      for (int i = 0; i < 3; ++i) J[i][k] += V[i][j] * c;
    }
  }
  if (n)
    for (int i = 0; i < 3; ++i) n[p * 3 + i] = (float)V[i][2];
  if (Jout)
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) Jout[p * 9 + i * 3 + k] = (float)J[i][k];
  if (meanc) meanc[p] = (float)(0.5 * (J[0][0] + J[1][1] + J[2][2]));
  if (gauss) {
    // -det [[J, n], [n^T, 0]] by cofactor expansion along the last row
    const double nv[3] = {V[0][2], V[1][2], V[2][2]};
    double M[4][4];
    for (int i = 0; i < 3; ++i) {
      for (int k = 0; k < 3; ++k) M[i][k] = J[i][k];
      M[i][3] = nv[i];
      M[3][i] = nv[i];
    }
    M[3][3] = 0.0;
    double det = 0.0;
    for (int c = 0; c < 4; ++c) {
      double m3[3][3];
      for (int i = 0; i < 3; ++i) {
        int cc = 0;
        for (int k = 0; k < 4; ++k) {
          if (k == c) continue;
          m3[i][cc++] = M[i][k];
        }
      }
      const double d3 = m3[0][0] * (m3[1][1] * m3[2][2] - m3[1][2] * m3[2][1]) -
                        m3[0][1] * (m3[1][0] * m3[2][2] - m3[1][2] * m3[2][0]) +
                        m3[0][2] * (m3[1][0] * m3[2][1] - m3[1][1] * m3[2][0]);
      det += (((3 + c) & 1) ? -1.0 : 1.0) * M[3][c] * d3;
    }
    gauss[p] = (float)(-det);
  }
}

int curvature(const float* H, const float* T, int64_t P, float* n, float* mean, float* gauss, float* J, cudaStream_t st) {
  if (P <= 0) return 0;
  curvature_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(H, T, P, n, mean, gauss, J);
  DUDF_LAUNCH_OK();
  return 0;
}

// ---- mean curvature from the directional third-order jet (dudf_mean_curvature) ----
// dirs9[p] = (v_0, v_1, n): the three directions the directional jet of point p is taken along
__global__ void __launch_bounds__(256) dirs9_kernel(const float* __restrict__ n, const float* __restrict__ dirs6, int64_t P, float* __restrict__ d9) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  for (int i = 0; i < 3; ++i) {
    d9[p * 9 + i] = dirs6[p * 6 + i * 2];
    d9[p * 9 + 3 + i] = dirs6[p * 6 + i * 2 + 1];
    d9[p * 9 + 6 + i] = n[p * 3 + i];
  }
}
// mean curvature = tr(dn/dx) / 2 = (T(v_0, v_0, n) / (lam_2 - lam_0) + T(v_1, v_1, n) / (lam_2 - lam_1)) / 2
__global__ void __launch_bounds__(256) mean_dir3_kernel(const float* __restrict__ jet, const float* __restrict__ lam, int64_t P, float* __restrict__ mean) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const double l0 = lam[p * 3], l1 = lam[p * 3 + 1], l2 = lam[p * 3 + 2];
  mean[p] = (float)(0.5 * ((double)jet[p * 10 + 8] / (l2 - l0) + (double)jet[p * 10 + 9] / (l2 - l1)));
}
int dirs9(const float* n, const float* dirs6, int64_t P, float* d9, cudaStream_t st) {
  if (P <= 0) return 0;
  dirs9_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(n, dirs6, P, d9);
  DUDF_LAUNCH_OK();
  return 0;
}
int mean_dir3(const float* jet, const float* lam, int64_t P, float* mean, cudaStream_t st) {
  if (P <= 0) return 0;
  mean_dir3_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(jet, lam, P, mean);
  DUDF_LAUNCH_OK();
  return 0;
}

// extract_fields fallback (src/render_mc.py:77-93): where the normalised gradient has norm < 0.04
// (only possible when grad f == 0 exactly) use the sign-aligned top eigenvector of the Hessian.
__global__ void __launch_bounds__(256) field_vectors_kernel(const float* __restrict__ g, const float* __restrict__ Hm, int64_t P,
                                                            float* vecs) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float gx = g[p * 3], gy = g[p * 3 + 1], gz = g[p * 3 + 2];
  if (sqrtf(gx * gx + gy * gy + gz * gz) < 0.04f) {
    double H[3][3], lam[3], V[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) H[i][j] = (double)Hm[p * 9 + i * 3 + j];
    eigh3<double>(H, lam, V);
    canon_sign(V, 2);
    const double dt = gx * V[0][2] + gy * V[1][2] + gz * V[2][2];
    const double sg = dt < 0.0 ? -1.0 : 1.0;
    gx = (float)(sg * V[0][2]); gy = (float)(sg * V[1][2]); gz = (float)(sg * V[2][2]);
  }
  vecs[p * 3] = gx; vecs[p * 3 + 1] = gy; vecs[p * 3 + 2] = gz;
}

int field_vectors(const float* g, const float* H, int64_t P, float* vecs, cudaStream_t st) {
  if (P <= 0) return 0;
  field_vectors_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(g, H, P, vecs);
  DUDF_LAUNCH_OK();
  return 0;
}

__global__ void __launch_bounds__(256) f32_to_f64_kernel(const float* __restrict__ s, double* __restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (double)s[i];
}
int f32_to_f64(const float* src, double* dst, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  f32_to_f64_kernel<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(src, dst, n);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

// Shared definitions for the DUDF B200 kernels (SIREN 3 -> 256 x L -> 1, sine activations).
//
// Channel ("jet") conventions used by every kernel in this directory.  A point carries NCH
// channels, ordered
//   NCH = 1 : f
//   NCH = 4 : f, d_x, d_y, d_z
//   NCH = 10: ... + d_xx, d_xy, d_xz, d_yy, d_yz, d_zz
//   NCH = 20: ... + d_xxx, d_xxy, d_xxz, d_xyy, d_xyz, d_xzz, d_yyy, d_yyz, d_yzz, d_zzz
// The linear map acts on every channel alike (bias on channel 0 only); the sine mixes the
// channels of one point (SURVEY.md §8 a-M; reference: src/model.py:29-30 differentiated the
// way src/diff_operators.py:187-212 does through autograd).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <algorithm>

#define DUDF_WIDTH 256
#define DUDF_MAX_LAYERS 16          // linear layers (hidden + output)

struct NetView {
  const float* W[DUDF_MAX_LAYERS];  // nn.Linear layout [out][in], fp32
  const float* b[DUDF_MAX_LAYERS];
  const float* Wt[DUDF_MAX_LAYERS]; // transposed copies [in][out] of the 256x256 layers (SIMT forward)
  int n_lin;                        // number of linear layers = n_hidden + 1
  float w0;                         // omega of the first sine layer (reference: w0)
  float ww;                         // omega of the other sine layers (reference: ww)
};

struct GradView {
  float* W[DUDF_MAX_LAYERS];
  float* b[DUDF_MAX_LAYERS];
};

__host__ __device__ constexpr int sym2(int i, int j) {           // index into (xx,xy,xz,yy,yz,zz)
  return i <= j ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : sym2(j, i);
}
__host__ __device__ constexpr int sym3_sorted(int i, int j, int k) {
  // i<=j<=k ; order xxx,xxy,xxz,xyy,xyz,xzz,yyy,yyz,yzz,zzz
  return i == 0 ? (j == 0 ? k : (j == 1 ? 2 + k : 5)) : (i == 1 ? (j == 1 ? 5 + k : 8) : 9);
}
__host__ __device__ constexpr int sym3(int i, int j, int k) {
  int a = i, b = j, c = k, t = 0;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) { t = b; b = c; c = t; }
  if (a > b) { t = a; a = b; b = t; }
  return sym3_sorted(a, b, c);
}

// ---------------------------------------------------------------------------------------------
// sin/cos of a moderately large fp32 argument with explicit Cody-Waite reduction by pi/2
// (first-layer arguments reach |30 (W0 x + b0)| ~ 40 rad; SURVEY.md §7 "hard parts").
// Accuracy ~1 ulp-ish for |x| < 1e4; uses FMA only (no MUFU), so it is the "fp32 path".
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sincos_precise(float x, float& s, float& c) {
  sincosf(x, &s, &c);               // libdevice: Payne-Hanek for huge args, Cody-Waite otherwise
}

// Fast variant for the tensor-core path: one reduction by 2*pi (3 FMAs), then MUFU.SIN/COS on
// an argument in [-pi, pi] (abs. error ~ 4e-7, far below the fp16 operand rounding).
__device__ __forceinline__ void sincos_fast(float x, float& s, float& c) {
  const float inv2pi = 0.15915494309189535f;
  // round to nearest even by the 1.5 * 2^23 trick: two FADDs on the FMA pipe (FRND shares the MUFU pipe: 8 clk per warp,
  // tools/micro/pipe_rates.cu); identical to rintf for |x| < 2^22 * 2 pi
  float q = (x * inv2pi + 12582912.f) - 12582912.f;
  float r = fmaf(q, -6.2831854820251465f, x);        // 2*pi hi
  r = fmaf(q, 1.7484555e-7f, r);                     // -(2*pi lo): 2pi = 6.28318548 - 1.7484555e-7
  s = __sinf(r);
  c = __cosf(r);
}

// fp32-grade variant without MUFU (split-precision tensor-core path): Cody-Waite reduction by pi/2 (FMA, two-term), the
// published single-precision minimax polynomials on [-pi/4, pi/4] (Cephes sinf / cosf coefficients), quadrant by the low
// bits of the rounded quotient.  ~1 ulp for |x| < 1e3; ~22 FMA-pipe / ALU instructions.
__device__ __forceinline__ void sincos_poly(float x, float& s, float& c) {
  const float t = fmaf(x, 0.63661977236758134f, 12582912.f);       // 1.5 * 2^23: the quotient rounded to nearest in the low bits
  const int qi = __float_as_int(t);
  const float q = t - 12582912.f;
  float r = fmaf(q, -1.57079637050628662109375f, x);
  r = fmaf(q, 4.37113882867379e-8f, r);                              // pi/2 = 1.5707963705062866 - 4.37113882867379e-8
  const float r2 = r * r;
  float sp = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, r2, -1.6666654611e-1f);
  const float sr = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, r2, 4.166664568298827e-2f);
  const float cr = fmaf(cp * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const bool swap = (qi & 1) != 0;
  const float ss = swap ? cr : sr, cc = swap ? sr : cr;
  s = __int_as_float(__float_as_int(ss) ^ ((qi & 2) << 30));
  c = __int_as_float(__float_as_int(cc) ^ (((qi + 1) & 2) << 30));
}

// ---------------------------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations (ascending eigenvalues, like
// torch.linalg.eigh / np.linalg.eigh used at src/loss_functions.py:142, src/render_st.py:59,
// src/render_mc.py:77, src/render_pc.py:65).  Reads the LOWER triangle (LAPACK 'L' default).
// V[i][k] = i-th component of eigenvector k.  Eigenvector signs are arbitrary by nature.
// ---------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ inline void eigh3(const T H[3][3], T lam[3], T V[3][3]) {
  T a[3][3];
  a[0][0] = H[0][0]; a[1][1] = H[1][1]; a[2][2] = H[2][2];
  a[0][1] = a[1][0] = H[1][0];
  a[0][2] = a[2][0] = H[2][0];
  a[1][2] = a[2][1] = H[2][1];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? T(1) : T(0);
  for (int sweep = 0; sweep < 12; ++sweep) {
    T off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    T diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= diag * (sizeof(T) == 4 ? T(1e-15) : T(1e-32)) || off == T(0)) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      T apq = a[p][q];
      if (apq == T(0)) continue;
      T theta = (a[q][q] - a[p][p]) / (T(2) * apq);
      T t = (theta >= T(0) ? T(1) : T(-1)) / (fabs(theta) + sqrt(theta * theta + T(1)));
      T c = T(1) / sqrt(t * t + T(1));
      T s = t * c;
      const int r = 3 - p - q;
      T arp = a[r][p], arq = a[r][q];
      a[p][p] -= t * apq;
      a[q][q] += t * apq;
      a[p][q] = a[q][p] = T(0);
      a[r][p] = a[p][r] = c * arp - s * arq;
      a[r][q] = a[q][r] = s * arp + c * arq;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        T vip = V[i][p], viq = V[i][q];
        V[i][p] = c * vip - s * viq;
        V[i][q] = s * vip + c * viq;
      }
    }
  }
  lam[0] = a[0][0]; lam[1] = a[1][1]; lam[2] = a[2][2];
  // sort ascending (3-element network), permuting eigenvector columns
#define DUDF_SWAP_COL(x, y)                                    \
  if (lam[x] > lam[y]) {                                       \
    T tl = lam[x]; lam[x] = lam[y]; lam[y] = tl;               \
    for (int i = 0; i < 3; ++i) { T tv = V[i][x]; V[i][x] = V[i][y]; V[i][y] = tv; } \
  }
  DUDF_SWAP_COL(0, 1) DUDF_SWAP_COL(1, 2) DUDF_SWAP_COL(0, 1)
#undef DUDF_SWAP_COL
}

// inverse of the hyperbolic scaling d*tanh(alpha d) near the surface: src/inverses.py:18-19
__device__ __forceinline__ float inv_tanh_dev(float f, float alpha) {
  return (f < 1.0f / alpha) ? sqrtf(f / alpha) : f;
}
// src/inverses.py:21-22
__device__ __forceinline__ float inv_siren_dev(float f, float min_step) {
  return (f > 0.0f) ? f : min_step;
}

// error plumbing for the C ABI
void dudf_set_error(const char* fmt, ...);
#define DUDF_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      dudf_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)
// every kernel launch of this library goes through this: error check + launch counter (dudf_launch_count)
void dudf_count_launch();
#define DUDF_LAUNCH_OK()                    \
  do {                                      \
    DUDF_CUDA_OK(cudaGetLastError());       \
    dudf_count_launch();                    \
  } while (0)
#define DUDF_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      dudf_set_error(__VA_ARGS__);         \
      return 2;                            \
    }                                      \
  } while (0)

// Device-side batch sampler for oriented point clouds: the step immediately before the hot path.
// Restates sampleTrainingDataPC / shortestDistance of the reference (src/dataset.py:72-131): a batch is
// [n_on cloud rows | n_far uniform domain points | n_near = cloud rows displaced along their normal by N(0, sigma)],
// distances 0 | nearest-cloud-point distance | |offset|, normals of the cloud rows | 0 | 0.
// The reference materialises the (n_far x n_cloud) matrix |x|^2 - 2 p.x twice (P @ X^T and the repeat of |x|^2, ~8 GB
// at 9 990 x 100 000 in fp64); here it is a tiled running minimum in registers: the cloud is read once per 1 024
// queries from shared memory, nothing is materialised.
// Random draws: counter-based Philox4x32-10 keyed by (seed, batch), so a row's draw does not depend on launch
// geometry; every draw can instead be supplied by the caller (parity tests feed the reference's numpy / torch draws).
#include <cstdlib>

#include "dudf_common.cuh"
#include "dudf_kernels.h"

namespace dudf {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

enum { STREAM_ON = 0, STREAM_FAR = 1, STREAM_NEAR = 2 };
__device__ __forceinline__ uint4 draw(const SampleArgs& a, uint32_t stream, int64_t i) {
  return philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((uint64_t)i >> 32), stream, (uint32_t)a.batch),
                       make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32) ^ (uint32_t)(a.batch >> 32)));
}
__device__ __forceinline__ int64_t on_cloud_index(const SampleArgs& a, int64_t row) {
  if (a.on_idx) return a.on_idx[row];
  const uint4 r = draw(a, STREAM_ON, row);
  return (int64_t)((((uint64_t)r.x << 32) | r.y) % (uint64_t)a.n_surf);
}

__global__ void __launch_bounds__(256) sample_rows_kernel(SampleArgs a) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t P = a.n_on + a.n_far + a.n_near;
  if (p >= P) return;
  float x[3], n[3] = {0.f, 0.f, 0.f}, d = 0.f;
  if (p < a.n_on) {
    const int64_t idx = on_cloud_index(a, p);
#pragma unroll
    for (int k = 0; k < 3; ++k) { x[k] = a.surf_pts[idx * 3 + k]; n[k] = a.surf_nrm[idx * 3 + k]; }
  } else if (p < a.n_on + a.n_far) {
    const int64_t i = p - a.n_on;
    if (a.far_pts) {
#pragma unroll
      for (int k = 0; k < 3; ++k) x[k] = a.far_pts[i * 3 + k];
    } else {
      const uint4 r = draw(a, STREAM_FAR, i);
      x[0] = a.lo[0] + (a.hi[0] - a.lo[0]) * u01(r.x);
      x[1] = a.lo[1] + (a.hi[1] - a.lo[1]) * u01(r.y);
      x[2] = a.lo[2] + (a.hi[2] - a.lo[2]) * u01(r.z);
    }
  } else {
    const int64_t i = p - a.n_on - a.n_far;
    const uint4 r = draw(a, STREAM_NEAR, i);
    const int64_t j = a.near_idx ? a.near_idx[i] : (int64_t)((((uint64_t)r.x << 32) | r.y) % (uint64_t)a.n_on);
    float off;
    if (a.near_off) off = a.near_off[i];
    else off = a.sigma * sqrtf(-2.0f * __logf(1.0f - u01(r.z))) * __cosf(6.283185307179586f * u01(r.w));   // Box-Muller
    const int64_t idx = on_cloud_index(a, j);
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = fmaf(a.surf_nrm[idx * 3 + k], off, a.surf_pts[idx * 3 + k]);
    d = fabsf(off);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { a.coords[p * 3 + k] = x[k]; a.normals[p * 3 + k] = n[k]; }
  a.dist[p] = d;
}

// ---- brute-force nearest-cloud-point distance ----
// Pass 1 (all queries): the reference's expansion min_x(|x|^2 - 2 p.x) + |p|^2 (src/dataset.py:72-78; 3 FMA + 1 min per pair) as a
// running minimum in registers, 8 queries per thread, the cloud streamed through shared memory in tiles of 1 024 points, split over
// the cloud to fill the GPU, atomicMin on the bit pattern (non-negative floats order like unsigned integers).
// The expansion carries an absolute rounding error of ~2e-7 on d^2 (|p|^2 and |x|^2 ~ 1 cancel): harmless for far points, not near
// the cloud (sqrt(2e-7) = 4e-4, i.e. up to 5e-4 on a distance near 0 — exactly where loss_s1's tanh target is most sensitive).
// Pass 2 (queries whose d^2 came out below NN_NEAR, a few % of the far rows, compacted on the device): the same tiled scan of the
// WHOLE cloud in the difference form sum (p - x)^2, exact to fp32 rounding — the true nearest point need not sit in the tile that won
// the approximate pass.  Packed fp32 arithmetic (sub / mul / fma .f32x2, FMNMX3 over two cloud points): 3.5 instructions per pair.
constexpr int NN_QPT = 8;            // queries per thread
constexpr int NN_THREADS = 128;      // 1 024 queries per block
constexpr int NN_TILE = 1024;        // cloud points per shared-memory tile
constexpr float NN_NEAR = 1.0e-3f;   // d^2 below which a query is re-scanned exactly (above: |error of d| <= 2e-7 / 2d <= 3.2e-6, 1e-4 relative)

__device__ __forceinline__ float nn_min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__global__ void __launch_bounds__(NN_THREADS) nn_min_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ X, int64_t nx,
                                                            int64_t per_split, unsigned int* __restrict__ key) {
  __shared__ float4 tile[NN_TILE];                              // x, y, z, |x|^2
  const int64_t q0 = (int64_t)blockIdx.x * (NN_THREADS * NN_QPT);
  const int64_t x0 = (int64_t)blockIdx.y * per_split, x1 = min(nx, x0 + per_split);
  float px[NN_QPT], py[NN_QPT], pz[NN_QPT], m[NN_QPT];
#pragma unroll
  for (int k = 0; k < NN_QPT; ++k) {
    const int64_t i = q0 + k * NN_THREADS + threadIdx.x;
    px[k] = py[k] = pz[k] = 0.f;
    if (i < nq) { px[k] = -2.f * q[i * 3]; py[k] = -2.f * q[i * 3 + 1]; pz[k] = -2.f * q[i * 3 + 2]; }
    m[k] = 3.0e38f;
  }
  for (int64_t t0 = x0; t0 < x1; t0 += NN_TILE) {
    const int cnt = (int)min((int64_t)NN_TILE, x1 - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < NN_TILE; j += NN_THREADS) {
      float4 v = make_float4(0.f, 0.f, 0.f, 3.0e38f);          // padding never wins the minimum
      if (j < cnt) {
        const float a = X[(t0 + j) * 3], b = X[(t0 + j) * 3 + 1], c = X[(t0 + j) * 3 + 2];
        v = make_float4(a, b, c, fmaf(a, a, fmaf(b, b, c * c)));
      }
      tile[j] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < NN_TILE; j += 2) {                      // two cloud points per FMNMX3: 3.5 instead of 4 instructions per pair
      const float4 v = tile[j], w = tile[j + 1];
#pragma unroll
      for (int k = 0; k < NN_QPT; ++k)
        m[k] = nn_min3(m[k], fmaf(px[k], v.x, fmaf(py[k], v.y, fmaf(pz[k], v.z, v.w))), fmaf(px[k], w.x, fmaf(py[k], w.y, fmaf(pz[k], w.z, w.w))));
    }
  }
#pragma unroll
  for (int k = 0; k < NN_QPT; ++k) {
    const int64_t i = q0 + k * NN_THREADS + threadIdx.x;
    if (i < nq && x1 > x0) atomicMin(&key[i], __float_as_uint(fmaxf(m[k] + 0.25f * (px[k] * px[k] + py[k] * py[k] + pz[k] * pz[k]), 0.f)));
  }
}

__global__ void __launch_bounds__(256) nn_init_kernel(unsigned int* __restrict__ key, int64_t nq, unsigned int* __restrict__ near_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) key[i] = 0x7f7fffffu;
  if (i == 0) *near_count = 0u;
}

// queries near the cloud -> compact list (order irrelevant); their keys restart at +inf for the exact pass
__global__ void __launch_bounds__(256) nn_select_kernel(unsigned int* __restrict__ key, int64_t nq, unsigned int* __restrict__ near_count,
                                                        unsigned int* __restrict__ near_list) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq && __uint_as_float(key[i]) < NN_NEAR) {
    near_list[atomicAdd(near_count, 1u)] = (unsigned int)i;
    key[i] = 0x7f7fffffu;
  }
}

__device__ __forceinline__ unsigned long long nn_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long nn_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long nn_mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long nn_sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// |p - x|^2 of two queries (packed) against one cloud point (broadcast pairs), summed like fmaf(dx, dx, fmaf(dy, dy, dz dz))
__device__ __forceinline__ unsigned long long nn_dist2(unsigned long long px, unsigned long long py, unsigned long long pz, unsigned long long xx,
                                                       unsigned long long yy, unsigned long long zz) {
  const unsigned long long dx = nn_sub2(px, xx), dy = nn_sub2(py, yy), dz = nn_sub2(pz, zz);
  return nn_fma2(dx, dx, nn_fma2(dy, dy, nn_mul2(dz, dz)));
}

// exact pass over the compacted near queries: grid sized for the worst case, blocks beyond the list exit.  QPT = 2 queries per thread
// (256 per block): the list is short, so the work is cut finer than in pass 1 to keep every SM busy
template <int QPT>
__global__ void __launch_bounds__(NN_THREADS) nn_exact_kernel(const float* __restrict__ q, const unsigned int* __restrict__ near_count,
                                                              const unsigned int* __restrict__ near_list, const float* __restrict__ X, int64_t nx,
                                                              int64_t per_split, unsigned int* __restrict__ key) {
  __shared__ ulonglong2 tile[2 * NN_TILE];                      // per cloud point (x, x, y, y) and (z, z, -, -): broadcast operands
  const unsigned int n_near = *near_count;
  const int64_t q0 = (int64_t)blockIdx.x * (NN_THREADS * QPT);
  if (q0 >= (int64_t)n_near) return;
  const int64_t x0 = (int64_t)blockIdx.y * per_split, x1 = min(nx, x0 + per_split);
  unsigned long long pxx[QPT / 2], pyy[QPT / 2], pzz[QPT / 2];
  int64_t qi[QPT];
#pragma unroll
  for (int k = 0; k < QPT / 2; ++k) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t slot = q0 + (2 * k + h) * NN_THREADS + threadIdx.x;
      const int64_t i = slot < (int64_t)n_near ? (int64_t)near_list[slot] : -1;
      qi[2 * k + h] = i;
      c[h][0] = c[h][1] = c[h][2] = 0.f;
      if (i >= 0) { c[h][0] = q[i * 3]; c[h][1] = q[i * 3 + 1]; c[h][2] = q[i * 3 + 2]; }
    }
    pxx[k] = nn_pack2(c[0][0], c[1][0]);
    pyy[k] = nn_pack2(c[0][1], c[1][1]);
    pzz[k] = nn_pack2(c[0][2], c[1][2]);
  }
  float m[QPT];
#pragma unroll
  for (int k = 0; k < QPT; ++k) m[k] = 3.0e38f;
  for (int64_t t0 = x0; t0 < x1; t0 += NN_TILE) {
    const int cnt = (int)min((int64_t)NN_TILE, x1 - t0);
    __syncthreads();
    for (int j = threadIdx.x; j < NN_TILE; j += NN_THREADS) {
      float a = 1.0e18f, b = 1.0e18f, c = 1.0e18f;           // padding: 3e36 away, never wins the minimum
      if (j < cnt) { a = X[(t0 + j) * 3]; b = X[(t0 + j) * 3 + 1]; c = X[(t0 + j) * 3 + 2]; }
      tile[2 * j] = make_ulonglong2(nn_pack2(a, a), nn_pack2(b, b));
      tile[2 * j + 1] = make_ulonglong2(nn_pack2(c, c), 0ull);
    }
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < NN_TILE; j += 2) {
      const ulonglong2 a0 = tile[2 * j], a1 = tile[2 * j + 1], b0 = tile[2 * j + 2], b1 = tile[2 * j + 3];
#pragma unroll
      for (int k = 0; k < QPT / 2; ++k) {
        const unsigned long long ta = nn_dist2(pxx[k], pyy[k], pzz[k], a0.x, a0.y, a1.x);
        const unsigned long long tb = nn_dist2(pxx[k], pyy[k], pzz[k], b0.x, b0.y, b1.x);
        float ta0, ta1, tb0, tb1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(ta0), "=f"(ta1) : "l"(ta));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(tb0), "=f"(tb1) : "l"(tb));
        m[2 * k] = nn_min3(m[2 * k], ta0, tb0);
        m[2 * k + 1] = nn_min3(m[2 * k + 1], ta1, tb1);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < QPT; ++k)
    if (qi[k] >= 0 && x1 > x0) atomicMin(&key[qi[k]], __float_as_uint(m[k]));
}

__global__ void __launch_bounds__(256) nn_finish_kernel(const unsigned int* __restrict__ key, int64_t nq, float* __restrict__ dist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) dist[i] = sqrtf(__uint_as_float(key[i]));
}

int nn_distance(const float* q, int64_t nq, const float* X, int64_t nx, float* dist, int sms, cudaStream_t st) {
  if (nq <= 0) return 0;
  DUDF_REQUIRE(nx > 0, "nearest-point distance: empty cloud");
  const char* scan = getenv("DUDF_NN_SCAN");      // A/B switch of bench.py: force the tiled scan
  if (nx >= CLOUD_INDEX_MIN_POINTS && nx <= CLOUD_INDEX_MAX_POINTS && nq * 16 >= nx && !(scan && scan[0] == '1')) {
    // large cloud, no index supplied: sorting it once (~0.1 ms at 200 000 points) and walking the box hierarchy is cheaper than the scan
    void* index = nullptr;
    DUDF_CUDA_OK(cudaMallocAsync(&index, (size_t)cloud_index_bytes(nx), st));
    int rc = cloud_index_build(X, nx, index, st);
    if (rc == 0) rc = cloud_index_query(q, nq, index, nx, dist, st);
    DUDF_CUDA_OK(cudaFreeAsync(index, st));
    return rc;
  }
  DUDF_REQUIRE(nq < (int64_t)0xffffffffu, "nearest-point distance: at most 2^32 - 2 queries per call");
  unsigned int* ws = nullptr;                  // stream-ordered scratch: key [nq], near list [nq], near count [1]
  DUDF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), (size_t)(2 * nq + 1) * sizeof(unsigned int), st));
  unsigned int *key = ws, *near_list = ws + nq, *near_count = ws + 2 * nq;
  nn_init_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(key, nq, near_count);
  DUDF_LAUNCH_OK();
  const int64_t qblocks = (nq + NN_THREADS * NN_QPT - 1) / (NN_THREADS * NN_QPT);
  int64_t splits = std::max<int64_t>(1, (4 * (int64_t)sms + qblocks - 1) / qblocks);        // ~4 blocks per SM
  splits = std::min<int64_t>(splits, (nx + NN_TILE - 1) / NN_TILE);
  int64_t per = (nx + splits - 1) / splits;
  per = (per + NN_TILE - 1) / NN_TILE * NN_TILE;
  splits = (nx + per - 1) / per;
  nn_min_kernel<<<dim3((unsigned)qblocks, (unsigned)splits), NN_THREADS, 0, st>>>(q, nq, X, nx, per, key);
  DUDF_LAUNCH_OK();
  nn_select_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(key, nq, near_count, near_list);
  DUDF_LAUNCH_OK();
  constexpr int XQ = 2;
  nn_exact_kernel<XQ><<<dim3((unsigned)((nq + NN_THREADS * XQ - 1) / (NN_THREADS * XQ)), (unsigned)splits), NN_THREADS, 0, st>>>(q, near_count, near_list, X,
                                                                                                                            nx, per, key);
  DUDF_LAUNCH_OK();
  nn_finish_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(key, nq, dist);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaFreeAsync(ws, st));
  return 0;
}

int sample_rows(const SampleArgs& a, cudaStream_t st) {
  const int64_t P = a.n_on + a.n_far + a.n_near;
  if (P <= 0) return 0;
  sample_rows_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(a);
  DUDF_LAUNCH_OK();
  return 0;
}

int sample_batch_pc(const SampleArgs& a, int sms, cudaStream_t st) {
  const int64_t P = a.n_on + a.n_far + a.n_near;
  if (P <= 0) return 0;
  int rc = sample_rows(a, st);
  if (rc) return rc;
  return nn_distance(a.coords + a.n_on * 3, a.n_far, a.surf_pts, a.n_surf, a.dist + a.n_on, sms, st);
}

int sample_batch_pc_indexed(const SampleArgs& a, const void* index, cudaStream_t st) {
  const int64_t P = a.n_on + a.n_far + a.n_near;
  if (P <= 0) return 0;
  int rc = sample_rows(a, st);
  if (rc) return rc;
  return cloud_index_query(a.coords + a.n_on * 3, a.n_far, index, a.n_surf, a.dist + a.n_on, st);
}

}  // namespace dudf

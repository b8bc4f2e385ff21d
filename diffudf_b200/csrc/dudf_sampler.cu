// Device-side batch sampler for oriented point clouds: the step immediately before the hot path.
// Restates sampleTrainingDataPC / shortestDistance of the reference (src/dataset.py:72-131): a batch is
// [n_on cloud rows | n_far uniform domain points | n_near = cloud rows displaced along their normal by N(0, sigma)],
// distances 0 | nearest-cloud-point distance | |offset|, normals of the cloud rows | 0 | 0.
// The reference materialises the (n_far x n_cloud) matrix |x|^2 - 2 p.x twice (P @ X^T and the repeat of |x|^2, ~8 GB
// at 9 990 x 100 000 in fp64); here it is a tiled running minimum in registers: the cloud is read once per 1 024
// queries from shared memory, nothing is materialised.
// Random draws: counter-based Philox4x32-10 keyed by (seed, batch), so a row's draw does not depend on launch
// geometry; every draw can instead be supplied by the caller (parity tests feed the reference's numpy / torch draws).
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
// Main pass: the reference's expansion min_x(|x|^2 - 2 p.x) + |p|^2 (3 FMA + 1 min per pair) as a running minimum in
// registers, 8 queries per thread, the cloud streamed through shared memory in tiles of 1 024 points; the winning TILE is
// remembered.  The expansion carries an absolute rounding error of ~2e-7 on d^2 (|p|^2 and |x|^2 ~ 1 cancel): harmless
// for far points, not near the cloud (sqrt(2e-7) = 4e-4).  So the finish pass re-scans the winning tile in the
// difference form sum (p - x)^2 (exact to fp32 rounding) for every query whose d^2 came out below NN_NEAR.
constexpr int NN_QPT = 8;            // queries per thread
constexpr int NN_THREADS = 128;      // 1 024 queries per block
constexpr int NN_TILE = 1024;        // cloud points per shared-memory tile (float4: x, y, z, |x|^2)
constexpr float NN_NEAR = 4.0e-3f;   // d^2 below which the winning tile is re-scanned

__global__ void __launch_bounds__(NN_THREADS) nn_min_kernel(const float* __restrict__ q, int64_t nq, const float* __restrict__ X, int64_t nx,
                                                            int64_t per_split, unsigned long long* __restrict__ key) {
  __shared__ float4 tile[NN_TILE];
  const int64_t q0 = (int64_t)blockIdx.x * (NN_THREADS * NN_QPT);
  const int64_t x0 = (int64_t)blockIdx.y * per_split, x1 = min(nx, x0 + per_split);
  float px[NN_QPT], py[NN_QPT], pz[NN_QPT], m[NN_QPT];
  uint32_t best[NN_QPT];
#pragma unroll
  for (int k = 0; k < NN_QPT; ++k) {
    const int64_t i = q0 + k * NN_THREADS + threadIdx.x;
    px[k] = py[k] = pz[k] = 0.f;
    if (i < nq) { px[k] = -2.f * q[i * 3]; py[k] = -2.f * q[i * 3 + 1]; pz[k] = -2.f * q[i * 3 + 2]; }
    m[k] = 3.0e38f;
    best[k] = 0u;
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
    float mt[NN_QPT];
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k) mt[k] = 3.0e38f;
#pragma unroll 4
    for (int j = 0; j < NN_TILE; ++j) {
      const float4 v = tile[j];
#pragma unroll
      for (int k = 0; k < NN_QPT; ++k) mt[k] = fminf(mt[k], fmaf(px[k], v.x, fmaf(py[k], v.y, fmaf(pz[k], v.z, v.w))));
    }
    const uint32_t tidx = (uint32_t)(t0 / NN_TILE);
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k)
      if (mt[k] < m[k]) { m[k] = mt[k]; best[k] = tidx; }
  }
#pragma unroll
  for (int k = 0; k < NN_QPT; ++k) {
    const int64_t i = q0 + k * NN_THREADS + threadIdx.x;
    if (i < nq && x1 > x0) {
      const float d2 = fmaxf(m[k] + 0.25f * (px[k] * px[k] + py[k] * py[k] + pz[k] * pz[k]), 0.f);
      // non-negative floats order like their bit patterns; the tile index rides in the low word
      atomicMin(&key[i], ((unsigned long long)__float_as_uint(d2) << 32) | best[k]);
    }
  }
}

__global__ void __launch_bounds__(256) nn_init_kernel(unsigned long long* __restrict__ key, int64_t nq) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) key[i] = ~0ull;
}

// one warp per query: sqrt of the expansion value, or the exact re-scan of the winning tile near the cloud
__global__ void __launch_bounds__(256) nn_finish_kernel(const unsigned long long* __restrict__ key, const float* __restrict__ q, int64_t nq,
                                                        const float* __restrict__ X, int64_t nx, float* __restrict__ dist) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= nq) return;
  const unsigned long long kv = key[i];
  float d2 = __uint_as_float((uint32_t)(kv >> 32));
  if (d2 < NN_NEAR) {
    // near the cloud the expansion's ~2e-7 absolute error on d^2 is up to 5e-4 on d: re-scan in the difference form.  The WHOLE cloud,
    // not only the winning tile of the approximate pass — the true nearest point may sit in another tile whose approximate minimum
    // lost by less than that error (a few % of the far rows are this close; one warp each, ~0.03 ms for a 200 k-point cloud)
    const int64_t t0 = 0, t1 = nx;
    const float qx = q[i * 3], qy = q[i * 3 + 1], qz = q[i * 3 + 2];
    float e = 3.0e38f;
    for (int64_t j = t0 + lane; j < t1; j += 32) {
      const float dx = qx - X[j * 3], dy = qy - X[j * 3 + 1], dz = qz - X[j * 3 + 2];
      e = fminf(e, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e = fminf(e, __shfl_xor_sync(0xffffffffu, e, o));
    d2 = e;
  }
  if (lane == 0) dist[i] = sqrtf(d2);
}

int nn_distance(const float* q, int64_t nq, const float* X, int64_t nx, float* dist, int sms, cudaStream_t st) {
  if (nq <= 0) return 0;
  DUDF_REQUIRE(nx > 0, "nearest-point distance: empty cloud");
  unsigned long long* key = nullptr;
  DUDF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&key), (size_t)nq * sizeof(unsigned long long), st));   // stream-ordered scratch
  nn_init_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(key, nq);
  DUDF_LAUNCH_OK();
  const int64_t qblocks = (nq + NN_THREADS * NN_QPT - 1) / (NN_THREADS * NN_QPT);
  int64_t splits = std::max<int64_t>(1, (4 * (int64_t)sms + qblocks - 1) / qblocks);        // ~4 blocks per SM
  splits = std::min<int64_t>(splits, (nx + NN_TILE - 1) / NN_TILE);
  int64_t per = (nx + splits - 1) / splits;
  per = (per + NN_TILE - 1) / NN_TILE * NN_TILE;                                            // splits start on tile boundaries
  splits = (nx + per - 1) / per;
  nn_min_kernel<<<dim3((unsigned)qblocks, (unsigned)splits), NN_THREADS, 0, st>>>(q, nq, X, nx, per, key);
  DUDF_LAUNCH_OK();
  nn_finish_kernel<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, st>>>(key, q, nq, X, nx, dist);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaFreeAsync(key, st));
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

}  // namespace dudf

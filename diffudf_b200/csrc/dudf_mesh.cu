// Mesh half of the device-side batch sampler: sampleTrainingData of the reference
// (src/dataset.py:14-70) asks Open3D's RaycastingScene for the distance of every off-surface row to the triangle mesh, and
// preprocessMesh (src/preprocess_mesh.py:29-40) draws the surface cloud with mesh.sample_points_uniformly(use_triangle_normal).
// Open3D is not part of the reference's sources; what it computes is restated here:
//   * point-to-triangle distance by brute force — every query against every triangle, triangles staged through shared memory
//     in a precomputed edge form, the running minimum in registers (beetle: 19 980 rows x 2 053 faces = 41 M pair tests per
//     batch).  The distance is UNSIGNED: the losses are even in d (d tanh(alpha d), |t + alpha d (1 - t^2)|), and the sign
//     Open3D derives from ray parity is meaningless for the open surfaces DUDF is about;
//   * area-weighted surface samples: triangle by inverse CDF of the areas, barycentric (1 - sqrt r1, sqrt r1 (1 - r2), sqrt r1 r2),
//     the triangle's normal.  Draws come from Philox (or from the caller, for parity tests).
#include "dudf_common.cuh"
#include "dudf_kernels.h"

namespace dudf {

struct TriForm {          // 16 floats per triangle
  float ax, ay, az;       // vertex a
  float ux, uy, uz;       // ab = b - a
  float vx, vy, vz;       // ac = c - a
  float uu, uv, vv;       // ab.ab, ab.ac, ac.ac
  float pad[4];
};
constexpr int MT_TILE = 512;          // triangles per shared-memory tile (32 KB)
constexpr int MT_THREADS = 256;

__global__ void __launch_bounds__(256) tri_form_kernel(const float* __restrict__ tri, int64_t nt, TriForm* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const float* p = tri + t * 9;
  TriForm f;
  f.ax = p[0]; f.ay = p[1]; f.az = p[2];
  f.ux = p[3] - p[0]; f.uy = p[4] - p[1]; f.uz = p[5] - p[2];
  f.vx = p[6] - p[0]; f.vy = p[7] - p[1]; f.vz = p[8] - p[2];
  f.uu = f.ux * f.ux + f.uy * f.uy + f.uz * f.uz;
  f.uv = f.ux * f.vx + f.uy * f.vy + f.uz * f.vz;
  f.vv = f.vx * f.vx + f.vy * f.vy + f.vz * f.vz;
  f.pad[0] = f.pad[1] = f.pad[2] = f.pad[3] = 0.f;
  out[t] = f;
}

// squared distance from p to the triangle (a, a + u, a + v): minimise |w - s u - t v|^2 over the triangle in (s, t), w = p - a.
// Region logic on the unconstrained minimiser; the edge cases are 1-D clamped projections.  Degenerate triangles fall back
// to their edges (the edge projections guard their own zero lengths).
__device__ __forceinline__ float tri_dist2(const TriForm& f, float px, float py, float pz) {
  const float wx = px - f.ax, wy = py - f.ay, wz = pz - f.az;
  const float wu = wx * f.ux + wy * f.uy + wz * f.uz;
  const float wv = wx * f.vx + wy * f.vy + wz * f.vz;
  const float det = f.uu * f.vv - f.uv * f.uv;
  float s = f.vv * wu - f.uv * wv;      // unconstrained minimiser times det
  float t = f.uu * wv - f.uv * wu;
  auto clamp01 = [](float x) { return fminf(fmaxf(x, 0.f), 1.f); };
  // squared length of the residual w - s u - t v, formed explicitly: the expanded quadratic ww - 2(s wu + t wv) + ... cancels
  // to ~1e-7 |w|^2, which is 1e-4 absolute on a distance of 1e-3 (the near-surface rows of a batch)
  auto resid2 = [&](float s1, float t1) {
    const float rx = fmaf(-t1, f.vx, fmaf(-s1, f.ux, wx)), ry = fmaf(-t1, f.vy, fmaf(-s1, f.uy, wy)), rz = fmaf(-t1, f.vz, fmaf(-s1, f.uz, wz));
    return fmaf(rx, rx, fmaf(ry, ry, rz * rz));
  };
  float d2;
  if (det > 1e-30f && s >= 0.f && t >= 0.f && s + t <= det) {
    // interior: one step of iterative refinement on the 2 x 2 normal equations — for sliver triangles (uu vv / det ~ 1e3) the
    // first solve carries ~1e-4 relative error in (s, t), i.e. an in-plane foot-point error of ~3e-5, first order in d near 0
    const float inv = 1.f / det;
    float s1 = s * inv, t1 = t * inv;
    const float rx = fmaf(-t1, f.vx, fmaf(-s1, f.ux, wx)), ry = fmaf(-t1, f.vy, fmaf(-s1, f.uy, wy)), rz = fmaf(-t1, f.vz, fmaf(-s1, f.uz, wz));
    const float ru = rx * f.ux + ry * f.uy + rz * f.uz, rv = rx * f.vx + ry * f.vy + rz * f.vz;
    s1 += (f.vv * ru - f.uv * rv) * inv;
    t1 += (f.uu * rv - f.uv * ru) * inv;
    d2 = resid2(s1, t1);
  } else {
    // candidates on the three edges (clamped 1-D projections)
    const float kab = f.uu > 0.f ? clamp01(wu / f.uu) : 0.f;
    const float kac = f.vv > 0.f ? clamp01(wv / f.vv) : 0.f;
    const float ee = f.uu - 2.f * f.uv + f.vv;                  // b + k (c - b):  e = v - u,  (w - u).e / e.e
    const float we = (wv - wu) - (f.uv - f.uu);
    const float kbc = ee > 0.f ? clamp01(we / ee) : 0.f;
    d2 = fminf(resid2(kab, 0.f), fminf(resid2(0.f, kac), resid2(1.f - kbc, kbc)));
  }
  return fmaxf(d2, 0.f);
}

// grid: (query blocks, triangle splits); one query per thread; minimum over this split's triangles, merged with atomicMin on
// the bit pattern (non-negative floats order like unsigned integers)
__global__ void __launch_bounds__(MT_THREADS) tri_min_kernel(const float* __restrict__ q, int64_t nq, const TriForm* __restrict__ T, int64_t nt,
                                                             int64_t per_split, unsigned int* __restrict__ key) {
  __shared__ TriForm tile[MT_TILE];
  const int64_t i = (int64_t)blockIdx.x * MT_THREADS + threadIdx.x;
  const int64_t t0s = (int64_t)blockIdx.y * per_split, t1s = min(nt, t0s + per_split);
  float px = 0.f, py = 0.f, pz = 0.f;
  if (i < nq) { px = q[i * 3]; py = q[i * 3 + 1]; pz = q[i * 3 + 2]; }
  float m = 3.0e38f;
  for (int64_t t0 = t0s; t0 < t1s; t0 += MT_TILE) {
    const int cnt = (int)min((int64_t)MT_TILE, t1s - t0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(T + t0);
    float4* dst = reinterpret_cast<float4*>(tile);
    for (int j = threadIdx.x; j < cnt * 4; j += MT_THREADS) dst[j] = src[j];
    __syncthreads();
    for (int j = 0; j < cnt; ++j) m = fminf(m, tri_dist2(tile[j], px, py, pz));
  }
  if (i < nq && t1s > t0s) atomicMin(&key[i], __float_as_uint(m));
}

__global__ void __launch_bounds__(256) tri_finish_kernel(const unsigned int* __restrict__ key, int64_t nq, float* __restrict__ dist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq) dist[i] = sqrtf(__uint_as_float(key[i]));
}

int mesh_distance(const float* q, int64_t nq, const float* tri, int64_t nt, float* dist, int sms, cudaStream_t st) {
  if (nq <= 0) return 0;
  DUDF_REQUIRE(nt > 0, "mesh distance: empty triangle list");
  unsigned char* ws = nullptr;
  const size_t form_bytes = (size_t)nt * sizeof(TriForm);
  DUDF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), form_bytes + (size_t)nq * sizeof(unsigned int), st));
  TriForm* forms = reinterpret_cast<TriForm*>(ws);
  unsigned int* key = reinterpret_cast<unsigned int*>(ws + form_bytes);
  tri_form_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(tri, nt, forms);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaMemsetAsync(key, 0x7f, (size_t)nq * sizeof(unsigned int), st));       // 0x7f7f7f7f = 3.4e38
  const int64_t qblocks = (nq + MT_THREADS - 1) / MT_THREADS;
  int64_t splits = std::max<int64_t>(1, (4 * (int64_t)sms + qblocks - 1) / qblocks);
  splits = std::min<int64_t>(splits, (nt + MT_TILE - 1) / MT_TILE);
  int64_t per = (nt + splits - 1) / splits;
  per = (per + MT_TILE - 1) / MT_TILE * MT_TILE;
  splits = (nt + per - 1) / per;
  tri_min_kernel<<<dim3((unsigned)qblocks, (unsigned)splits), MT_THREADS, 0, st>>>(q, nq, forms, nt, per, key);
  DUDF_LAUNCH_OK();
  tri_finish_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(key, nq, dist);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cudaFreeAsync(ws, st));
  return 0;
}

// ---- area-weighted surface samples ----
__device__ __forceinline__ uint4 mesh_philox(uint4 c, uint2 k) {
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

// cdf: [nt] inclusive prefix sums of the triangle areas divided by the total (host-built, float64 -> float32).
// draws (optional): [n][3] = (u_triangle, r1, r2) in [0, 1)
__global__ void __launch_bounds__(256) mesh_sample_kernel(const float* __restrict__ tri, const float* __restrict__ cdf, int64_t nt, int64_t n,
                                                          uint64_t seed, const float* __restrict__ draws, float* __restrict__ pts,
                                                          float* __restrict__ nrm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float u0, r1, r2;
  if (draws) { u0 = draws[i * 3]; r1 = draws[i * 3 + 1]; r2 = draws[i * 3 + 2]; }
  else {
    const uint4 r = mesh_philox(make_uint4((uint32_t)i, (uint32_t)((uint64_t)i >> 32), 7u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    u0 = (float)(r.x >> 8) * (1.0f / 16777216.0f);
    r1 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
    r2 = (float)(r.z >> 8) * (1.0f / 16777216.0f);
  }
  int64_t lo = 0, hi = nt - 1;            // first triangle whose cumulative share exceeds u0
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cdf[mid] > u0) hi = mid; else lo = mid + 1;
  }
  const float* p = tri + lo * 9;
  const float sr = sqrtf(r1);
  const float a = 1.f - sr, b = sr * (1.f - r2), c = sr * r2;
  const float ux = p[3] - p[0], uy = p[4] - p[1], uz = p[5] - p[2];
  const float vx = p[6] - p[0], vy = p[7] - p[1], vz = p[8] - p[2];
  float nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
  const float len = fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-30f);
#pragma unroll
  for (int k = 0; k < 3; ++k) pts[i * 3 + k] = a * p[k] + b * p[3 + k] + c * p[6 + k];
  nrm[i * 3] = nx / len; nrm[i * 3 + 1] = ny / len; nrm[i * 3 + 2] = nz / len;
}

int mesh_sample_surface(const float* tri, const float* cdf, int64_t nt, int64_t n, uint64_t seed, const float* draws, float* pts, float* nrm,
                        cudaStream_t st) {
  if (n <= 0) return 0;
  DUDF_REQUIRE(nt > 0, "mesh sampler: empty triangle list");
  mesh_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tri, cdf, nt, n, seed, draws, pts, nrm);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

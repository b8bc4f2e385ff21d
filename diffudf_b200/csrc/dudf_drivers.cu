// Device-resident query drivers: the sphere-tracing loop of src/render_st.py:136-172 (propagate_rays) and the projection
// loop of src/render_pc.py:43-53 (Sampler.generate_point_cloud) around the field-query kernels.  The reference runs these
// loops on the host with numpy masks and one chunked evaluate() + device->host copy per iteration; here the ray state
// (float64 positions, active list, hit mask) stays in HBM, every iteration is one value query of the compacted active rays
// plus one fused advance / classify kernel and a stream compaction, and the host only reads back the 4-byte active count.
// All kernels here are HBM-bound, one thread per ray / point, coalesced.
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_device.cuh"

namespace dudf {

// src/inverses.py:3-22 on one value (a is |f| for sphere tracing, f for the projection, exactly as the callers pass it)
__device__ __forceinline__ float inverse_dev(int gt_mode, float a, float alpha, double sqrt_alpha, float min_step) {
  if (gt_mode == DUDF_GT_TANH) return inv_tanh_dev(a, alpha);
  if (gt_mode == DUDF_GT_SIREN) return inv_siren_dev(a, min_step);
  // 'squared': the float32 array is divided in place by the float64 scalar np.sqrt(alpha) (computed in float64, stored as float32)
  return (float)((double)((a > 0.f) ? sqrtf(a) : min_step) / (double)sqrt_alpha);
}

__global__ void __launch_bounds__(256) drv_gather_kernel(const double* __restrict__ pos, const int* __restrict__ idx, int64_t n,
                                                         float* __restrict__ x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = idx ? idx[i] : i;
  x[i * 3 + 0] = (float)pos[r * 3 + 0];
  x[i * 3 + 1] = (float)pos[r * 3 + 1];
  x[i * 3 + 2] = (float)pos[r * 3 + 2];
}

// one sphere-tracing step of the active rays (render_st.py:150-166): step = inverse(|f|), pos += dir * step in float64
// (separate multiply and add, like the array expression), hit when the step (tanh / squared) or the value (siren) is below
// the threshold while the new position is strictly inside (-1, 1)^3; a ray keeps marching while it is inside and not hit
__global__ void __launch_bounds__(256) drv_advance_kernel(double* __restrict__ pos, const double* __restrict__ dir, const int* __restrict__ idx,
                                                          const float* __restrict__ f, int64_t n, int gt_mode, float alpha, double sqrt_alpha,
                                                          float thr, unsigned char* __restrict__ hit, unsigned char* __restrict__ keep) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = idx[i];
  const float fv = f[i];
  const float step = inverse_dev(gt_mode, fabsf(fv), alpha, sqrt_alpha, 0.01f);
  const double sd = (double)step;
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double p = __dadd_rn(pos[r * 3 + k], __dmul_rn(dir[r * 3 + k], sd));
    pos[r * 3 + k] = p;
    inside = inside && (p > -1.0) && (p < 1.0);
  }
  const bool below = (gt_mode == DUDF_GT_SIREN) ? (fv < thr) : (fabsf(step) < thr);
  if (below && inside) hit[r] = 1;
  keep[i] = (!below && inside) ? 1 : 0;
}

__global__ void __launch_bounds__(256) drv_mark_kernel(const int* __restrict__ idx, int64_t n, unsigned char* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[idx[i]] = 1;
}

// one projection step (render_pc.py:46-53).  The reference works on the float64 arrays evaluate() returns (fp32 results
// widened): steps = inverse(f), x -= steps * g / |g|, all in float64
__device__ __forceinline__ double inverse_dev64(int gt_mode, double a, double alpha, double min_step) {
  if (gt_mode == DUDF_GT_TANH) return (a < 1.0 / alpha) ? sqrt(a / alpha) : a;
  if (gt_mode == DUDF_GT_SIREN) return (a > 0.0) ? a : min_step;
  return ((a > 0.0) ? sqrt(a) : min_step) / sqrt(alpha);
}
__global__ void __launch_bounds__(256) drv_project_kernel(double* __restrict__ x, const float* __restrict__ f, const float* __restrict__ g, int64_t n,
                                                          int gt_mode, double alpha, double* __restrict__ steps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double step = inverse_dev64(gt_mode, (double)f[i], alpha, 0.0);       // no abs(): a negative value gives NaN, like the reference
  const double gx = g[i * 3], gy = g[i * 3 + 1], gz = g[i * 3 + 2];
  const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz)));
  x[i * 3 + 0] = __dsub_rn(x[i * 3 + 0], __dmul_rn(step, gx / nrm));
  x[i * 3 + 1] = __dsub_rn(x[i * 3 + 1], __dmul_rn(step, gy / nrm));
  x[i * 3 + 2] = __dsub_rn(x[i * 3 + 2], __dmul_rn(step, gz / nrm));
  steps[i] = step;
}

// ---- shading of the hit points: phong_shading / ward_reflectance of src/render_st.py:174-245, one thread per hit, float64 like the
// ---- reference's numpy arrays; sums of three products in numpy's order, no fused multiply-adds
struct ShadeArgs {
  const long long* rows;      // [H] ray index of every hit, ascending (np.nonzero(hits))
  const double* samples;      // [R][3] ray positions (t0)
  const double* normals;      // [H][3]
  const double* pc1;          // [H][3] principal directions (Ward), or null
  const double* pc2;
  const double* color_map;    // [H][3] or null (grey 0.7 / 0.7 / 0.2)
  double light[3], camera[3];
  double shininess, alpha1, alpha2;
  int method;                 // 0 Blinn-Phong, 1 Ward
  double* colors;             // [R][3], rows of rays that did not hit are left as the caller filled them (ones)
};
__device__ __forceinline__ double dot3(const double* a, const double* b) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
}
__device__ __forceinline__ void normalize3(double* v) {
  const double n = sqrt(dot3(v, v));
  v[0] /= n; v[1] /= n; v[2] /= n;
}
__global__ void __launch_bounds__(256) drv_shade_kernel(ShadeArgs a, int64_t H) {
  const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  const int64_t r = a.rows[h];
  const double p[3] = {a.samples[r * 3], a.samples[r * 3 + 1], a.samples[r * 3 + 2]};
  const double n[3] = {a.normals[h * 3], a.normals[h * 3 + 1], a.normals[h * 3 + 2]};
  double Ld[3] = {a.light[0] - p[0], a.light[1] - p[1], a.light[2] - p[2]};
  normalize3(Ld);
  const double nl = dot3(n, Ld);
  const double lambertian = fmax(nl, 0.0);
  double specular = 0.0;
  if (a.method == 0) {
    const double I[3] = {-1.0 * Ld[0], -1.0 * Ld[1], -1.0 * Ld[2]};
    const double k = 2.0 * dot3(n, I);
    const double Rv[3] = {I[0] - k * n[0], I[1] - k * n[1], I[2] - k * n[2]};
    double V[3] = {p[0], p[1], p[2]};
    normalize3(V);
    const double sa = fmax(dot3(Rv, V), 0.0);
    if (a.shininess > 0.0 && lambertian > 0.0) specular = pow(sa, a.shininess);
  } else {
    double Vd[3] = {a.camera[0] - p[0], a.camera[1] - p[1], a.camera[2] - p[2]};
    normalize3(Vd);
    double Hh[3] = {Vd[0] + Ld[0], Vd[1] + Ld[1], Vd[2] + Ld[2]};
    normalize3(Hh);
    const double pc1[3] = {a.pc1[h * 3], a.pc1[h * 3 + 1], a.pc1[h * 3 + 2]};
    const double pc2[3] = {a.pc2[h * 3], a.pc2[h * 3 + 1], a.pc2[h * 3 + 2]};
    const double weight = 1.0 / (4.0 * 3.141592653589793 * a.alpha1 * a.alpha2 * sqrt(nl * dot3(n, Vd)));
    const double t1 = dot3(Hh, pc1) / a.alpha1, t2 = dot3(Hh, pc2) / a.alpha2;
    double sp = weight * exp(-2.0 * (t1 * t1 + t2 * t2) / (1.0 + dot3(n, Hh)));
    if (isnan(sp)) sp = 0.0;                                    // np.nan_to_num
    else if (isinf(sp)) sp = sp > 0.0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    specular = sp * 0.1;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double cm = a.color_map ? a.color_map[h * 3 + c] : 1.0;
    const double dc = a.color_map ? cm * 0.7 : 0.7, am = a.color_map ? cm * 0.2 : 0.2;
    double v = __dadd_rn(__dadd_rn(__dmul_rn(dc, lambertian), __dmul_rn(dc, specular)), am);
    v = fmin(fmax(v, 0.0), 0.9);
    a.colors[r * 3 + c] = v;
  }
}

static inline unsigned blocks_for(int64_t n) { return (unsigned)((n + 255) / 256); }

size_t drv_select_temp_bytes(int64_t R) {
  size_t a = 0, b = 0;
  cub::CountingInputIterator<int> it(0);
  cub::DeviceSelect::Flagged(nullptr, a, it, (const unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)R);
  cub::DeviceSelect::Flagged(nullptr, b, (const int*)nullptr, (const unsigned char*)nullptr, (int*)nullptr, (int*)nullptr, (int)R);
  return (a > b ? a : b) + 256;
}

int drv_select_initial(void* temp, size_t temp_bytes, const unsigned char* active, int64_t R, int* idx, int* d_count, cudaStream_t st) {
  cub::CountingInputIterator<int> it(0);
  DUDF_CUDA_OK(cub::DeviceSelect::Flagged(temp, temp_bytes, it, active, idx, d_count, (int)R, st));
  dudf_count_launch();
  return 0;
}
int drv_select(void* temp, size_t temp_bytes, const int* idx_in, const unsigned char* keep, int64_t n, int* idx_out, int* d_count,
               cudaStream_t st) {
  DUDF_CUDA_OK(cub::DeviceSelect::Flagged(temp, temp_bytes, idx_in, keep, idx_out, d_count, (int)n, st));
  dudf_count_launch();
  return 0;
}
int drv_gather(const double* pos, const int* idx, int64_t n, float* x, cudaStream_t st) {
  if (n <= 0) return 0;
  drv_gather_kernel<<<blocks_for(n), 256, 0, st>>>(pos, idx, n, x);
  DUDF_LAUNCH_OK();
  return 0;
}
int drv_advance(double* pos, const double* dir, const int* idx, const float* f, int64_t n, int gt_mode, float alpha, float thr,
                unsigned char* hit, unsigned char* keep, cudaStream_t st) {
  if (n <= 0) return 0;
  drv_advance_kernel<<<blocks_for(n), 256, 0, st>>>(pos, dir, idx, f, n, gt_mode, alpha, sqrt((double)alpha), thr, hit, keep);
  DUDF_LAUNCH_OK();
  return 0;
}
int drv_mark(const int* idx, int64_t n, unsigned char* mask, cudaStream_t st) {
  if (n <= 0) return 0;
  drv_mark_kernel<<<blocks_for(n), 256, 0, st>>>(idx, n, mask);
  DUDF_LAUNCH_OK();
  return 0;
}
int drv_project(double* x, const float* f, const float* g, int64_t n, int gt_mode, float alpha, double* steps, cudaStream_t st) {
  if (n <= 0) return 0;
  drv_project_kernel<<<blocks_for(n), 256, 0, st>>>(x, f, g, n, gt_mode, (double)alpha, steps);
  DUDF_LAUNCH_OK();
  return 0;
}

int drv_shade(const long long* rows, int64_t H, const double* samples, const double* normals, const double* pc1, const double* pc2,
              const double* color_map, const double* light, const double* camera, int method, double shininess, double alpha1, double alpha2,
              double* colors, cudaStream_t st) {
  if (H <= 0) return 0;
  ShadeArgs a;
  a.rows = rows; a.samples = samples; a.normals = normals; a.pc1 = pc1; a.pc2 = pc2; a.color_map = color_map;
  for (int k = 0; k < 3; ++k) { a.light[k] = light[k]; a.camera[k] = camera ? camera[k] : 0.0; }
  a.shininess = shininess; a.alpha1 = alpha1; a.alpha2 = alpha2; a.method = method; a.colors = colors;
  drv_shade_kernel<<<blocks_for(H), 256, 0, st>>>(a, H);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

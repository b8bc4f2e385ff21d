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

}  // namespace dudf

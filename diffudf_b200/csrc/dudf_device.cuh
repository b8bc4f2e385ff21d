// Device helpers shared by the CUDA-core and tensor-core kernels.
#pragma once
#include "dudf_common.cuh"
#include "dudf_kernels.h"

namespace dudf {

// grid coordinates as src/render_mc.py:36-49 computes them in fp32: idx * (2/(N-1)) + (-1)
__device__ __forceinline__ void grid_point(int64_t idx, int N, float vs, float* p) {
  int64_t i2 = idx % N, i1 = (idx / N) % N, i0 = idx / ((int64_t)N * N);
  p[0] = __fmaf_rn((float)i0, vs, 0.f) + (-1.f);
  p[1] = __fmaf_rn((float)i1, vs, 0.f) + (-1.f);
  p[2] = __fmaf_rn((float)i2, vs, 0.f) + (-1.f);
}

// finalise one point from its NCH raw channels (shared by the SIMT and tensor-core kernels)
template <int NCH>
__device__ __forceinline__ void finalize_point(const QueryOut& o, int64_t p, const float* v) {
  float f = v[0];
  if (o.flags & DUDF_Q_ABS_INV_TANH) f = inv_tanh_dev(fabsf(f), o.alpha);
  if (o.f) o.f[p] = f;
  if (o.packed) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) o.packed[p * NCH + c] = v[c];
  }
  if constexpr (NCH >= 4) {
    if (o.g) {
      float gx = v[1], gy = v[2], gz = v[3];
      if (o.flags & DUDF_Q_NEG_NORMALIZE) {         // -F.normalize(grad): src/render_mc.py:74-75
        float nrm = fmaxf(sqrtf(gx * gx + gy * gy + gz * gz), 1e-12f);
        gx = -gx / nrm; gy = -gy / nrm; gz = -gz / nrm;
      }
      o.g[p * 3 + 0] = gx; o.g[p * 3 + 1] = gy; o.g[p * 3 + 2] = gz;
    }
  }
  if constexpr (NCH >= 10) {
    if (o.H) {
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.H[p * 9 + i * 3 + j] = v[4 + sym2(i, j)];
    }
  }
  if constexpr (NCH >= 20) {
    if (o.T) {
#pragma unroll
      for (int c = 0; c < 10; ++c) o.T[p * 10 + c] = v[10 + c];
    }
  }
}


}  // namespace dudf

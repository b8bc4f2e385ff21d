// Per-row loss terms and adjoint seeds of loss_s1 (src/loss_functions.py:123-155) and loss_siren (:82-104),
// shared by the stand-alone loss kernel (dudf_misc.cu) and the fused training kernel (dudf_tc_train.cu).
// Seeds use the stored-variable convention: off-diagonal Hessian channels carry Hbar_ij + Hbar_ji.
#pragma once
#include "dudf_common.cuh"
#include "dudf_kernels.h"

namespace dudf {

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.f) - (x < 0.f)); }
__device__ __forceinline__ double sgnd(double x) { return (double)((x > 0.0) - (x < 0.0)); }

struct LossRowCfg {
  int mode;                 // DUDF_LOSS_S1 or DUDF_LOSS_SIREN
  float w[4];
  float alpha;
  float invP;               // 1 / P_global
  float up[4];              // dL/d(term) (ones when the terms are summed directly)
};

// v: the row's jet channels (f, grad[3], hess sym6), nch of them valid; nrm: ground-truth normal (read for on-surface rows);
// t: the row's unweighted-sum shares of the four terms, already multiplied by w/P;  sd: seeds d(sum_k up_k term_k)/dv.
__device__ __forceinline__ void loss_row(const LossRowCfg& a, int nch, const float* v, float d, const float* nrm, double (&t)[4], float (&sd)[10]) {
  const float f = v[0];
  const bool on = (d == 0.f);
  const float invP = a.invP;
#pragma unroll
  for (int c = 0; c < 10; ++c) sd[c] = 0.f;
  t[0] = t[1] = t[2] = t[3] = 0.0;
  if (a.mode == DUDF_LOSS_S1) {
    const float th = tanhf(a.alpha * d);
    const float tdf = d * th;
    if (on) {
      t[0] = fabsf(f);
      sd[0] = a.up[0] * a.w[0] * invP * sgnf(f);
    } else {
      t[1] = fabsf(tdf - f);
      sd[0] = -a.up[1] * a.w[1] * invP * sgnf(tdf - f);
    }
    if (a.w[3] != 0.f && nch >= 4) {
      const float gx = v[1], gy = v[2], gz = v[3];
      const float gn = sqrtf(gx * gx + gy * gy + gz * gz);
      const float tgt = fabsf(th + d * a.alpha * (1.f - th * th));
      t[3] = fabsf(gn - tgt);
      if (gn > 0.f) {
        const float k = a.up[3] * a.w[3] * invP * sgnf(gn - tgt) / gn;
        sd[1] = k * gx; sd[2] = k * gy; sd[3] = k * gz;
      }
    }
    if (a.w[2] != 0.f && nch >= 10 && on) {
      double H[3][3], lam[3], V[3][3];
      H[0][0] = v[4]; H[0][1] = H[1][0] = v[5]; H[0][2] = H[2][0] = v[6];
      H[1][1] = v[7]; H[1][2] = H[2][1] = v[8]; H[2][2] = v[9];
      eigh3<double>(H, lam, V);
      const double n0 = nrm[0], n1 = nrm[1], n2 = nrm[2];
      const double nn = fmax(sqrt(n0 * n0 + n1 * n1 + n2 * n2), 1e-8);
      const double vx = V[0][2], vy = V[1][2], vz = V[2][2];
      const double vn = fmax(sqrt(vx * vx + vy * vy + vz * vz), 1e-8);
      const double cs = (n0 * vx + n1 * vy + n2 * vz) / (nn * vn);
      t[2] = 1.0 - fabs(cs);
      const double coef = -(double)a.up[2] * a.w[2] * invP * sgnd(cs);
      const double nb[3] = {coef * (n0 / (nn * vn) - cs * vx / (vn * vn)), coef * (n1 / (nn * vn) - cs * vy / (vn * vn)),
                            coef * (n2 / (nn * vn) - cs * vz / (vn * vn))};
      double Hb[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      const double vv[3] = {vx, vy, vz};
      for (int j = 0; j < 2; ++j) {
        const double cj = (V[0][j] * nb[0] + V[1][j] * nb[1] + V[2][j] * nb[2]) / (lam[2] - lam[j]);
        for (int r = 0; r < 3; ++r)
          for (int q = 0; q < 3; ++q) Hb[r][q] += cj * 0.5 * (V[r][j] * vv[q] + vv[r] * V[q][j]);
      }
      sd[4] = (float)Hb[0][0]; sd[5] = (float)(Hb[0][1] + Hb[1][0]); sd[6] = (float)(Hb[0][2] + Hb[2][0]);
      sd[7] = (float)Hb[1][1]; sd[8] = (float)(Hb[1][2] + Hb[2][1]); sd[9] = (float)Hb[2][2];
    }
  } else {  // DUDF_LOSS_SIREN
    if (on) {
      t[0] = fabsf(f);
      sd[0] = a.up[0] * a.w[0] * invP * sgnf(f);
    } else {
      const float e = expf(-1e2f * fabsf(f));
      t[1] = e;
      sd[0] = a.up[1] * a.w[1] * invP * (-1e2f) * sgnf(f) * e;
    }
    const float gx = v[1], gy = v[2], gz = v[3];
    const float gr = sqrtf(gx * gx + gy * gy + gz * gz);
    const float gn = fmaxf(gr, 1e-8f);
    if (on) {
      const float n0 = nrm[0], n1 = nrm[1], n2 = nrm[2];
      const float nn = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-8f);
      const float cs = (gx * n0 + gy * n1 + gz * n2) / (gn * nn);
      t[2] = 1.f - cs;
      const float k = -a.up[2] * a.w[2] * invP;
      sd[1] = k * (n0 / (gn * nn) - cs * gx / (gn * gn));
      sd[2] = k * (n1 / (gn * nn) - cs * gy / (gn * gn));
      sd[3] = k * (n2 / (gn * nn) - cs * gz / (gn * gn));
    }
    t[3] = (double)(gr - 1.f) * (double)(gr - 1.f);
    if (gr > 0.f) {
      const float k = a.up[3] * a.w[3] * invP * 2.f * (gr - 1.f) / gr;
      sd[1] += k * gx; sd[2] += k * gy; sd[3] += k * gz;
    }
  }
  t[0] *= a.w[0] * (double)invP; t[1] *= a.w[1] * (double)invP;
  t[2] *= a.w[2] * (double)invP; t[3] *= a.w[3] * (double)invP;
}

}  // namespace dudf

// Internal (C++) interface between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "dudf_common.cuh"
#include "../../include/dudf_b200.h"

namespace dudf {

// where a query writes its per-point results (any pointer may be null)
struct QueryOut {
  float* f;        // [P]
  float* g;        // [P][3]
  float* H;        // [P][3][3]
  float* T;        // [P][10] symmetric third derivatives
  float* packed;   // [P][NCH] raw channels (training path)
  int flags;       // DUDF_Q_*
  float alpha;
  unsigned long long* trace;   // diagnostics (tools/trace_probe.py): event log of CTA 0, or null
};

// ---- fp32 CUDA-core path (dudf_simt.cu) ----
int simt_forward(const NetView& net, int nch, const float* x, int64_t P, int gridN, int64_t grid_first,
                 const QueryOut& out, float* Zst, float* Ast, int64_t ctot, int64_t col0, int sms, cudaStream_t st);
int simt_backward(const NetView& net, const GradView& grad, int nch, const float* x, int64_t P, const float* seeds,
                  const float* Zst, float* Zbst, int64_t ctot, int64_t col0, int sms, cudaStream_t st);
int simt_wgrad(const NetView& net, const GradView& grad, const float* Zbst, const float* Ast, int64_t ctot,
               int64_t ncols, int sms, cudaStream_t st);
int simt_tile_points(int nch);   // points per tile of the SIMT kernels
int simt_tile_cols(int nch);

// ---- tcgen05 path (dudf_tc.cu) ----
struct TcPacked;                 // opaque device image of the fp16 operand tiles
size_t tc_packed_bytes(int n_lin);
int tc_pack(const NetView& net, void* packed, cudaStream_t st);
int tc_forward(const void* packed, const NetView& net, int nch, const float* x, int64_t P, int gridN,
               int64_t grid_first, const QueryOut& out, int sms, cudaStream_t st);
// ---- split-precision tcgen05 path (dudf_tcx.cu): hi + lo fp16 operands, fp32-grade results ----
size_t tcx_packed_bytes(int n_lin);
int tcx_pack(const NetView& net, void* packed, cudaStream_t st);
int tcx_forward_dir3(const void* packed, const NetView& net, const float* x, const float* dirs, int64_t P, float* out_packed, int sms,
                     cudaStream_t st);
int tcx_forward(const void* packed, const NetView& net, int nch, const float* x, int64_t P, int gridN,
                int64_t grid_first, const QueryOut& out, int sms, cudaStream_t st);
int tc_selftest(int variant, float* max_err, cudaStream_t st);
void tc_set_trace(unsigned long long* buf);
unsigned long long* tc_get_trace();
int tc_mma_bench(int variant, int ctas, int iters, float* clk_host, cudaStream_t st);

// ---- loss epilogues, optimiser, utilities (dudf_misc.cu) ----
struct LossArgs {
  int mode;                 // DUDF_LOSS_*
  const float* packed;      // [P][NCH] forward outputs
  int nch;
  const float* normals;     // [P][3]
  const float* dist;        // [P]
  int64_t P;                // rows of this call
  int64_t P_global;         // divisor of the means
  float w[4];
  float alpha;
  const float* upstream;    // [4] dL/d(term) or null (ones)
  float* seeds;             // [P][NCH] out
  double* terms;            // [4] accumulated (atomicAdd)
  double* s2_stats;         // [3] n, sum, sumsq (mode S2)
  float* seed_absmax;       // [1] running max of |stored seed| (atomicMax), may be null
};
// ---- tcgen05 training path (dudf_tc_train.cu) ----
int tc_train_pair_cols(int nch);
int tc_train_pair_points(int nch);
struct TcSegment {          // one row segment of a training batch (consecutive stash columns, 256 per sub-tile pair)
  const float* x;           // [rows][3]
  int64_t rows;
  int nch;
  float* packed;            // forward: [rows][nch] raw channels out
  const float* seeds;       // backward: [rows][nch]
  const float* normals;     // fused step: [rows][3]
  const float* dist;        // fused step: [rows]
};
struct TcFusedLoss {        // loss configuration of the fused step (dudf_tc_train.cu)
  int mode;                 // DUDF_LOSS_S1 or DUDF_LOSS_SIREN
  float w[4];
  float alpha;
  int64_t P_global;
  double* terms;            // [4] accumulated
  const float* amax_prev;   // max|stored seed| of the previous step (loss scale)
  float* amax_next;         // this step's, raised with atomicMax
  int flags;                // DUDF_FUSED_* tuning bits
};
size_t tc_fused_scratch_bytes(const NetView& net, int sms);
int tc_train_fused(const void* packed, const NetView& net, const GradView& grad, const TcSegment* segs, int nseg, const TcFusedLoss& fl,
                   float* scratch, void* Aimg, void* Zimg, int64_t ld, int sms, cudaStream_t st);
int tc_train_forward(const void* packed, const NetView& net, const TcSegment* segs, int nseg, float* Ust, void* Aimg, int64_t ld,
                     int64_t col0, int sms, cudaStream_t st);
int tcx_train_forward(const void* packed, const NetView& net, const TcSegment* segs, int nseg, float* Ust, void* Aimg, int64_t ld,
                      int64_t col0, int sms, cudaStream_t st);
int tc_train_backward(const void* packed, const NetView& net, const GradView& grad, const TcSegment* segs, int nseg,
                      const float* seed_absmax, const float* Ust, void* Zimg, int64_t ld, int64_t col0, int sms, cudaStream_t st);
int tc_train_wgrad(const NetView& net, const GradView& grad, const void* Zimg, const void* Aimg, int64_t ld, const float* seed_absmax,
                   int sms, cudaStream_t st, int layer_lo = 0, int layer_hi = 0);
// ---- device-side batch sampler for oriented point clouds (dudf_sampler.cu; src/dataset.py:72-131) ----
struct SampleArgs {
  const float* surf_pts;    // [n_surf][3]
  const float* surf_nrm;    // [n_surf][3]
  int64_t n_surf, n_on, n_far, n_near;
  float sigma;              // std of the normal offsets of the near rows (0.01 in the reference)
  float lo[3], hi[3];       // domain of the far rows
  uint64_t seed, batch;     // Philox key / counter prefix
  const int64_t* on_idx;    // optional caller-supplied draws (parity tests): [n_on] cloud indices,
  const float* far_pts;     //   [n_far][3] domain points,
  const int64_t* near_idx;  //   [n_near] indices into the ON rows,
  const float* near_off;    //   [n_near] offsets
  float* coords;            // out [P][3]
  float* normals;           // out [P][3]
  float* dist;              // out [P]
};
int sample_batch_pc(const SampleArgs& a, int sms, cudaStream_t st);
int sample_rows(const SampleArgs& a, cudaStream_t st);       // positions / normals / |offset| only
// ---- mesh half of the sampler (dudf_mesh.cu; src/dataset.py:14-70, src/preprocess_mesh.py:29-40) ----
int mesh_distance(const float* q, int64_t nq, const float* tri, int64_t nt, float* dist, int sms, cudaStream_t st);
int mesh_sample_surface(const float* tri, const float* cdf, int64_t nt, int64_t n, uint64_t seed, const float* draws, float* pts, float* nrm,
                        cudaStream_t st);
int nn_distance(const float* q, int64_t nq, const float* X, int64_t nx, float* dist, int sms, cudaStream_t st);
// ---- spatial index of the cloud (dudf_cloud_index.cu): Morton-sorted points under a 32-ary box hierarchy, exact fp32 nearest distance ----
constexpr int CLOUD_INDEX_MAX_LEVELS = 5;
constexpr int64_t CLOUD_INDEX_MAX_POINTS = (int64_t)1 << 30;
constexpr int64_t CLOUD_INDEX_MIN_POINTS = 2048;        // below this nn_distance scans the cloud by brute force
int64_t cloud_index_bytes(int64_t n);                   // size of the caller-owned device buffer, -1 if n is out of range
int cloud_index_build(const float* X, int64_t n, void* index, cudaStream_t st);
int cloud_index_query(const float* q, int64_t nq, const void* index, int64_t n, float* dist, cudaStream_t st);
int sample_batch_pc_indexed(const SampleArgs& a, const void* index, cudaStream_t st);
// ---- device-resident query drivers (dudf_drivers.cu; src/render_st.py:136-172, src/render_pc.py:43-53) ----
size_t drv_select_temp_bytes(int64_t R);
int drv_select_initial(void* temp, size_t temp_bytes, const unsigned char* active, int64_t R, int* idx, int* d_count, cudaStream_t st);
int drv_select(void* temp, size_t temp_bytes, const int* idx_in, const unsigned char* keep, int64_t n, int* idx_out, int* d_count,
               cudaStream_t st);
int drv_gather(const double* pos, const int* idx, int64_t n, float* x, cudaStream_t st);
int drv_advance(double* pos, const double* dir, const int* idx, const float* f, int64_t n, int gt_mode, float alpha, float thr,
                unsigned char* hit, unsigned char* keep, cudaStream_t st);
int drv_mark(const int* idx, int64_t n, unsigned char* mask, cudaStream_t st);
int drv_project(double* x, const float* f, const float* g, int64_t n, int gt_mode, float alpha, double* steps, cudaStream_t st);
int drv_shade(const long long* rows, int64_t H, const double* samples, const double* normals, const double* pc1, const double* pc2,
              const double* color_map, const double* light, const double* camera, int method, double shininess, double alpha1, double alpha2,
              double* colors, cudaStream_t st);
// ---- CAP-UDF marching cubes (dudf_capmc.cu; src/render_mc.py:201-256) ----
size_t cap_scan_temp_bytes(int64_t nblocks);
int cap_units(int64_t ncell);      // blocks of consecutive cells that form the unit of the output order
int cap_classify(const float* df, const float* vecs, int N, float thr, unsigned char* code, long long* block_count, long long* block_offset,
                 void* temp, size_t temp_bytes, cudaStream_t st);
int cap_emit(const float* df, const unsigned char* code, int N, const long long* block_offset, double* tris, cudaStream_t st);
int loss_seeds(const LossArgs& a, cudaStream_t st);
int s2_finish(const double* stats, float w0, float w1, double* terms, cudaStream_t st);
int loss_s2_stats(const float* packed, const float* dist, int64_t P, double* stats, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, int64_t t,
              cudaStream_t st, const float* unsafe = nullptr, long long* skipped = nullptr);
int adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float b1, float b2, float eps, cudaStream_t st);
int adam_step_peers(float* p, const float* const* peer_grads, int world, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps,
                    int64_t t, int guarded, long long* skipped, float* g_sum_out, cudaStream_t st);
int dirs9(const float* n, const float* dirs6, int64_t P, float* d9, cudaStream_t st);
int mean_dir3(const float* jet, const float* lam, int64_t P, float* mean, cudaStream_t st);
int scale_guard(const float* amax_prev, const float* amax_next, float limit, float* flag, cudaStream_t st);
int transpose256(const float* W, float* Wt, cudaStream_t st);
int eig_normals(const float* H, const float* ref_dir, int ref_mode, int64_t P, float* n, float* dirs, float* lam,
                cudaStream_t st);
int curvature(const float* H, const float* T, int64_t P, float* n, float* mean, float* gauss, float* J, cudaStream_t st);
int field_vectors(const float* g, const float* H, int64_t P, float* vecs, cudaStream_t st);
int f32_to_f64(const float* src, double* dst, int64_t n, cudaStream_t st);

}  // namespace dudf

/* dudf_b200.h — C ABI of the B200-native DUDF hot path (libdudf_b200.so).
 *
 * The reference (LIA-DiTella/DiffUDF) has no FFI: its boundary is the Python call surface of
 * src/model.py, src/diff_operators.py, src/loss_functions.py, src/evaluate.py and the field queries
 * of src/render_mc.py / render_st.py / render_pc.py.  Each entry point below names the reference
 * interface it replaces; diffudf_b200/*.py binds them with ctypes and mirrors the reference
 * signatures (INTEGRATION.md shows the binding a reference maintainer would add).
 *
 * Conventions: every pointer is a DEVICE pointer unless its name ends in _host; tensors are dense
 * fp32 row-major; `stream` is a cudaStream_t passed as void* (NULL = default stream); the caller
 * owns all buffers; functions return 0 on success, non-zero on failure with the message available
 * from dudf_last_error().  There is no CPU fallback: calls fail when no sm_100 device is present.
 */
#ifndef DUDF_B200_H
#define DUDF_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dudf_ctx dudf_ctx;

/* arithmetic of the hidden-layer contractions */
#define DUDF_PRECISION_FP32 0  /* fp32 FFMA on CUDA cores; tolerance 1e-5 vs the fp64 oracle          */
#define DUDF_PRECISION_TC16 1  /* tcgen05 MMA, fp16 operands, fp32 TMEM accumulation; 1e-3 class (f), 2e-3 (derivatives) */
#define DUDF_PRECISION_TCX3 2  /* tcgen05 MMA, hi+lo fp16 operand split (3 MMAs per product), fp32-grade jets (1e-5);
                                  training: split forward, single-pass reverse sweep / weight gradient (gradients 1e-3) */

/* post-processing flags of the field queries */
#define DUDF_Q_ABS_INV_TANH 1  /* f <- inv_tanh(|f|, alpha)            src/inverses.py:18-19, render_mc.py:71 */
#define DUDF_Q_NEG_NORMALIZE 2 /* g <- -g / max(|g|, 1e-12)            src/render_mc.py:74-75                 */

/* loss selectors */
#define DUDF_LOSS_S1 0    /* src/loss_functions.py:123-155 */
#define DUDF_LOSS_S2 1    /* src/loss_functions.py:106-121 */
#define DUDF_LOSS_SIREN 2 /* src/loss_functions.py:82-104  */

int dudf_version(void);
const char* dudf_last_error(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t dudf_launch_count(void);

/* Network container: replaces SIREN.__init__ / load_state_dict / .to(device) (src/model.py:85-113).
 * n_hidden sine layers of width 256 (anything else is rejected), w0 = first-layer omega, ww = others. */
int dudf_create(int n_hidden, float w0, float ww, dudf_ctx** out);
int dudf_destroy(dudf_ctx* ctx);
/* W_host[i] / b_host[i] (HOST arrays of n_hidden+1 DEVICE pointers) use the nn.Linear layout of the
 * state_dict keys net.{i}.0.weight / net.{i}.0.bias.  Copies and re-packs the operand images. */
int dudf_set_weights(dudf_ctx* ctx, const float* const* W_host, const float* const* b_host, void* stream);
/* Zero-copy variant for callers that keep the parameters alive (the Python module does): dudf_bind_weights records the
 * device pointers, dudf_refresh_weights rebuilds the derived operand images after the parameters changed
 * (optimizer.step()): the fp32 transposes used by the CUDA-core forward and / or the fp16 tensor-core images. */
#define DUDF_REFRESH_FP32 1
#define DUDF_REFRESH_TC16 2
#define DUDF_REFRESH_TCX3 4
int dudf_bind_weights(dudf_ctx* ctx, const float* const* W_host, const float* const* b_host);
int dudf_refresh_weights(dudf_ctx* ctx, int what, void* stream);

/* Field query at arbitrary points: SIREN.forward + diff_operators.gradient / hessian
 * (src/model.py:116-135, src/diff_operators.py:187-212) and the third derivatives that
 * render_st.compute_curvature obtains through jacobian() (src/render_st.py:42-55).
 * order 0: f | 1: f, grad | 2: + Hessian | 3: + third derivatives (10 symmetric components, fp32 path only).
 * Any output pointer may be NULL.  x: [P][3]; f: [P]; g: [P][3]; H: [P][3][3]; T: [P][10]. */
int dudf_query_points(dudf_ctx* ctx, const float* x, int64_t P, int order, int flags, float alpha, float* f,
                      float* g, float* H, float* T, int precision, void* stream);

/* Dense grid query: the coordinate generation and evaluate() call of extract_fields
 * (src/render_mc.py:36-49,69-75).  Flat index i -> (i0,i1,i2), i2 fastest, coord = idx*2/(N-1) - 1.
 * Evaluates indices [first, first+count).  df: [count] = inv_tanh(|f|); vecs: [count][3] = -normalize(grad f)
 * (flags select the post-processing; with flags = 0 raw f and grad are written). */
int dudf_query_grid(dudf_ctx* ctx, int N, int64_t first, int64_t count, int flags, float alpha, float* df,
                    float* vecs, float* H, int precision, void* stream);

/* Top eigenvector of the Hessian ("eigen-normal") and the two principal directions:
 * torch.linalg.eigh call sites src/loss_functions.py:142-143, src/render_st.py:59-62, src/render_mc.py:77-84,
 * src/render_pc.py:65.  ref_mode 0: LAPACK-free canonical sign (largest |component| positive);
 * 1: flip so that dot(n, ref_dir) >= 0 (render_mc.py:80-84); 2: flip so that dot(n, ref_dir) <= 0
 * (render_st.py:104-105).  n: [P][3]; dirs: [P][3][2] (may be NULL); lam: [P][3] ascending (may be NULL). */
int dudf_eig_normals(const float* H, const float* ref_dir, int ref_mode, int64_t P, float* n, float* dirs,
                     float* lam, void* stream);

/* Mean and Gaussian curvature of the level set from H and the third derivatives T
 * (src/render_st.py:42-55: jacobian of the eigen-normal, src/diff_operators.py:214-227).  n: [P][3] canonical
 * sign; mean, gauss: [P]; J: [P][3][3] = d n_i / d x_k.  Any output may be NULL. */
int dudf_curvature(const float* H, const float* T, int64_t P, float* n, float* mean, float* gauss, float* J,
                   void* stream);

/* vecs <- where(|g| == 0, eigen-normal sign-aligned, g) : the fallback branch of extract_fields
 * (src/render_mc.py:77-93); g is the already normalised, negated gradient. */
int dudf_field_vectors(const float* g, const float* H, int64_t P, float* vecs, void* stream);

/* gt_mode of src/inverses.py:3-22 */
#define DUDF_GT_TANH 0     /* inv_tanh(f) = f < 1/alpha ? sqrt(f/alpha) : f */
#define DUDF_GT_SIREN 1    /* inv_siren(f) = f > 0 ? f : min_step           */
#define DUDF_GT_SQUARED 2  /* (f > 0 ? sqrt(f) : min_step) / sqrt(alpha)    */

/* Sphere tracing, the loop of propagate_rays (src/render_st.py:136-172), device resident.  Per iteration: one value query of
 * the rays still marching, step = inverse(gt_mode, |f|, alpha, min_step 0.01), pos += dir * step in float64; a ray is HIT when
 * the step (value, for DUDF_GT_SIREN) is below thr and the new position lies strictly inside (-1,1)^3, it is DROPPED when it
 * leaves the cube, otherwise it keeps marching; at most max_it iterations.
 * pos: [R][3] float64, updated in place (t0 of the reference); dir: [R][3] float64; active: [R] bytes, in: the rays to march
 * (mask_rays), out: the rays still marching after max_it; hit: [R] bytes, OR-ed with the rays that hit.  queries_host (may be
 * NULL) receives the number of value queries.  The stream is synchronised once per iteration (4-byte active count). */
int dudf_march_rays(dudf_ctx* ctx, double* pos, const double* dir, unsigned char* active, unsigned char* hit, int64_t R,
                    int gt_mode, float alpha, float thr, int max_it, int precision, int64_t* queries_host, void* stream);

/* Projection onto the zero level set, the inner loop of Sampler.generate_point_cloud (src/render_pc.py:43-53): num_steps
 * times x -= inverse(gt_mode, f, alpha, min_step 0) * grad f / |grad f| (no abs(), like the reference), in float64 on the widened
 * fp32 field values exactly as the reference computes on the float64 arrays of evaluate().  x: [P][3] float64, updated in
 * place; steps: [P] float64, the last step lengths; g: [P][3] the gradients of the last step; H: [P][3][3] the Hessians of the
 * last step, or NULL.  No host synchronisation. */
int dudf_project_points(dudf_ctx* ctx, double* x, int64_t P, int num_steps, int gt_mode, float alpha, double* steps, float* g,
                        float* H, int precision, void* stream);

/* Shading of the hit points: phong_shading (method 0) / ward_reflectance (method 1) of src/render_st.py:174-245, in float64 like
 * the reference's numpy arrays.  rows: [H] indices of the rays that hit, ascending (np.nonzero(hits)); samples: [R][3] ray
 * positions; normals: [H][3]; pc1, pc2: [H][3] principal directions (Ward only, else NULL); color_map: [H][3] or NULL (grey);
 * light_host / camera_host: 3 HOST doubles (camera: Ward only); colors: [R][3], only the rows of `rows` are written (the caller fills
 * the array with ones first, like np.ones_like(samples)). */
int dudf_shade_hits(const long long* rows, int64_t H, const double* samples, const double* normals, const double* pc1, const double* pc2,
                    const double* color_map, const double* light_host, const double* camera_host, int method, double shininess,
                    double alpha1, double alpha2, double* colors, void* stream);

/* CAP-UDF marching cubes: extract_mesh_CAP(ndf, grad, resolution) of src/render_mc.py:201-256 on the device, fed by the
 * outputs of dudf_query_grid.  df: [N][N][N] distances, vecs: [N][N][N][3] (negated, normalised) gradients.  A cell whose smallest
 * corner distance exceeds `threshold` (0.008 in the reference) is skipped; corner c is negative when
 * dot(vecs[corner 0], vecs[corner c]) < 0; cells with a negative corner are triangulated by marching cubes at iso-value 0 (vertices
 * at the linear zero crossings, float64, mapped to [-1,1]^3).  Triangles come out in the reference's cell order (i, j, k
 * lexicographic) as a soup: tris [n][3 vertices][3] float64.  The per-cell triangulation table is generated from the published
 * algorithm (tools/gen_mc_table.py); PyMCubes' table (mcubes.marching_cubes, render_mc.py:231) is not part of the reference tree.
 * Call with tris == NULL to classify the cells and get the triangle count in *n_tris_host, then with a buffer of at least that
 * many triangles (capacity, in triangles) to emit them from the classification the context holds. */
int dudf_cap_mesh(dudf_ctx* ctx, const float* df, const float* vecs, int N, float threshold, double* tris, int64_t capacity,
                  int64_t* n_tris_host, void* stream);

/* evaluate() of src/evaluate.py:5-37 with HOST buffers: chunks of at most max_batch points, fp32 compute, results
 * widened to float64 on the device and copied into the caller's arrays (any of them may be NULL).  The chunks run through a
 * two-slot pipeline (kernels on `stream`, device -> pinned staging on a copy stream, staging -> the caller's pageable arrays on
 * the host); the call returns when every result is in place. */
int dudf_evaluate_host(dudf_ctx* ctx, const float* x_host, int64_t N, int order, double* f_host, double* g_host,
                       double* H_host, int64_t max_batch, int precision, void* stream);

/* Training primitives.  loss_s1 / loss_s2 / loss_siren (src/loss_functions.py:82-155) followed by
 * train_loss.backward() (train.py:221) decompose into
 *   1. dudf_jet_forward : model(x) + gradient/hessian as forward jets, stashing what the reverse sweep needs;
 *   2. dudf_loss        : the loss terms (values) and, given dL/d(term), the per-row adjoint seeds;
 *   3. dudf_jet_backward: reverse sweep through the jet network (bias / first / last layer gradients);
 *   4. dudf_jet_wgrad   : weight gradients of the hidden 256x256 layers (contraction over all columns).
 * `packed` / `seeds` are [P][NCH] with NCH = 1, 4, 10 for order 0, 1, 2: channels f, d_x, d_y, d_z, d_xx, d_xy,
 * d_xz, d_yy, d_yz, d_zz (a seed of an off-diagonal channel is dL/dH_ij + dL/dH_ji).  The stashes Z, A, Zb are
 * caller-owned fp32 arrays [n_hidden][256][ld]; dudf_stash_columns gives the padded number of columns a call
 * uses starting at col0 (ld and col0 multiples of 4).  Gradients ACCUMULATE into gW / gb (HOST arrays of
 * n_hidden+1 DEVICE pointers, nn.Linear layouts). */
int64_t dudf_stash_columns(int order, int64_t P, int precision);
int dudf_jet_forward(dudf_ctx* ctx, const float* x, int64_t P, int order, float* packed, void* Z, void* A, int64_t ld,
                     int64_t col0, int precision, void* stream);
int dudf_jet_backward(dudf_ctx* ctx, const float* x, int64_t P, int order, const float* seeds, const float* seed_absmax,
                      const void* Z, void* Zb, int64_t ld, int64_t col0, float* const* gW_host, float* const* gb_host,
                      int precision, void* stream);
int dudf_jet_wgrad(dudf_ctx* ctx, const void* Zb, const void* A, int64_t ld, int64_t ncols, const float* seed_absmax,
                   float* const* gW_host, int precision, void* stream);
/* The same contraction for the hidden layers [layer_lo, layer_hi) only (1 <= layer_lo <= layer_hi <= n_hidden), tensor-core
 * precisions: lets a data-parallel caller launch the layers in groups and all-reduce each finished group while the next runs. */
int dudf_jet_wgrad_layers(dudf_ctx* ctx, const void* Zb, const void* A, int64_t ld, const float* seed_absmax, float* const* gW,
                          int layer_lo, int layer_hi, int precision, void* stream);
/* Several row segments of one batch in one call (one launch on the tensor-core path, which balances its persistent
 * grid over all segments): loss_s1 evaluates the Hessian jet (order 2) on the on-surface rows and the gradient jet on
 * the others.  Tensor-core path: at most 2 segments, contiguous stash columns, the order-2 segment first. */
typedef struct dudf_segment {
  const float* x;      /* [rows][3] */
  int64_t rows;
  int order;           /* 0, 1, 2 */
  float* packed;       /* forward: [rows][NCH] out */
  const float* seeds;  /* backward: [rows][NCH] in */
  int64_t col0;        /* first stash column */
} dudf_segment;
int dudf_jet_forward_multi(dudf_ctx* ctx, const dudf_segment* segs_host, int nseg, void* Z, void* A, int64_t ld, int precision,
                           void* stream);
int dudf_jet_backward_multi(dudf_ctx* ctx, const dudf_segment* segs_host, int nseg, const float* seed_absmax, const void* Z,
                            void* Zb, int64_t ld, float* const* gW_host, float* const* gb_host, int precision, void* stream);
/* With DUDF_PRECISION_TC16 the stashes change type: Z is fp32 [n_hidden][ld][256] (column-group / thread-major, private
 * to the kernel pair; ld a multiple of 64), A and Zb are fp16 operand images [n_hidden][ld / 64][256 neurons][64 columns]
 * (32 KB each, 128-byte swizzled rows: the K-major operands of the weight-gradient GEMM; zero-initialised by the
 * caller), and the reverse sweep runs
 * under a power-of-two loss scale derived from seed_absmax (1 device float, zeroed by the caller before the dudf_loss
 * calls of a step, which raise it with atomicMax; all dudf_loss calls of a step must precede its first backward). */
/* Fused training step on the tensor-core path (loss_s1 / loss_siren; train.py:204-216 = loss_fn + backward):
 * ONE launch runs forward jets, the loss terms, their adjoint seeds and the reverse sweep for every sub-tile pair
 * inside a CTA (the pre-activation stash stays in a per-CTA, L2-resident scratch), then the weight-gradient GEMM.
 * Segments as in dudf_jet_forward_multi (order-2 segment first; their columns start at column 0 of the images).
 * terms (4 device doubles) and the gradients are ACCUMULATED.  The reverse sweep's power-of-two loss scale is taken
 * from amax_prev = max|stored seed| of the PREVIOUS step with the same loss configuration (1 device float, e.g. the
 * seed_absmax a dudf_loss pass or an earlier fused step produced); amax_next (zeroed by the caller) receives this
 * step's.  scratch: dudf_fused_scratch_bytes(ctx) bytes.  A, Zb: fp16 operand images as for dudf_jet_wgrad, ld columns. */
typedef struct dudf_train_segment {
  const float* x;        /* [rows][3] */
  const float* normals;  /* [rows][3] */
  const float* dist;     /* [rows] */
  int64_t rows;
  int order;             /* 1 or 2 */
  float* packed;         /* optional out: [rows][NCH] jets (may be NULL) */
} dudf_train_segment;
#define DUDF_FUSED_DISCARD 1      /* drop consumed scratch lines from L2 instead of letting them be written back */
#define DUDF_FUSED_IMG_EVICT_FIRST 2 /* operand-image stores carry an L2 evict-first hint */
#define DUDF_FUSED_NO_WGRAD 4     /* stop after the fused launch; the caller runs dudf_jet_wgrad(…, amax_prev, …) itself */
int64_t dudf_fused_scratch_bytes(const dudf_ctx* ctx);
int dudf_train_step_fused(dudf_ctx* ctx, int mode, const dudf_train_segment* segs_host, int nseg, int64_t P_global,
                          const float* w_host, float alpha, double* terms, const float* amax_prev, float* amax_next, void* scratch,
                          void* A, void* Zb, int64_t ld, float* const* gW_host, float* const* gb_host, int flags, void* stream);
/* Loss epilogue over P rows.  w_host: 4 host floats (loss weights); P_global: divisor of the means (sum of rows over
 * data-parallel ranks).  terms (4 device doubles, may be NULL) is ACCUMULATED with this call's share of each term in
 * the order of the reference dicts: S1 {sdf_on_surf, sdf_off_surf, hessian_constraint, grad_constraint}, SIREN
 * {sdf_on_surf, sdf_off_surf, normal_constraint, grad_constraint}.  seeds (may be NULL) receives
 * d(sum_k upstream[k] term_k)/d(packed); upstream: 4 device floats or NULL (ones).  For S2 the seeds need the
 * (all-reduced) statistics n, sum, sum of squares of the on-surface predictions: dudf_loss_s2_stats accumulates them
 * into 3 device doubles, dudf_loss_s2_finish turns them into the two S2 terms. */
int dudf_loss(int mode, const float* packed, int nch, const float* normals, const float* dist, int64_t P, int64_t P_global,
              const float* w_host, float alpha, const float* upstream, float* seeds, float* seed_absmax, double* terms,
              double* s2_stats, void* stream);
int dudf_loss_s2_stats(const float* packed, const float* dist, int64_t P, double* stats, void* stream);
int dudf_loss_s2_finish(const double* stats, float w0, float w1, double* terms, void* stream);
/* ---- Batch sampler for oriented point clouds (src/dataset.py:72-131, PointCloud(onlyPCloud=True)) ----
 * dudf_sample_batch_pc replaces sampleTrainingDataPC: writes one [n_on | n_far | n_near] batch (coords [P][3], normals
 * [P][3], dist [P], fp32, P = n_on + n_far + n_near; the reference's samplesFar = samplesOffSurface // 2 split is the
 * caller's) entirely on the device: cloud rows, uniform points of [lo, hi] with their nearest-cloud-point distance,
 * cloud rows displaced along the normal by N(0, sigma) with distance |offset|.  Draws come from Philox4x32-10 keyed by
 * (seed, batch_index) unless the caller supplies them (any of on_idx / far_pts / near_idx / near_off may be NULL).
 * dudf_nearest_distance replaces shortestDistance (:72-78): dist[i] = min_j |q_i - X_j| without materialising the
 * n_q x n_x matrix.  All pointers device. */
int dudf_sample_batch_pc(const float* surf_pts, const float* surf_normals, int64_t n_surf, int64_t n_on, int64_t n_far, int64_t n_near,
                         float sigma, const float* lo_host, const float* hi_host, uint64_t seed, uint64_t batch_index,
                         const int64_t* on_idx, const float* far_pts, const int64_t* near_idx, const float* near_off, float* coords,
                         float* normals, float* dist, void* stream);
int dudf_nearest_distance(const float* queries, int64_t n_q, const float* cloud, int64_t n_x, float* dist, void* stream);
/* Spatial index of the cloud for the same distances (the reference calls shortestDistance on the SAME cloud for every batch,
 * src/dataset.py:116-118, 176-185): the points sorted along a Morton curve under a 32-ary hierarchy of boxes, built once into a
 * caller-owned, 16-byte aligned device buffer of dudf_cloud_index_bytes(n_x) bytes (-1: n_x outside 1 .. 2^30).  Queries walk it
 * best-first, one warp each; distances are EXACT in fp32 (difference form, monotone lower bounds).  dudf_nearest_distance builds
 * a temporary index by itself for clouds of 2 048 points and more; dudf_sample_batch_pc_indexed is dudf_sample_batch_pc with the
 * far rows measured through a prebuilt index (what PointCloud does for every batch). */
int64_t dudf_cloud_index_bytes(int64_t n_x);
int dudf_cloud_index_build(const float* cloud, int64_t n_x, void* index, void* stream);
int dudf_nearest_distance_indexed(const float* queries, int64_t n_q, const void* index, int64_t n_x, float* dist, void* stream);
int dudf_sample_batch_pc_indexed(const float* surf_pts, const float* surf_normals, int64_t n_surf, const void* index, int64_t n_on,
                                 int64_t n_far, int64_t n_near, float sigma, const float* lo_host, const float* hi_host, uint64_t seed,
                                 uint64_t batch_index, const int64_t* on_idx, const float* far_pts, const int64_t* near_idx,
                                 const float* near_off, float* coords, float* normals, float* dist, void* stream);
/* ---- Mesh half of the batch sampler (src/dataset.py:14-70, PointCloud(onlyPCloud=False); src/preprocess_mesh.py:29-40) ----
 * dudf_mesh_distance replaces scene.compute_signed_distance (Open3D RaycastingScene, :35,50) by the UNSIGNED distance
 * dist[i] = min_t |q_i - triangle_t| (brute force; the losses are even in the distance and the ray-parity sign is undefined
 * for open surfaces).  triangles: [n_tri][3 vertices][3] fp32, device.
 * dudf_sample_batch_mesh writes one [n_on | n_far | n_near] batch like dudf_sample_batch_pc, with the distance of ALL
 * off-surface rows (far and near) measured to the mesh, as sampleTrainingData does.
 * dudf_mesh_sample_surface replaces mesh.sample_points_uniformly(n, use_triangle_normal=True): triangle by inverse CDF of the
 * areas (cdf: [n_tri] inclusive prefix sums / total, device), barycentric (1 - sqrt r1, sqrt r1 (1 - r2), sqrt r1 r2), triangle
 * normal.  draws ([n][3] = u_triangle, r1, r2 in [0,1)) may be NULL (Philox keyed by seed). */
int dudf_mesh_distance(const float* queries, int64_t n_q, const float* triangles, int64_t n_tri, float* dist, void* stream);
int dudf_sample_batch_mesh(const float* surf_pts, const float* surf_normals, int64_t n_surf, const float* triangles, int64_t n_tri,
                           int64_t n_on, int64_t n_far, int64_t n_near, float sigma, const float* lo_host, const float* hi_host,
                           uint64_t seed, uint64_t batch_index, const int64_t* on_idx, const float* far_pts, const int64_t* near_idx,
                           const float* near_off, float* coords, float* normals, float* dist, void* stream);
int dudf_mesh_sample_surface(const float* triangles, const float* cdf, int64_t n_tri, int64_t n, uint64_t seed, const float* draws,
                             float* points, float* normals, void* stream);
/* torch.optim.Adam.step as configured in train.py:334-337 (betas, eps given explicitly, no weight decay);
 * t is the 1-based step count.  Flat fp32 arrays of n elements. */
int dudf_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                   float eps, int64_t t, void* stream);
/* The same update for a step that is replayed as a CUDA graph (kernel arguments are frozen at capture, so the scalars that change
 * from step to step live on the device): state = 6 device floats, 8-byte aligned: [0] learning rate (fp32, written by the caller when
 * the schedule changes it), [2..3] the 64-bit count of steps taken so far (the kernel uses t = count + 1 and advances it), [4] an
 * internal block ticket (zero-initialised).  Bias corrections are formed in double as dudf_adam_step forms them on the host. */
int dudf_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2, float eps,
                       void* stream);
/* Safety net of the fused tensor-core step (dudf_train_step_fused scales its fp16 adjoints with the PREVIOUS step's seed
 * magnitude): dudf_scale_guard adds 1 to *flag when S(amax_prev) * amax_next exceeds `limit` (nominal range (1024, 2048], fp16
 * saturates at 65504) or is not finite; the flag lives in the slot behind the flat gradient so that the data-parallel
 * all-reduce sums it over the ranks.  dudf_adam_step_guarded is dudf_adam_step that leaves p, m, v untouched and increments
 * *skipped (device int64, may be NULL) when *unsafe_flag != 0 — the GradScaler rule of the reference ecosystem's mixed-precision
 * training (the reference itself trains in fp32, train.py:204-222, and has no such window). */
int dudf_scale_guard(const float* amax_prev, const float* amax_next, float limit, float* flag, void* stream);
int dudf_adam_step_guarded(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                           float eps, int64_t t, const float* unsafe_flag, int64_t* skipped, void* stream);

/* Eigen-normals, principal directions and MEAN curvature of P points on tensor cores (split-precision mode): replaces
 * compute_normals_and_cd + compute_curvature(..., 'mean') of src/render_st.py:42-62 (autograd Hessian, torch.linalg.eigh, autograd
 * Jacobian of the eigen-normal).  Hessian jet -> 3x3 eigen-solve -> a 10-channel DIRECTIONAL third-order jet along (v_0, v_1, n) per
 * point (SURVEY 8 a-M) -> tr(dn/dx) / 2 = (T(v_0,v_0,n) / (lam_2 - lam_0) + T(v_1,v_1,n) / (lam_2 - lam_1)) / 2.
 * normals [P][3], dirs [P][3][2] (may be NULL), mean [P]: fp32 device arrays, eigenvector signs canonical as dudf_eig_normals. */
int dudf_mean_curvature(dudf_ctx* ctx, const float* x, int64_t P, float* normals, float* dirs, float* mean, void* stream);

/* Data-parallel optimiser step with the gradient all-reduce fused in (the reference has no distributed code — this is
 * the one exchange step of the sharded training path): peer_grads = HOST array of `world` DEVICE pointers, rank r's flat gradient
 * buffer of n + 1 floats (element n = the dudf_scale_guard flag) mapped into this process (CUDA IPC / symmetric memory over
 * NVLink).  The caller orders the launch behind a cross-rank barrier (every rank's gradient complete) and alternates two buffer
 * sets per step.  Each rank sums the peers in rank order (bit-identical on all ranks) and applies dudf_adam_step's update;
 * guarded != 0: skip + count when the summed flag is non-zero; g_sum_out (optional, n floats): the summed gradient. */
int dudf_adam_step_peers(float* p, const float* const* peer_grads, int world, float* m, float* v, int64_t n, float lr, float beta1,
                         float beta2, float eps, int64_t t, int guarded, int64_t* skipped, float* g_sum_out, void* stream);

/* MeshUDF marching cubes — HOST code, plain C++17, no device involved.  Replaces
 * _marching_cubes_lewiner_cy.marching_cubes_udf(im, grads, luts, st, classic, avg_thresh, max_thresh, mask)
 * (src/marching_cubes/_marching_cubes_lewiner_cy.pyx:1116-1774) as called by udf_mc_lewiner (_marching_cubes_lewiner.py:80-141)
 * from extract_mesh_MESHUDF (src/render_mc.py:127-133): im [nz][ny][nx] fp32 distances, grads [nz][ny][nx][3] fp32, mask
 * [nz][ny][nx] bytes or NULL; luts_blob = Lewiner's look-up tables as written by tools/export_lewiner_luts.py
 * (diffudf_b200/data/lewiner_luts.bin).  The result arrays are malloc'ed by the library (vertices in (x, y, z) grid units,
 * normalised normals, values, faces in emission order — what Cell.get_vertices / get_normals / get_values / get_faces return) and
 * released with dudf_meshudf_free.  Same serial visit order as the reference: identical arrays on identical fields. */
typedef struct dudf_meshudf_result {
  float* vertices;
  float* normals;
  float* values;
  int32_t* faces;
  int64_t n_vertices, n_faces;
} dudf_meshudf_result;
int dudf_meshudf_mc(const float* im, const float* grads, int nz, int ny, int nx, int step, float avg_thresh, float max_thresh,
                    const unsigned char* mask, const void* luts_blob, int64_t luts_bytes, dudf_meshudf_result* out);
void dudf_meshudf_free(dudf_meshudf_result* r);

/* bring-up / regression tests of the tcgen05 building blocks (tests/test_gpu_umma.py) */
int dudf_selftest_umma(int variant, float* max_err_host);
/* Tensor-pipe micro-benchmark (tools/umma_bench.py): average clocks per tcgen05.mma (K = 16) over `ctas` CTAs that each
 * issue iters x 16 instructions on resident operands; variants in dudf_tc.cu. */
int dudf_bench_umma(int variant, int ctas, int iters, float* clk_per_mma_host);
/* Diagnostics (tools/trace_probe.py): CTA 0 of the tensor-core query kernel logs (tag, clock) events of its MMA warp and of
 * epilogue warp 0 into device_buffer (2 x 8192 uint64); NULL switches the log off. */
int dudf_debug_set_trace(void* device_buffer);

#ifdef __cplusplus
}
#endif
#endif

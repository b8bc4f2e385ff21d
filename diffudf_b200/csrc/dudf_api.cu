// C-ABI layer of libdudf_b200.so (declared in include/dudf_b200.h): context, workspaces and the
// orchestration of the kernels.  No torch types; pointers + sizes only.
#include <cstdarg>
#include <cstring>
#include <string>
#include <vector>
#include "dudf_common.cuh"
#include "dudf_kernels.h"

static thread_local std::string g_err;
void dudf_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

static long long g_launches = 0;
void dudf_count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }

namespace dudf {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      dudf_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
      return 1;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace dudf

struct dudf_ctx {
  int n_hidden = 0, n_lin = 0, device = 0, sms = 148;
  float w0 = 30.f, ww = 30.f;
  bool weights_set = false;
  float* Wd[DUDF_MAX_LAYERS] = {};
  float* bd[DUDF_MAX_LAYERS] = {};
  float* Wtd[DUDF_MAX_LAYERS] = {};
  const float* Wp[DUDF_MAX_LAYERS] = {};   // weights in use: the owned copies (dudf_set_weights) or borrowed (dudf_bind_weights)
  const float* bp[DUDF_MAX_LAYERS] = {};
  void* tc_packed = nullptr;
  void* tcx_packed = nullptr;
  dudf::DevBuf ws_out, ws_x64, ws_drv, ws_cap;
  // dudf_evaluate_host: pinned result staging (two slots), the stream of the device -> host copies and their events
  void* ev_pinned = nullptr;
  size_t ev_pinned_cap = 0;
  cudaStream_t ev_copy_stream = nullptr;
  cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int cap_N = 0;                 // grid size of the classification held in ws_cap (0: none)
  const float* cap_df = nullptr;
  int64_t cap_ntris = 0;
  NetView view() const {
    NetView v;
    memset(&v, 0, sizeof(v));
    for (int i = 0; i < n_lin; ++i) { v.W[i] = Wp[i]; v.b[i] = bp[i]; v.Wt[i] = Wtd[i]; }
    v.n_lin = n_lin;
    v.w0 = w0;
    v.ww = ww;
    return v;
  }
};

using namespace dudf;

extern "C" {

int dudf_version(void) { return 100; }
const char* dudf_last_error(void) { return g_err.c_str(); }
int64_t dudf_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int dudf_create(int n_hidden, float w0, float ww, dudf_ctx** out) {
  DUDF_REQUIRE(out != nullptr, "dudf_create: null output pointer");
  DUDF_REQUIRE(n_hidden >= 1 && n_hidden + 1 <= DUDF_MAX_LAYERS, "dudf_create: n_hidden=%d unsupported (1..%d)", n_hidden,
               DUDF_MAX_LAYERS - 1);
  int dev = 0;
  DUDF_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DUDF_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  DUDF_REQUIRE(prop.major == 10, "dudf_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", dev,
               prop.major, prop.minor);
  dudf_ctx* c = new dudf_ctx();
  c->n_hidden = n_hidden;
  c->n_lin = n_hidden + 1;
  c->device = dev;
  c->sms = prop.multiProcessorCount;
  c->w0 = w0;
  c->ww = ww;
  for (int i = 0; i < c->n_lin; ++i) {
    const size_t in = (i == 0) ? 3 : 256, outn = (i == c->n_lin - 1) ? 1 : 256;
    DUDF_CUDA_OK(cudaMalloc(&c->Wd[i], in * outn * sizeof(float)));
    DUDF_CUDA_OK(cudaMalloc(&c->bd[i], outn * sizeof(float)));
    if (i > 0 && i < c->n_lin - 1) DUDF_CUDA_OK(cudaMalloc(&c->Wtd[i], 256 * 256 * sizeof(float)));
  }
  DUDF_CUDA_OK(cudaMalloc(&c->tc_packed, tc_packed_bytes(c->n_lin)));
  DUDF_CUDA_OK(cudaMalloc(&c->tcx_packed, tcx_packed_bytes(c->n_lin)));
  // stream-ordered scratch (cudaMallocAsync in the samplers, the mesh distance, dudf_mean_curvature, the ray / projection drivers):
  // keep freed blocks in the device's default pool instead of returning them to the driver at every synchronisation (the default
  // release threshold is 0: a 25 MB workspace re-mapped per call cost 10 ... 850 ms on a multi-GPU box)
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  *out = c;
  return 0;
}

int dudf_destroy(dudf_ctx* c) {
  if (!c) return 0;
  for (int i = 0; i < c->n_lin; ++i) {
    cudaFree(c->Wd[i]);
    cudaFree(c->bd[i]);
    if (c->Wtd[i]) cudaFree(c->Wtd[i]);
  }
  cudaFree(c->tc_packed);
  cudaFree(c->tcx_packed);
  c->ws_out.release();
  c->ws_x64.release();
  if (c->ev_pinned) cudaFreeHost(c->ev_pinned);
  if (c->ev_copy_stream) cudaStreamDestroy(c->ev_copy_stream);
  for (int k = 0; k < 2; ++k) {
    if (c->ev_ready[k]) cudaEventDestroy(c->ev_ready[k]);
    if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
  }
  delete c;
  return 0;
}

int dudf_refresh_weights(dudf_ctx* c, int what, void* stream) {
  DUDF_REQUIRE(c && c->Wp[0], "dudf_refresh_weights: no weights bound");
  cudaStream_t st = (cudaStream_t)stream;
  if (what & DUDF_REFRESH_FP32) {
    for (int i = 0; i < c->n_lin; ++i)
      if (c->Wtd[i]) {
        int rc = transpose256(c->Wp[i], c->Wtd[i], st);
        if (rc) return rc;
      }
  }
  if (what & DUDF_REFRESH_TC16) {
    int rc = tc_pack(c->view(), c->tc_packed, st);
    if (rc) return rc;
  }
  if (what & DUDF_REFRESH_TCX3) {
    int rc = tcx_pack(c->view(), c->tcx_packed, st);
    if (rc) return rc;
  }
  c->weights_set = true;
  return 0;
}

int dudf_bind_weights(dudf_ctx* c, const float* const* W, const float* const* b) {
  DUDF_REQUIRE(c && W && b, "dudf_bind_weights: null argument");
  for (int i = 0; i < c->n_lin; ++i) {
    DUDF_REQUIRE(W[i] && b[i], "dudf_bind_weights: null pointer for layer %d", i);
    c->Wp[i] = W[i];
    c->bp[i] = b[i];
  }
  return 0;
}

int dudf_set_weights(dudf_ctx* c, const float* const* W, const float* const* b, void* stream) {
  DUDF_REQUIRE(c && W && b, "dudf_set_weights: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < c->n_lin; ++i) {
    const size_t in = (i == 0) ? 3 : 256, outn = (i == c->n_lin - 1) ? 1 : 256;
    DUDF_REQUIRE(W[i] && b[i], "dudf_set_weights: null pointer for layer %d", i);
    DUDF_CUDA_OK(cudaMemcpyAsync(c->Wd[i], W[i], in * outn * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DUDF_CUDA_OK(cudaMemcpyAsync(c->bd[i], b[i], outn * sizeof(float), cudaMemcpyDeviceToDevice, st));
    c->Wp[i] = c->Wd[i];
    c->bp[i] = c->bd[i];
  }
  return dudf_refresh_weights(c, DUDF_REFRESH_FP32 | DUDF_REFRESH_TC16 | DUDF_REFRESH_TCX3, stream);
}

static int order_to_nch(int order) { return order == 0 ? 1 : order == 1 ? 4 : order == 2 ? 10 : order == 3 ? 20 : -1; }

static int run_forward(dudf_ctx* c, int nch, const float* x, int64_t P, int gridN, int64_t first, const QueryOut& out,
                       int precision, cudaStream_t st) {
  if (precision == DUDF_PRECISION_TC16) {
    DUDF_REQUIRE(nch != 20, "third-order jets are only available with DUDF_PRECISION_FP32");
    return tc_forward(c->tc_packed, c->view(), nch, x, P, gridN, first, out, c->sms, st);
  }
  if (precision == DUDF_PRECISION_TCX3) {
    DUDF_REQUIRE(nch != 20, "third-order jets are only available with DUDF_PRECISION_FP32");
    return tcx_forward(c->tcx_packed, c->view(), nch, x, P, gridN, first, out, c->sms, st);
  }
  DUDF_REQUIRE(precision == DUDF_PRECISION_FP32, "unknown precision %d", precision);
  return simt_forward(c->view(), nch, x, P, gridN, first, out, nullptr, nullptr, 0, 0, c->sms, st);
}

int dudf_query_points(dudf_ctx* c, const float* x, int64_t P, int order, int flags, float alpha, float* f, float* g,
                      float* H, float* T, int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_query_points: weights not set");
  DUDF_REQUIRE(x != nullptr || P == 0, "dudf_query_points: null coordinates");
  const int nch = order_to_nch(order);
  DUDF_REQUIRE(nch > 0, "dudf_query_points: order %d unsupported (0..3)", order);
  if (P <= 0) return 0;
  QueryOut o{f, g, H, T, nullptr, flags, alpha};
  return run_forward(c, nch, x, P, 0, 0, o, precision, (cudaStream_t)stream);
}

int dudf_query_grid(dudf_ctx* c, int N, int64_t first, int64_t count, int flags, float alpha, float* df, float* vecs,
                    float* H, int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_query_grid: weights not set");
  DUDF_REQUIRE(N >= 2, "dudf_query_grid: N=%d", N);
  DUDF_REQUIRE(first >= 0 && count >= 0 && first + count <= (int64_t)N * N * N, "dudf_query_grid: range out of bounds");
  if (count == 0) return 0;
  const int nch = H ? 10 : (vecs ? 4 : 1);
  QueryOut o{df, vecs, H, nullptr, nullptr, flags, alpha};
  return run_forward(c, nch, nullptr, count, N, first, o, precision, (cudaStream_t)stream);
}

int dudf_mean_curvature(dudf_ctx* c, const float* x, int64_t P, float* normals, float* dirs, float* mean, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_mean_curvature: weights not set");
  DUDF_REQUIRE((x && normals && mean) || P == 0, "dudf_mean_curvature: null argument");
  if (P <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // workspace: H [P][9], lam [P][3], dirs6 [P][6], dirs9 [P][9], jet [P][10]
  float* ws = nullptr;
  DUDF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), (size_t)P * 37 * sizeof(float), st));
  float *H = ws, *lam = ws + P * 9, *d6 = ws + P * 12, *d9 = ws + P * 18, *jet = ws + P * 27;
  QueryOut o{nullptr, nullptr, H, nullptr, nullptr, 0, 0.f};
  int rc = tcx_forward(c->tcx_packed, c->view(), 10, x, P, 0, 0, o, c->sms, st);
  if (!rc) rc = eig_normals(H, nullptr, 0, P, normals, dirs ? dirs : d6, lam, st);
  if (!rc) rc = dirs9(normals, dirs ? dirs : d6, P, d9, st);
  if (!rc) rc = tcx_forward_dir3(c->tcx_packed, c->view(), x, d9, P, jet, c->sms, st);
  if (!rc) rc = mean_dir3(jet, lam, P, mean, st);
  cudaFreeAsync(ws, st);
  return rc;
}

int dudf_eig_normals(const float* H, const float* ref_dir, int ref_mode, int64_t P, float* n, float* dirs, float* lam,
                     void* stream) {
  DUDF_REQUIRE(H && n, "dudf_eig_normals: null argument");
  return eig_normals(H, ref_dir, ref_mode, P, n, dirs, lam, (cudaStream_t)stream);
}

int dudf_curvature(const float* H, const float* T, int64_t P, float* n, float* mean, float* gauss, float* J, void* stream) {
  DUDF_REQUIRE(H && T, "dudf_curvature: null argument");
  return curvature(H, T, P, n, mean, gauss, J, (cudaStream_t)stream);
}

int dudf_field_vectors(const float* g, const float* H, int64_t P, float* vecs, void* stream) {
  DUDF_REQUIRE(g && H && vecs, "dudf_field_vectors: null argument");
  return field_vectors(g, H, P, vecs, (cudaStream_t)stream);
}

int dudf_march_rays(dudf_ctx* c, double* pos, const double* dir, unsigned char* active, unsigned char* hit, int64_t R, int gt_mode,
                    float alpha, float thr, int max_it, int precision, int64_t* queries_host, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_march_rays: weights not set");
  DUDF_REQUIRE(R >= 0 && R < (int64_t)1 << 31, "dudf_march_rays: R=%lld out of range", (long long)R);
  DUDF_REQUIRE(gt_mode >= DUDF_GT_TANH && gt_mode <= DUDF_GT_SQUARED, "dudf_march_rays: gt_mode %d", gt_mode);
  if (queries_host) *queries_host = 0;
  if (R == 0) return 0;
  DUDF_REQUIRE(pos && dir && active && hit, "dudf_march_rays: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  // workspace: idx[2][R] int | count int (+pad) | keep[R] bytes | x[R][3] float | f[R] float | cub temp
  const size_t temp_bytes = drv_select_temp_bytes(R);
  const size_t off_cnt = 2 * (size_t)R * sizeof(int);
  const size_t off_keep = off_cnt + 256;
  const size_t off_x = (off_keep + (size_t)R + 255) / 256 * 256;
  const size_t off_f = off_x + (size_t)R * 3 * sizeof(float);
  const size_t off_tmp = (off_f + (size_t)R * sizeof(float) + 255) / 256 * 256;
  if (c->ws_drv.ensure(off_tmp + temp_bytes)) return 1;
  unsigned char* ws = (unsigned char*)c->ws_drv.p;
  int* idx[2] = {(int*)ws, (int*)ws + R};
  int* d_count = (int*)(ws + off_cnt);
  unsigned char* keep = ws + off_keep;
  float* x = (float*)(ws + off_x);
  float* f = (float*)(ws + off_f);
  void* temp = ws + off_tmp;
  int rc = drv_select_initial(temp, temp_bytes, active, R, idx[0], d_count, st);
  if (rc) return rc;
  int n = 0;
  DUDF_CUDA_OK(cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  DUDF_CUDA_OK(cudaStreamSynchronize(st));
  int cur = 0;
  int64_t nq = 0;
  for (int it = 0; it < max_it && n > 0; ++it) {
    if ((rc = drv_gather(pos, idx[cur], n, x, st))) return rc;
    QueryOut o{f, nullptr, nullptr, nullptr, nullptr, 0, 0.f};
    if ((rc = run_forward(c, 1, x, n, 0, 0, o, precision, st))) return rc;
    nq += n;
    if ((rc = drv_advance(pos, dir, idx[cur], f, n, gt_mode, alpha, thr, hit, keep, st))) return rc;
    if ((rc = drv_select(temp, temp_bytes, idx[cur], keep, n, idx[cur ^ 1], d_count, st))) return rc;
    cur ^= 1;
    DUDF_CUDA_OK(cudaMemcpyAsync(&n, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    DUDF_CUDA_OK(cudaStreamSynchronize(st));
  }
  DUDF_CUDA_OK(cudaMemsetAsync(active, 0, (size_t)R, st));
  if ((rc = drv_mark(idx[cur], n, active, st))) return rc;
  if (queries_host) *queries_host = nq;
  return 0;
}

int dudf_project_points(dudf_ctx* c, double* x, int64_t P, int num_steps, int gt_mode, float alpha, double* steps, float* g, float* H,
                        int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_project_points: weights not set");
  DUDF_REQUIRE(gt_mode >= DUDF_GT_TANH && gt_mode <= DUDF_GT_SQUARED, "dudf_project_points: gt_mode %d", gt_mode);
  DUDF_REQUIRE(P >= 0 && num_steps >= 0, "dudf_project_points: negative size");
  if (P == 0 || num_steps == 0) return 0;
  DUDF_REQUIRE(x && steps && g, "dudf_project_points: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (c->ws_drv.ensure((size_t)P * 4 * sizeof(float))) return 1;
  float* x32 = (float*)c->ws_drv.p;
  float* f = x32 + 3 * P;
  for (int s = 0; s < num_steps; ++s) {
    int rc = drv_gather(x, nullptr, P, x32, st);
    if (rc) return rc;
    const bool last = (s == num_steps - 1);
    QueryOut o{f, g, (last ? H : nullptr), nullptr, nullptr, 0, 0.f};
    if ((rc = run_forward(c, (last && H) ? 10 : 4, x32, P, 0, 0, o, precision, st))) return rc;
    if ((rc = drv_project(x, f, g, P, gt_mode, alpha, steps, st))) return rc;
  }
  return 0;
}

int dudf_shade_hits(const long long* rows, int64_t H, const double* samples, const double* normals, const double* pc1, const double* pc2,
                    const double* color_map, const double* light_host, const double* camera_host, int method, double shininess,
                    double alpha1, double alpha2, double* colors, void* stream) {
  DUDF_REQUIRE(H >= 0, "dudf_shade_hits: negative hit count");
  if (H == 0) return 0;
  DUDF_REQUIRE(rows && samples && normals && light_host && colors, "dudf_shade_hits: null argument");
  DUDF_REQUIRE(method == 0 || method == 1, "dudf_shade_hits: method %d (0 Blinn-Phong, 1 Ward)", method);
  DUDF_REQUIRE(method == 0 || (pc1 && pc2 && camera_host), "dudf_shade_hits: Ward reflectance needs the principal directions and the camera");
  return drv_shade(rows, H, samples, normals, pc1, pc2, color_map, light_host, camera_host, method, shininess, alpha1, alpha2, colors,
                   (cudaStream_t)stream);
}

int dudf_cap_mesh(dudf_ctx* c, const float* df, const float* vecs, int N, float threshold, double* tris, int64_t capacity,
                  int64_t* n_tris_host, void* stream) {
  DUDF_REQUIRE(c != nullptr, "dudf_cap_mesh: null context");
  DUDF_REQUIRE(df && vecs && n_tris_host, "dudf_cap_mesh: null argument");
  DUDF_REQUIRE(N >= 2 && N <= 1290, "dudf_cap_mesh: N=%d out of range (2..1290)", N);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t M = N - 1, ncell = M * M * M, nblocks = cap_units(ncell);
  const size_t off_cnt = ((size_t)ncell + 255) / 256 * 256;
  const size_t off_off = off_cnt + (size_t)nblocks * sizeof(long long);
  const size_t off_tmp = off_off + (size_t)nblocks * sizeof(long long);
  const size_t temp_bytes = cap_scan_temp_bytes(nblocks);
  if (tris == nullptr) {
    c->cap_N = 0;
    if (c->ws_cap.ensure(off_tmp + temp_bytes)) return 1;
    unsigned char* ws = (unsigned char*)c->ws_cap.p;
    int rc = cap_classify(df, vecs, N, threshold, ws, (long long*)(ws + off_cnt), (long long*)(ws + off_off), ws + off_tmp, temp_bytes, st);
    if (rc) return rc;
    long long last_cnt = 0, last_off = 0;
    DUDF_CUDA_OK(cudaMemcpyAsync(&last_cnt, (long long*)(ws + off_cnt) + (nblocks - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
    DUDF_CUDA_OK(cudaMemcpyAsync(&last_off, (long long*)(ws + off_off) + (nblocks - 1), sizeof(long long), cudaMemcpyDeviceToHost, st));
    DUDF_CUDA_OK(cudaStreamSynchronize(st));
    c->cap_N = N;
    c->cap_df = df;
    c->cap_ntris = last_cnt + last_off;
    *n_tris_host = c->cap_ntris;
    return 0;
  }
  DUDF_REQUIRE(c->cap_N == N && c->cap_df == df, "dudf_cap_mesh: emit without a matching classification (call with tris == NULL first)");
  DUDF_REQUIRE(capacity >= c->cap_ntris, "dudf_cap_mesh: buffer holds %lld triangles, %lld needed", (long long)capacity, (long long)c->cap_ntris);
  *n_tris_host = c->cap_ntris;
  if (c->cap_ntris == 0) return 0;
  unsigned char* ws = (unsigned char*)c->ws_cap.p;
  return cap_emit(df, ws, N, (const long long*)(ws + off_off), tris, st);
}

int dudf_evaluate_host(dudf_ctx* c, const float* x_host, int64_t N, int order, double* f_host, double* g_host,
                       double* H_host, int64_t max_batch, int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_evaluate_host: weights not set");
  DUDF_REQUIRE(order >= 0 && order <= 2, "dudf_evaluate_host: order %d", order);
  if (N <= 0) return 0;
  // Chunks of at most max_batch points (the reference's memory cap, src/evaluate.py:13) run through a two-slot pipeline: while the
  // kernels of chunk k run on the caller's stream, the fp64 results of chunk k - 1 cross PCIe into pinned staging on a copy stream
  // and the host moves those of chunk k - 2 into the caller's (pageable) arrays.  Pageable destinations made every copy synchronous
  // and serialised transfer and compute (2 M points: 40-50 M queries/s against the kernel's 107 M).
  constexpr int64_t PIPE = 1 << 18;
  if (max_batch <= 0) max_batch = 1 << 20;
  int64_t B = max_batch < N ? max_batch : N;
  if (B > PIPE && N > PIPE) B = PIPE;
  // per slot: fp32 chunk x[3B] f[B] g[3B] H[9B]; fp64 chunk f[B] g[3B] H[9B]; pinned staging mirrors the fp64 chunk
  const size_t f32_slot = (size_t)B * 16 * sizeof(float), f64_slot = (size_t)B * 13 * sizeof(double);
  if (c->ws_out.ensure(2 * f32_slot)) return 1;
  if (c->ws_x64.ensure(2 * f64_slot)) return 1;
  if (c->ev_pinned_cap < 2 * f64_slot) {
    if (c->ev_pinned) cudaFreeHost(c->ev_pinned);
    c->ev_pinned = nullptr;
    c->ev_pinned_cap = 0;
    DUDF_CUDA_OK(cudaMallocHost(&c->ev_pinned, 2 * f64_slot));
    c->ev_pinned_cap = 2 * f64_slot;
  }
  if (!c->ev_copy_stream) {
    DUDF_CUDA_OK(cudaStreamCreateWithFlags(&c->ev_copy_stream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      DUDF_CUDA_OK(cudaEventCreateWithFlags(&c->ev_ready[k], cudaEventDisableTiming));
      DUDF_CUDA_OK(cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
  }
  const int nch = order_to_nch(order);
  const bool want_g = g_host && order >= 1, want_H = H_host && order >= 2;
  cudaStream_t st = (cudaStream_t)stream;      // the caller's stream: ordered behind its weight refresh / parameter updates
  int64_t heads[2] = {0, 0}, counts[2] = {0, 0};
  auto drain = [&](int slot) -> int {          // results of the chunk in `slot`: pinned staging -> the caller's arrays
    if (counts[slot] == 0) return 0;
    DUDF_CUDA_OK(cudaEventSynchronize(c->ev_done[slot]));
    const double* src = reinterpret_cast<const double*>(static_cast<char*>(c->ev_pinned) + slot * f64_slot);
    const int64_t n = counts[slot], head = heads[slot];
    if (f_host) memcpy(f_host + head, src, (size_t)n * sizeof(double));
    if (want_g) memcpy(g_host + head * 3, src + B, (size_t)n * 3 * sizeof(double));
    if (want_H) memcpy(H_host + head * 9, src + 4 * B, (size_t)n * 9 * sizeof(double));
    counts[slot] = 0;
    return 0;
  };
  int64_t k = 0;
  for (int64_t head = 0; head < N; head += B, ++k) {
    const int slot = (int)(k & 1);
    int rc = drain(slot);                      // chunk k - 2 has left this slot's device and pinned buffers
    if (rc) return rc;
    const int64_t n = (N - head < B) ? N - head : B;
    float* xd = reinterpret_cast<float*>(static_cast<char*>(c->ws_out.p) + slot * f32_slot);
    float *fd = xd + 3 * B, *gd = fd + B, *Hd = gd + 3 * B;
    double* f64 = reinterpret_cast<double*>(static_cast<char*>(c->ws_x64.p) + slot * f64_slot);
    double *g64 = f64 + B, *H64 = g64 + 3 * B;
    double* pin = reinterpret_cast<double*>(static_cast<char*>(c->ev_pinned) + slot * f64_slot);
    DUDF_CUDA_OK(cudaMemcpyAsync(xd, x_host + head * 3, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    QueryOut o{fd, order >= 1 ? gd : nullptr, order >= 2 ? Hd : nullptr, nullptr, nullptr, 0, 0.f};
    if ((rc = run_forward(c, nch, xd, n, 0, 0, o, precision, st))) return rc;
    if (f_host && (rc = f32_to_f64(fd, f64, n, st))) return rc;
    if (want_g && (rc = f32_to_f64(gd, g64, n * 3, st))) return rc;
    if (want_H && (rc = f32_to_f64(Hd, H64, n * 9, st))) return rc;
    DUDF_CUDA_OK(cudaEventRecord(c->ev_ready[slot], st));
    DUDF_CUDA_OK(cudaStreamWaitEvent(c->ev_copy_stream, c->ev_ready[slot], 0));
    if (f_host) DUDF_CUDA_OK(cudaMemcpyAsync(pin, f64, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->ev_copy_stream));
    if (want_g) DUDF_CUDA_OK(cudaMemcpyAsync(pin + B, g64, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->ev_copy_stream));
    if (want_H) DUDF_CUDA_OK(cudaMemcpyAsync(pin + 4 * B, H64, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, c->ev_copy_stream));
    DUDF_CUDA_OK(cudaEventRecord(c->ev_done[slot], c->ev_copy_stream));
    heads[slot] = head;
    counts[slot] = n;
  }
  int rc = drain((int)(k & 1));                // the older of the two chunks in flight first
  if (rc) return rc;
  if ((rc = drain((int)((k + 1) & 1)))) return rc;
  DUDF_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// training primitives: jet forward with stash, loss epilogues, reverse sweep, weight gradients.
// The stashes Z (pre-activations), A (activations) and Zb (pre-activation adjoints) are caller-owned
// fp32 arrays [n_hidden][256][ld]; a call works on the column range starting at col0.
// ---------------------------------------------------------------------------------------------
int64_t dudf_stash_columns(int order, int64_t P, int precision) {
  const int nch = order_to_nch(order);
  if (nch < 0 || nch > 10 || P < 0) return -1;
  if (precision == DUDF_PRECISION_TC16 || precision == DUDF_PRECISION_TCX3) {
    const int pp = tc_train_pair_points(nch), pc = tc_train_pair_cols(nch);
    return (P + pp - 1) / pp * pc;
  }
  const int pt = simt_tile_points(nch), nc = simt_tile_cols(nch);
  return (P + pt - 1) / pt * nc;
}

static int fill_grad_view(dudf_ctx* c, float* const* gW, float* const* gb, GradView& gv, const char* who) {
  memset(&gv, 0, sizeof(gv));
  for (int i = 0; i < c->n_lin; ++i) {
    DUDF_REQUIRE(gW[i] && (!gb || gb[i]), "%s: null gradient pointer for layer %d", who, i);
    gv.W[i] = gW[i];
    gv.b[i] = gb ? gb[i] : nullptr;
  }
  return 0;
}

// validates a segment list and, for the tensor-core path, that the segments occupy consecutive stash columns
static int check_segments(const dudf_segment* segs, int nseg, int64_t ld, int precision, bool forward, const char* who) {
  DUDF_REQUIRE(segs && nseg >= 1 && nseg <= 2, "%s: 1 or 2 segments", who);
  DUDF_REQUIRE(precision >= DUDF_PRECISION_FP32 && precision <= DUDF_PRECISION_TCX3, "%s: unknown precision %d", who, precision);
  int64_t next = -1;
  for (int i = 0; i < nseg; ++i) {
    const dudf_segment& s = segs[i];
    DUDF_REQUIRE(s.order >= 0 && s.order <= 2, "%s: order %d (0..2)", who, s.order);
    DUDF_REQUIRE(s.rows > 0 && s.x && (forward ? (s.packed != nullptr) : (s.seeds != nullptr)), "%s: empty or null segment", who);
    const int64_t cols = dudf_stash_columns(s.order, s.rows, precision);
    DUDF_REQUIRE(s.col0 % 4 == 0 && s.col0 + cols <= ld, "%s: stash too small", who);
    if (precision != DUDF_PRECISION_FP32 && i > 0)
      DUDF_REQUIRE(s.col0 == next && segs[0].order == 2, "%s: tensor-core segments must be contiguous, Hessian segment first", who);
    next = s.col0 + cols;
  }
  return 0;
}

int dudf_jet_forward_multi(dudf_ctx* c, const dudf_segment* segs, int nseg, void* Z, void* A, int64_t ld, int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_jet_forward: weights not set");
  DUDF_REQUIRE(Z && A && ld % 4 == 0, "dudf_jet_forward: null or misaligned stash");
  int rc = check_segments(segs, nseg, ld, precision, true, "dudf_jet_forward");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != DUDF_PRECISION_FP32) {
    TcSegment ts[2];
    for (int i = 0; i < nseg; ++i) ts[i] = TcSegment{segs[i].x, segs[i].rows, order_to_nch(segs[i].order), segs[i].packed, nullptr};
    if (precision == DUDF_PRECISION_TCX3) return tcx_train_forward(c->tcx_packed, c->view(), ts, nseg, (float*)Z, A, ld, segs[0].col0, c->sms, st);
    return tc_train_forward(c->tc_packed, c->view(), ts, nseg, (float*)Z, A, ld, segs[0].col0, c->sms, st);
  }
  for (int i = 0; i < nseg; ++i) {
    QueryOut o{nullptr, nullptr, nullptr, nullptr, segs[i].packed, 0, 0.f};
    rc = simt_forward(c->view(), order_to_nch(segs[i].order), segs[i].x, segs[i].rows, 0, 0, o, (float*)Z, (float*)A, ld, segs[i].col0, c->sms, st);
    if (rc) return rc;
  }
  return 0;
}

int dudf_jet_backward_multi(dudf_ctx* c, const dudf_segment* segs, int nseg, const float* seed_absmax, const void* Z, void* Zb, int64_t ld,
                            float* const* gW, float* const* gb, int precision, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_jet_backward: weights not set");
  DUDF_REQUIRE(Z && Zb && gW && gb, "dudf_jet_backward: null argument");
  int rc = check_segments(segs, nseg, ld, precision, false, "dudf_jet_backward");
  if (rc) return rc;
  GradView gv;
  rc = fill_grad_view(c, gW, gb, gv, "dudf_jet_backward");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision != DUDF_PRECISION_FP32) {      // TCX3: split forward, single-pass reverse sweep (tools/precision_study.py)
    TcSegment ts[2];
    for (int i = 0; i < nseg; ++i) ts[i] = TcSegment{segs[i].x, segs[i].rows, order_to_nch(segs[i].order), nullptr, segs[i].seeds};
    return tc_train_backward(c->tc_packed, c->view(), gv, ts, nseg, seed_absmax, (const float*)Z, Zb, ld, segs[0].col0, c->sms, st);
  }
  for (int i = 0; i < nseg; ++i) {
    rc = simt_backward(c->view(), gv, order_to_nch(segs[i].order), segs[i].x, segs[i].rows, segs[i].seeds, (const float*)Z, (float*)Zb, ld,
                       segs[i].col0, c->sms, st);
    if (rc) return rc;
  }
  return 0;
}

int dudf_jet_forward(dudf_ctx* c, const float* x, int64_t P, int order, float* packed, void* Z, void* A, int64_t ld,
                     int64_t col0, int precision, void* stream) {
  if (P <= 0) return 0;
  dudf_segment s{x, P, order, packed, nullptr, col0};
  return dudf_jet_forward_multi(c, &s, 1, Z, A, ld, precision, stream);
}

int dudf_jet_backward(dudf_ctx* c, const float* x, int64_t P, int order, const float* seeds, const float* seed_absmax, const void* Z,
                      void* Zb, int64_t ld, int64_t col0, float* const* gW, float* const* gb, int precision, void* stream) {
  if (P <= 0) return 0;
  dudf_segment s{x, P, order, nullptr, seeds, col0};
  return dudf_jet_backward_multi(c, &s, 1, seed_absmax, Z, Zb, ld, gW, gb, precision, stream);
}

int dudf_jet_wgrad_layers(dudf_ctx* c, const void* Zb, const void* A, int64_t ld, const float* seed_absmax, float* const* gW, int layer_lo,
                          int layer_hi, int precision, void* stream) {
  DUDF_REQUIRE(c && Zb && A && gW, "dudf_jet_wgrad_layers: null argument");
  DUDF_REQUIRE(precision == DUDF_PRECISION_TC16 || precision == DUDF_PRECISION_TCX3, "dudf_jet_wgrad_layers: tensor-core precisions only");
  DUDF_REQUIRE(layer_lo >= 1 && layer_hi <= c->n_lin - 1 && layer_lo <= layer_hi, "dudf_jet_wgrad_layers: layer range [%d, %d) outside [1, %d)",
               layer_lo, layer_hi, c->n_lin - 1);
  GradView gv;
  int rc = fill_grad_view(c, gW, nullptr, gv, "dudf_jet_wgrad_layers");
  if (rc) return rc;
  return tc_train_wgrad(c->view(), gv, Zb, A, ld, seed_absmax, c->sms, (cudaStream_t)stream, layer_lo, layer_hi);
}

int dudf_jet_wgrad(dudf_ctx* c, const void* Zb, const void* A, int64_t ld, int64_t ncols, const float* seed_absmax, float* const* gW,
                   int precision, void* stream) {
  DUDF_REQUIRE(c && Zb && A && gW, "dudf_jet_wgrad: null argument");
  DUDF_REQUIRE(precision >= DUDF_PRECISION_FP32 && precision <= DUDF_PRECISION_TCX3, "dudf_jet_wgrad: unknown precision %d", precision);
  GradView gv;
  int rc = fill_grad_view(c, gW, nullptr, gv, "dudf_jet_wgrad");
  if (rc) return rc;
  if (precision != DUDF_PRECISION_FP32) return tc_train_wgrad(c->view(), gv, Zb, A, ld, seed_absmax, c->sms, (cudaStream_t)stream);
  return simt_wgrad(c->view(), gv, (const float*)Zb, (const float*)A, ld, ncols, c->sms, (cudaStream_t)stream);
}

int64_t dudf_fused_scratch_bytes(const dudf_ctx* c) {
  if (!c) return -1;
  return (int64_t)tc_fused_scratch_bytes(c->view(), c->sms);
}

int dudf_train_step_fused(dudf_ctx* c, int mode, const dudf_train_segment* segs, int nseg, int64_t P_global, const float* w_host, float alpha,
                          double* terms, const float* amax_prev, float* amax_next, void* scratch, void* A, void* Zb, int64_t ld,
                          float* const* gW, float* const* gb, int flags, void* stream) {
  DUDF_REQUIRE(c && c->weights_set, "dudf_train_step_fused: weights not set");
  DUDF_REQUIRE(segs && nseg >= 1 && nseg <= 2 && w_host && terms && amax_prev && amax_next && scratch && A && Zb && gW && gb,
               "dudf_train_step_fused: null argument");
  DUDF_REQUIRE(P_global > 0, "dudf_train_step_fused: P_global must be positive");
  GradView gv;
  int rc = fill_grad_view(c, gW, gb, gv, "dudf_train_step_fused");
  if (rc) return rc;
  TcSegment ts[2];
  for (int i = 0; i < nseg; ++i) {
    const dudf_train_segment& s = segs[i];
    DUDF_REQUIRE(s.rows > 0 && s.x && s.normals && s.dist, "dudf_train_step_fused: empty or null segment");
    DUDF_REQUIRE(s.order == 1 || s.order == 2, "dudf_train_step_fused: order %d (1 or 2)", s.order);
    DUDF_REQUIRE(i == 0 || segs[0].order == 2, "dudf_train_step_fused: Hessian segment first");
    ts[i] = TcSegment{s.x, s.rows, order_to_nch(s.order), s.packed, nullptr, s.normals, s.dist};
  }
  TcFusedLoss fl;
  fl.mode = mode; fl.alpha = alpha; fl.P_global = P_global; fl.terms = terms; fl.amax_prev = amax_prev; fl.amax_next = amax_next;
  fl.flags = flags;
  for (int k = 0; k < 4; ++k) fl.w[k] = w_host[k];
  cudaStream_t st = (cudaStream_t)stream;
  rc = tc_train_fused(c->tc_packed, c->view(), gv, ts, nseg, fl, (float*)scratch, A, Zb, ld, c->sms, st);
  if (rc || (flags & DUDF_FUSED_NO_WGRAD)) return rc;
  return tc_train_wgrad(c->view(), gv, Zb, A, ld, amax_prev, c->sms, st);
}

int dudf_loss(int mode, const float* packed, int nch, const float* normals, const float* dist, int64_t P, int64_t P_global,
              const float* w_host, float alpha, const float* upstream, float* seeds, float* seed_absmax, double* terms,
              double* s2_stats, void* stream) {
  DUDF_REQUIRE(packed && dist && w_host, "dudf_loss: null argument");
  DUDF_REQUIRE(mode == DUDF_LOSS_S1 || mode == DUDF_LOSS_S2 || mode == DUDF_LOSS_SIREN, "dudf_loss: unknown mode %d", mode);
  DUDF_REQUIRE(nch == 1 || nch == 4 || nch == 10, "dudf_loss: nch %d", nch);
  DUDF_REQUIRE(mode != DUDF_LOSS_SIREN || nch >= 4, "dudf_loss: loss_siren needs gradients");
  DUDF_REQUIRE(normals || (mode == DUDF_LOSS_S2), "dudf_loss: normals required");
  LossArgs a;
  a.mode = mode; a.packed = packed; a.nch = nch; a.normals = normals; a.dist = dist; a.P = P; a.P_global = P_global;
  for (int k = 0; k < 4; ++k) a.w[k] = w_host[k];
  a.alpha = alpha; a.upstream = upstream; a.seeds = seeds; a.terms = terms; a.s2_stats = s2_stats; a.seed_absmax = seed_absmax;
  return loss_seeds(a, (cudaStream_t)stream);
}

int dudf_loss_s2_stats(const float* packed, const float* dist, int64_t P, double* stats, void* stream) {
  DUDF_REQUIRE(packed && dist && stats, "dudf_loss_s2_stats: null argument");
  return loss_s2_stats(packed, dist, P, stats, (cudaStream_t)stream);
}

int dudf_loss_s2_finish(const double* stats, float w0, float w1, double* terms, void* stream) {
  DUDF_REQUIRE(stats && terms, "dudf_loss_s2_finish: null argument");
  return s2_finish(stats, w0, w1, terms, (cudaStream_t)stream);
}

static int device_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int dudf_sample_batch_pc(const float* surf_pts, const float* surf_normals, int64_t n_surf, int64_t n_on, int64_t n_far, int64_t n_near,
                         float sigma, const float* lo_host, const float* hi_host, uint64_t seed, uint64_t batch_index,
                         const int64_t* on_idx, const float* far_pts, const int64_t* near_idx, const float* near_off, float* coords,
                         float* normals, float* dist, void* stream) {
  DUDF_REQUIRE(surf_pts && surf_normals && coords && normals && dist, "dudf_sample_batch_pc: null argument");
  DUDF_REQUIRE(n_surf > 0 && n_on >= 0 && n_far >= 0 && n_near >= 0, "dudf_sample_batch_pc: bad sizes");
  DUDF_REQUIRE(n_near == 0 || n_on > 0, "dudf_sample_batch_pc: near rows are displaced ON rows; n_on must be positive");
  SampleArgs a;
  a.surf_pts = surf_pts; a.surf_nrm = surf_normals; a.n_surf = n_surf; a.n_on = n_on; a.n_far = n_far; a.n_near = n_near;
  a.sigma = sigma; a.seed = seed; a.batch = batch_index;
  for (int k = 0; k < 3; ++k) { a.lo[k] = lo_host ? lo_host[k] : -1.f; a.hi[k] = hi_host ? hi_host[k] : 1.f; }
  a.on_idx = on_idx; a.far_pts = far_pts; a.near_idx = near_idx; a.near_off = near_off;
  a.coords = coords; a.normals = normals; a.dist = dist;
  return sample_batch_pc(a, device_sms(), (cudaStream_t)stream);
}

int64_t dudf_cloud_index_bytes(int64_t n_x) { return cloud_index_bytes(n_x); }

int dudf_cloud_index_build(const float* cloud, int64_t n_x, void* index, void* stream) {
  DUDF_REQUIRE(cloud && index, "dudf_cloud_index_build: null argument");
  return cloud_index_build(cloud, n_x, index, (cudaStream_t)stream);
}

int dudf_nearest_distance_indexed(const float* queries, int64_t n_q, const void* index, int64_t n_x, float* dist, void* stream) {
  DUDF_REQUIRE(queries && index && dist, "dudf_nearest_distance_indexed: null argument");
  return cloud_index_query(queries, n_q, index, n_x, dist, (cudaStream_t)stream);
}

int dudf_sample_batch_pc_indexed(const float* surf_pts, const float* surf_normals, int64_t n_surf, const void* index, int64_t n_on,
                                 int64_t n_far, int64_t n_near, float sigma, const float* lo_host, const float* hi_host, uint64_t seed,
                                 uint64_t batch_index, const int64_t* on_idx, const float* far_pts, const int64_t* near_idx,
                                 const float* near_off, float* coords, float* normals, float* dist, void* stream) {
  DUDF_REQUIRE(surf_pts && surf_normals && index && coords && normals && dist, "dudf_sample_batch_pc_indexed: null argument");
  DUDF_REQUIRE(n_surf > 0 && n_on >= 0 && n_far >= 0 && n_near >= 0, "dudf_sample_batch_pc_indexed: bad sizes");
  DUDF_REQUIRE(n_near == 0 || n_on > 0, "dudf_sample_batch_pc_indexed: near rows are displaced ON rows; n_on must be positive");
  SampleArgs a;
  a.surf_pts = surf_pts; a.surf_nrm = surf_normals; a.n_surf = n_surf; a.n_on = n_on; a.n_far = n_far; a.n_near = n_near;
  a.sigma = sigma; a.seed = seed; a.batch = batch_index;
  for (int k = 0; k < 3; ++k) { a.lo[k] = lo_host ? lo_host[k] : -1.f; a.hi[k] = hi_host ? hi_host[k] : 1.f; }
  a.on_idx = on_idx; a.far_pts = far_pts; a.near_idx = near_idx; a.near_off = near_off;
  a.coords = coords; a.normals = normals; a.dist = dist;
  return sample_batch_pc_indexed(a, index, (cudaStream_t)stream);
}

int dudf_mesh_distance(const float* queries, int64_t n_q, const float* triangles, int64_t n_tri, float* dist, void* stream) {
  DUDF_REQUIRE(queries && triangles && dist, "dudf_mesh_distance: null argument");
  return mesh_distance(queries, n_q, triangles, n_tri, dist, device_sms(), (cudaStream_t)stream);
}

int dudf_sample_batch_mesh(const float* surf_pts, const float* surf_normals, int64_t n_surf, const float* triangles, int64_t n_tri,
                           int64_t n_on, int64_t n_far, int64_t n_near, float sigma, const float* lo_host, const float* hi_host,
                           uint64_t seed, uint64_t batch_index, const int64_t* on_idx, const float* far_pts, const int64_t* near_idx,
                           const float* near_off, float* coords, float* normals, float* dist, void* stream) {
  DUDF_REQUIRE(surf_pts && surf_normals && triangles && coords && normals && dist, "dudf_sample_batch_mesh: null argument");
  DUDF_REQUIRE(n_surf > 0 && n_tri > 0 && n_on >= 0 && n_far >= 0 && n_near >= 0, "dudf_sample_batch_mesh: bad sizes");
  DUDF_REQUIRE(n_near == 0 || n_on > 0, "dudf_sample_batch_mesh: near rows are displaced ON rows; n_on must be positive");
  SampleArgs a;
  a.surf_pts = surf_pts; a.surf_nrm = surf_normals; a.n_surf = n_surf; a.n_on = n_on; a.n_far = n_far; a.n_near = n_near;
  a.sigma = sigma; a.seed = seed; a.batch = batch_index;
  for (int k = 0; k < 3; ++k) { a.lo[k] = lo_host ? lo_host[k] : -1.f; a.hi[k] = hi_host ? hi_host[k] : 1.f; }
  a.on_idx = on_idx; a.far_pts = far_pts; a.near_idx = near_idx; a.near_off = near_off;
  a.coords = coords; a.normals = normals; a.dist = dist;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = sample_rows(a, st);
  if (rc) return rc;
  return mesh_distance(coords + n_on * 3, n_far + n_near, triangles, n_tri, dist + n_on, device_sms(), st);
}

int dudf_mesh_sample_surface(const float* triangles, const float* cdf, int64_t n_tri, int64_t n, uint64_t seed, const float* draws,
                             float* points, float* normals, void* stream) {
  DUDF_REQUIRE(triangles && cdf && points && normals, "dudf_mesh_sample_surface: null argument");
  return mesh_sample_surface(triangles, cdf, n_tri, n, seed, draws, points, normals, (cudaStream_t)stream);
}

int dudf_nearest_distance(const float* queries, int64_t n_q, const float* cloud, int64_t n_x, float* dist, void* stream) {
  DUDF_REQUIRE(queries && cloud && dist, "dudf_nearest_distance: null argument");
  return nn_distance(queries, n_q, cloud, n_x, dist, device_sms(), (cudaStream_t)stream);
}

int dudf_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   int64_t t, void* stream) {
  DUDF_REQUIRE(p && g && m && v, "dudf_adam_step: null argument");
  DUDF_REQUIRE(t >= 1, "dudf_adam_step: step count must be >= 1");
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, t, (cudaStream_t)stream);
}

int dudf_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2, float eps,
                       void* stream) {
  DUDF_REQUIRE(p && g && m && v && state, "dudf_adam_step_dev: null argument");
  DUDF_REQUIRE(((uintptr_t)state & 7) == 0, "dudf_adam_step_dev: state must be 8-byte aligned");
  return adam_step_dev(p, g, m, v, n, state, beta1, beta2, eps, (cudaStream_t)stream);
}

int dudf_adam_step_guarded(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                           int64_t t, const float* unsafe_flag, int64_t* skipped, void* stream) {
  DUDF_REQUIRE(p && g && m && v && unsafe_flag, "dudf_adam_step_guarded: null argument");
  DUDF_REQUIRE(t >= 1, "dudf_adam_step_guarded: step count must be >= 1");
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, t, (cudaStream_t)stream, unsafe_flag, (long long*)skipped);
}

int dudf_adam_step_peers(float* p, const float* const* peer_grads, int world, float* m, float* v, int64_t n, float lr, float beta1,
                         float beta2, float eps, int64_t t, int guarded, int64_t* skipped, float* g_sum_out, void* stream) {
  DUDF_REQUIRE(p && peer_grads && m && v, "dudf_adam_step_peers: null argument");
  DUDF_REQUIRE(t >= 1, "dudf_adam_step_peers: step count must be >= 1");
  return adam_step_peers(p, peer_grads, world, m, v, n, lr, beta1, beta2, eps, t, guarded, (long long*)skipped, g_sum_out, (cudaStream_t)stream);
}

int dudf_scale_guard(const float* amax_prev, const float* amax_next, float limit, float* flag, void* stream) {
  DUDF_REQUIRE(amax_prev && amax_next && flag, "dudf_scale_guard: null argument");
  return scale_guard(amax_prev, amax_next, limit, flag, (cudaStream_t)stream);
}

int dudf_selftest_umma(int variant, float* max_err_host) {
  DUDF_REQUIRE(max_err_host != nullptr, "dudf_selftest_umma: null output");
  return tc_selftest(variant, max_err_host, 0);
}

int dudf_bench_umma(int variant, int ctas, int iters, float* clk_per_mma_host) {
  DUDF_REQUIRE(clk_per_mma_host != nullptr, "dudf_bench_umma: null output");
  return tc_mma_bench(variant, ctas, iters, clk_per_mma_host, 0);
}

int dudf_debug_set_trace(void* device_buffer) {
  tc_set_trace((unsigned long long*)device_buffer);
  return 0;
}

}  // extern "C"

// Spatial index of a point cloud for the batch sampler's nearest-cloud-point distance (SURVEY.md §8f row 1;
// shortestDistance, src/dataset.py:72-78, called for the n_far uniform rows of every batch, :116-118).
// The reference forms the whole n_far x n_cloud matrix; dudf_sampler.cu's brute-force kernels stream the cloud once per 1 024
// queries (2e9 pair distances per 29 970-row batch of a 200 000-point cloud: 0.4-0.5 ms, as long as a third of the training step).
// Here the cloud is sorted ONCE along a Morton curve and covered by a 32-ary hierarchy of axis-aligned boxes; a query is served by
// one warp that walks the hierarchy best-first: the 32 lanes measure the 32 children of a node, the warp descends into the closest
// child that can still beat the running minimum.  ~10^2 node visits instead of 2e5 pair distances per query.
//
// The result is EXACT in fp32: point distances are the difference form fma(dx, dx, fma(dy, dy, dz dz)); a box's lower bound is the
// same expression on max(lo - p, p - hi, 0), and since fp32 subtraction, multiplication and fma are monotone, the bound never
// exceeds the computed distance of a point inside the box — pruning on `bound >= best` cannot discard the minimiser.
//
// Layout of the caller-owned index buffer (cloud_index_bytes(n)): [64-byte header | float4 pts[n32] | level 1 boxes | level 2 ...],
// pts = (x, y, z, original row as int bits) in Morton order, padded to a multiple of 1 024 with points at 1e18; boxes = (lo, hi) float4
// pairs, level k holds ceil(count_{k-1} / 32) boxes padded to a multiple of 32 with inverted (empty) boxes; the top level is the
// first with at most 32 boxes.
#include <cub/device/device_radix_sort.cuh>

#include "dudf_common.cuh"
#include "dudf_kernels.h"

namespace dudf {

namespace {

constexpr float IDX_FAR = 1.0e18f;      // padding coordinate: (1e18)^2 * 3 is finite in fp32 and never wins a minimum

struct IndexView {
  const float4* pts;
  const float4* box[CLOUD_INDEX_MAX_LEVELS + 1];   // box[k]: level k, k = 1 .. top
  int top;
};
struct IndexLayout {
  int64_t n32;                                   // padded points
  int64_t count[CLOUD_INDEX_MAX_LEVELS + 1];     // boxes per level (count[0] = n)
  int64_t padded[CLOUD_INDEX_MAX_LEVELS + 1];
  size_t off_pts, off_box[CLOUD_INDEX_MAX_LEVELS + 1], bytes;
  int top;
};

int layout_of(int64_t n, IndexLayout& l) {
  l.n32 = ((n + 31) / 32 + 31) / 32 * 1024;      // every leaf of the padded leaf level has its 32 points (padding: points at 1e18)
  l.count[0] = n;
  l.padded[0] = l.n32;
  l.off_pts = 64;
  size_t off = l.off_pts + (size_t)l.n32 * sizeof(float4);
  for (int k = 1; k <= CLOUD_INDEX_MAX_LEVELS; ++k) {
    l.count[k] = (l.count[k - 1] + 31) / 32;
    l.padded[k] = (l.count[k] + 31) / 32 * 32;
    l.off_box[k] = off;
    off += (size_t)l.padded[k] * 2 * sizeof(float4);
    if (l.count[k] <= 32) {
      l.top = k;
      l.bytes = off;
      return 0;
    }
  }
  return 1;
}

IndexView view_of(const void* index, const IndexLayout& l) {
  IndexView v;
  const char* base = static_cast<const char*>(index);
  v.pts = reinterpret_cast<const float4*>(base + l.off_pts);
  for (int k = 0; k <= CLOUD_INDEX_MAX_LEVELS; ++k) v.box[k] = nullptr;
  for (int k = 1; k <= l.top; ++k) v.box[k] = reinterpret_cast<const float4*>(base + l.off_box[k]);
  v.top = l.top;
  return v;
}

// floats <-> unsigned integers of the same order (for atomicMin / atomicMax on the bounding box)
__device__ __forceinline__ unsigned int ord_of(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_of(unsigned int o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

__global__ void idx_init_kernel(unsigned int* hdr) {
  if (threadIdx.x < 3) hdr[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) hdr[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) idx_bounds_kernel(const float* __restrict__ X, int64_t n, unsigned int* __restrict__ hdr) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float v = X[i * 3 + k];
      lo[k] = fminf(lo[k], v);
      hi[k] = fmaxf(hi[k], v);
    }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&hdr[k], ord_of(lo[k]));
      atomicMax(&hdr[3 + k], ord_of(hi[k]));
    }
  }
}

__device__ __forceinline__ unsigned int spread10(unsigned int v) {   // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) idx_morton_kernel(const float* __restrict__ X, int64_t n, const unsigned int* __restrict__ hdr,
                                                         unsigned int* __restrict__ key, int* __restrict__ row) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float lo = float_of(hdr[k]), hi = float_of(hdr[3 + k]);
    const float t = hi > lo ? (X[i * 3 + k] - lo) / (hi - lo) : 0.f;
    c[k] = (unsigned int)fminf(fmaxf(t * 1024.f, 0.f), 1023.f);
  }
  key[i] = (spread10(c[0]) << 2) | (spread10(c[1]) << 1) | spread10(c[2]);
  row[i] = (int)i;
}

__global__ void __launch_bounds__(256) idx_gather_kernel(const float* __restrict__ X, int64_t n, int64_t n32, const int* __restrict__ row,
                                                         float4* __restrict__ pts) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n32) return;
  float4 v = make_float4(IDX_FAR, IDX_FAR, IDX_FAR, __int_as_float(-1));
  if (j < n) {
    const int64_t r = row[j];
    v = make_float4(X[r * 3], X[r * 3 + 1], X[r * 3 + 2], __int_as_float((int)r));
  }
  pts[j] = v;
}

// one warp per box of level `k`: union of its 32 children (points for k = 1, boxes of level k - 1 otherwise)
__global__ void __launch_bounds__(256) idx_boxes_kernel(const float4* __restrict__ child, int64_t n_child_valid, int64_t n_child_padded,
                                                        int leaf, float4* __restrict__ box, int64_t n_box_padded) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_box_padded) return;
  const int64_t c = w * 32 + lane;
  float lo[3] = {IDX_FAR, IDX_FAR, IDX_FAR}, hi[3] = {-IDX_FAR, -IDX_FAR, -IDX_FAR};
  if (leaf) {
    if (c < n_child_valid) {
      const float4 p = child[c];
      lo[0] = hi[0] = p.x; lo[1] = hi[1] = p.y; lo[2] = hi[2] = p.z;
    }
  } else if (c < n_child_padded) {
    const float4 a = child[2 * c], b = child[2 * c + 1];
    lo[0] = a.x; lo[1] = a.y; lo[2] = a.z;
    hi[0] = b.x; hi[1] = b.y; hi[2] = b.z;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], s));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], s));
    }
  if (lane == 0) {
    box[2 * w] = make_float4(lo[0], lo[1], lo[2], 0.f);
    box[2 * w + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
}

__device__ __forceinline__ float sq3(float a, float b, float c) { return fmaf(a, a, fmaf(b, b, c * c)); }

// pops the remaining child with the smallest bound below `best` (-1: none left); `key` is this lane's bound, all-ones once taken
__device__ __forceinline__ int idx_pop(unsigned int& key, float best, int lane) {
  const unsigned int k = key < __float_as_uint(best) ? key : 0xffffffffu;
  const unsigned int m = __reduce_min_sync(0xffffffffu, k);
  if (m == 0xffffffffu) return -1;
  const int c = __ffs(__ballot_sync(0xffffffffu, k == m)) - 1;
  if (lane == c) key = 0xffffffffu;
  return c;
}

// Best-first walk of the children of `node` (a box of level LEVEL; LEVEL == top + 1 is the virtual root over the top level).
// `best` is warp-uniform.  A walk is a chain of dependent loads (boxes -> chosen child -> its boxes ...), one warp per query and one
// wave of warps per batch, so its latency is the kernel's duration: above the leaves the walk descends into the closest child first
// (that fixes a tight `best`), at the leaf level it then takes the surviving leaves FOUR at a time (four independent loads in flight).
template <int LEVEL>
__device__ __forceinline__ void idx_visit(const IndexView& v, int64_t node, float px, float py, float pz, float& best, int lane) {
  const float4 lo = v.box[LEVEL - 1][2 * (node * 32 + lane)], hi = v.box[LEVEL - 1][2 * (node * 32 + lane) + 1];
  const float lb = sq3(fmaxf(fmaxf(lo.x - px, px - hi.x), 0.f), fmaxf(fmaxf(lo.y - py, py - hi.y), 0.f), fmaxf(fmaxf(lo.z - pz, pz - hi.z), 0.f));
  unsigned int key = __float_as_uint(lb);                          // lb >= 0: bit patterns order like floats
  if constexpr (LEVEL == 2) {
    bool first = true;
    while (true) {
      const int c0 = idx_pop(key, best, lane);
      if (c0 < 0) break;                                           // no remaining leaf can beat the running minimum
      int c1 = c0, c2 = c0, c3 = c0;
      if (!first) {
        const int a = idx_pop(key, best, lane);
        if (a >= 0) {
          c1 = a;
          const int b = idx_pop(key, best, lane);
          if (b >= 0) {
            c2 = b;
            const int c = idx_pop(key, best, lane);
            if (c >= 0) c3 = c;
          }
        }
      }
      first = false;
      const float4* leaf = v.pts + node * 1024 + lane;
      const float4 p0 = leaf[c0 * 32], p1 = leaf[c1 * 32], p2 = leaf[c2 * 32], p3 = leaf[c3 * 32];
      const float d2 = fminf(fminf(sq3(px - p0.x, py - p0.y, pz - p0.z), sq3(px - p1.x, py - p1.y, pz - p1.z)),
                             fminf(sq3(px - p2.x, py - p2.y, pz - p2.z), sq3(px - p3.x, py - p3.y, pz - p3.z)));
      best = fminf(best, __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(d2))));
    }
  } else {
    while (true) {
      const int c = idx_pop(key, best, lane);
      if (c < 0) break;
      idx_visit<LEVEL - 1>(v, node * 32 + c, px, py, pz, best, lane);
    }
  }
}

__global__ void __launch_bounds__(256) idx_query_kernel(IndexView v, const float* __restrict__ q, int64_t nq, float* __restrict__ dist) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nq) return;
  const float px = q[w * 3], py = q[w * 3 + 1], pz = q[w * 3 + 2];
  float best = 3.0e38f;
  switch (v.top) {
    case 1: idx_visit<2>(v, 0, px, py, pz, best, lane); break;
    case 2: idx_visit<3>(v, 0, px, py, pz, best, lane); break;
    case 3: idx_visit<4>(v, 0, px, py, pz, best, lane); break;
    case 4: idx_visit<5>(v, 0, px, py, pz, best, lane); break;
    default: idx_visit<6>(v, 0, px, py, pz, best, lane); break;
  }
  if (lane == 0) dist[w] = sqrtf(best);
}

}  // namespace

int64_t cloud_index_bytes(int64_t n) {
  IndexLayout l;
  if (n <= 0 || n > CLOUD_INDEX_MAX_POINTS || layout_of(n, l)) return -1;
  return (int64_t)l.bytes;
}

int cloud_index_build(const float* X, int64_t n, void* index, cudaStream_t st) {
  IndexLayout l;
  DUDF_REQUIRE(n > 0 && n <= CLOUD_INDEX_MAX_POINTS && layout_of(n, l) == 0, "cloud index: 1 .. 2^30 points");
  DUDF_REQUIRE((reinterpret_cast<uintptr_t>(index) & 15) == 0, "cloud index: the buffer must be 16-byte aligned");
  char* base = static_cast<char*>(index);
  unsigned int* hdr = reinterpret_cast<unsigned int*>(base);
  float4* pts = reinterpret_cast<float4*>(base + l.off_pts);
  size_t temp_bytes = 0;
  DUDF_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned int*)nullptr, (unsigned int*)nullptr, (const int*)nullptr,
                                               (int*)nullptr, (int)n, 0, 30, st));
  const size_t keys = ((size_t)n * sizeof(unsigned int) + 255) / 256 * 256;
  char* ws = nullptr;                            // stream-ordered scratch: keys in / out, rows in / out, sort temporaries
  DUDF_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&ws), 4 * keys + temp_bytes, st));
  unsigned int *key_in = reinterpret_cast<unsigned int*>(ws), *key_out = reinterpret_cast<unsigned int*>(ws + keys);
  int *row_in = reinterpret_cast<int*>(ws + 2 * keys), *row_out = reinterpret_cast<int*>(ws + 3 * keys);
  idx_init_kernel<<<1, 32, 0, st>>>(hdr);
  DUDF_LAUNCH_OK();
  const unsigned nb = (unsigned)std::min<int64_t>((n + 255) / 256, 1184);
  idx_bounds_kernel<<<nb, 256, 0, st>>>(X, n, hdr);
  DUDF_LAUNCH_OK();
  idx_morton_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(X, n, hdr, key_in, row_in);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cub::DeviceRadixSort::SortPairs(ws + 4 * keys, temp_bytes, key_in, key_out, row_in, row_out, (int)n, 0, 30, st));
  dudf_count_launch();
  idx_gather_kernel<<<(unsigned)((l.n32 + 255) / 256), 256, 0, st>>>(X, n, l.n32, row_out, pts);
  DUDF_LAUNCH_OK();
  for (int k = 1; k <= l.top; ++k) {
    float4* box = reinterpret_cast<float4*>(base + l.off_box[k]);
    const float4* child = k == 1 ? pts : reinterpret_cast<const float4*>(base + l.off_box[k - 1]);
    idx_boxes_kernel<<<(unsigned)((l.padded[k] * 32 + 255) / 256), 256, 0, st>>>(child, l.count[k - 1], l.padded[k - 1], k == 1, box, l.padded[k]);
    DUDF_LAUNCH_OK();
  }
  DUDF_CUDA_OK(cudaFreeAsync(ws, st));
  return 0;
}

int cloud_index_query(const float* q, int64_t nq, const void* index, int64_t n, float* dist, cudaStream_t st) {
  if (nq <= 0) return 0;
  IndexLayout l;
  DUDF_REQUIRE(n > 0 && n <= CLOUD_INDEX_MAX_POINTS && layout_of(n, l) == 0, "cloud index: 1 .. 2^30 points");
  DUDF_REQUIRE(nq <= ((int64_t)1 << 31) * 8 - 8, "cloud index: too many queries for one launch");
  idx_query_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(view_of(index, l), q, nq, dist);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

// CAP-UDF marching cubes on the device: extract_mesh_CAP of src/render_mc.py:201-256, the step right after the dense grid
// query.  The reference walks the (N-1)^3 cells in a Python triple loop (hours at 512^3); cells are independent:
//   * a cell is skipped when the smallest of its 8 distances exceeds the threshold (0.008);
//   * corner c is NEGATIVE when dot(grad[corner 0], grad[corner c]) < 0, its value is -ndf, otherwise +ndf;
//   * if any value is negative the signed 2x2x2 block goes through marching cubes at iso-value 0: vertices at the linear zero
//     crossing of the sign-changing edges (float64, like the float64 block the reference builds), shifted by the cell index and
//     mapped to [-1,1]^3 by v / (N-1) * 2 - 1.
// Two passes, both HBM-bound (one thread per cell, k fastest -> coalesced reads, neighbours share lines through L1/L2):
//   classify  stages the corner distances of a 4 x 8 x 32 tile of cells in shared memory, reads the gradients of near-surface
//             cells only and writes one case byte per cell; a count kernel sums the triangles of every block of 256 consecutive
//             cells and an exclusive scan of those counts gives every block its output offset;
//   emit      re-reads the case bytes (16 per thread), scans the per-thread counts inside the block and writes the triangles of
//             cell (i,j,k) in the reference's order (cells in i, j, k lexicographic order, triangles in table order).
// The case table is generated from first principles (tools/gen_mc_table.py); PyMCubes' own table is not available (parity with
// its triangulation is unpinned): results are compared as unordered triangle sets against oracle/cap_mc.py.
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>
#include "dudf_common.cuh"
#include "dudf_kernels.h"
#include "dudf_mc_table.h"

namespace dudf {

constexpr int CAP_BLOCK = 256;

__device__ __forceinline__ void cap_cell_ijk(int64_t cell, int M, int& i, int& j, int& k) {
  k = (int)(cell % M);
  j = (int)((cell / M) % M);
  i = (int)(cell / ((int64_t)M * M));
}

// classify: one block per tile of 4 x 8 x 32 cells; the 5 x 9 x 33 corner distances of the tile go through shared memory (1.45
// global loads per cell instead of 8), each thread classifies 4 cells (one per i-plane of the tile); gradients are read from
// global memory for the few cells within the threshold only
constexpr int CAP_TI = 4, CAP_TJ = 8, CAP_TK = 32;
__global__ void __launch_bounds__(CAP_BLOCK) cap_classify_kernel(const float* __restrict__ df, const float* __restrict__ vecs, int N, float thr,
                                                                 unsigned char* __restrict__ code) {
  __shared__ float sd[CAP_TI + 1][CAP_TJ + 1][CAP_TK + 1];
  const int M = N - 1;
  const int tk = (M + CAP_TK - 1) / CAP_TK, tj = (M + CAP_TJ - 1) / CAP_TJ;
  const int bk = blockIdx.x % tk, bj = (blockIdx.x / tk) % tj, bi = blockIdx.x / (tk * tj);
  const int i0 = bi * CAP_TI, j0 = bj * CAP_TJ, k0 = bk * CAP_TK;
  for (int t = threadIdx.x; t < (CAP_TI + 1) * (CAP_TJ + 1) * (CAP_TK + 1); t += CAP_BLOCK) {
    const int kk = t % (CAP_TK + 1), jj = (t / (CAP_TK + 1)) % (CAP_TJ + 1), ii = t / ((CAP_TK + 1) * (CAP_TJ + 1));
    const int i = i0 + ii, j = j0 + jj, k = k0 + kk;
    sd[ii][jj][kk] = (i < N && j < N && k < N) ? df[((int64_t)i * N + j) * N + k] : 3.4e38f;
  }
  __syncthreads();
  const int kk = threadIdx.x % CAP_TK, jj = threadIdx.x / CAP_TK;
  const int j = j0 + jj, k = k0 + kk;
  if (j >= M || k >= M) return;
#pragma unroll
  for (int ii = 0; ii < CAP_TI; ++ii) {
    const int i = i0 + ii;
    if (i >= M) break;
    float v[8];
    float vmin = 3.4e38f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = sd[ii + ((c >> 2) & 1)][jj + ((c >> 1) & 1)][kk + (c & 1)];
      vmin = fminf(vmin, v[c]);
    }
    unsigned cc = 0;
    if (!(vmin > thr)) {
      const int64_t base = ((int64_t)i * N + j) * N + k;
      const float gx = vecs[base * 3], gy = vecs[base * 3 + 1], gz = vecs[base * 3 + 2];
#pragma unroll
      for (int c = 1; c < 8; ++c) {
        const int64_t p = base + ((c >> 2) & 1) * (int64_t)N * N + ((c >> 1) & 1) * N + (c & 1);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(gx, vecs[p * 3]), __fmul_rn(gy, vecs[p * 3 + 1])), __fmul_rn(gz, vecs[p * 3 + 2]));
        if (d < 0.f && v[c] > 0.f) cc |= 1u << c;       // res = -val is negative only for val > 0 (res.min() < 0)
      }
    }
    code[((int64_t)i * M + j) * M + k] = (unsigned char)cc;
  }
}

// triangles per unit of CAP_UNIT = 4096 consecutive cells (the unit of the output order): 16 case bytes per thread, one 16-byte load
constexpr int CAP_PER_THREAD = 16;
constexpr int CAP_UNIT = CAP_BLOCK * CAP_PER_THREAD;
__device__ __forceinline__ uint4 cap_load_codes(const unsigned char* __restrict__ code, int64_t first, int64_t ncell) {
  if (first + CAP_PER_THREAD <= ncell) return *reinterpret_cast<const uint4*>(code + first);       // the code array is 256-byte aligned
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  for (int q = 0; q < CAP_PER_THREAD; ++q)
    if (first + q < ncell) w[q >> 2] |= (uint32_t)code[first + q] << (8 * (q & 3));
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ unsigned cap_code_of(const uint4& c, int q) {
  const uint32_t w = (q < 4) ? c.x : (q < 8) ? c.y : (q < 12) ? c.z : c.w;
  return (w >> (8 * (q & 3))) & 0xffu;
}
__global__ void __launch_bounds__(CAP_BLOCK) cap_count_kernel(const unsigned char* __restrict__ code, int64_t ncell, long long* __restrict__ block_count) {
  const int64_t first = ((int64_t)blockIdx.x * CAP_BLOCK + threadIdx.x) * CAP_PER_THREAD;
  const uint4 c = cap_load_codes(code, first, ncell);
  int n = 0;
  if (c.x | c.y | c.z | c.w) {
#pragma unroll
    for (int q = 0; q < CAP_PER_THREAD; ++q) n += MC_NTRI[cap_code_of(c, q)];
  }
  typedef cub::BlockReduce<int, CAP_BLOCK> Reduce;
  __shared__ Reduce::TempStorage tmp;
  const int sum = Reduce(tmp).Sum(n);
  if (threadIdx.x == 0) block_count[blockIdx.x] = sum;
}

__global__ void __launch_bounds__(CAP_BLOCK) cap_emit_kernel(const float* __restrict__ df, const unsigned char* __restrict__ code, int N, int64_t ncell,
                                                             const long long* __restrict__ block_offset, double* __restrict__ tris) {
  const int M = N - 1;
  const int64_t first = ((int64_t)blockIdx.x * CAP_BLOCK + threadIdx.x) * CAP_PER_THREAD;
  const uint4 codes = cap_load_codes(code, first, ncell);
  int n = 0;
  if (codes.x | codes.y | codes.z | codes.w) {
#pragma unroll
    for (int q = 0; q < CAP_PER_THREAD; ++q) n += MC_NTRI[cap_code_of(codes, q)];
  }
  typedef cub::BlockScan<int, CAP_BLOCK> Scan;
  __shared__ Scan::TempStorage tmp;
  int off;
  Scan(tmp).ExclusiveSum(n, off);
  if (n == 0) return;
  const double den = (double)(N - 1);             // v / (N-1) * 2 + (-1), in the order the reference evaluates it
  double* out = tris + (block_offset[blockIdx.x] + off) * 9;
  for (int q = 0; q < CAP_PER_THREAD; ++q) {
    const unsigned cc = cap_code_of(codes, q);
    const int nt = MC_NTRI[cc];
    if (nt == 0) continue;
    int i, j, k;
    cap_cell_ijk(first + q, M, i, j, k);
    const int64_t base = ((int64_t)i * N + j) * N + k;
    double val[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double a = (double)df[base + ((c >> 2) & 1) * (int64_t)N * N + ((c >> 1) & 1) * N + (c & 1)];
      val[c] = ((cc >> c) & 1u) ? -a : a;
    }
    for (int t = 0; t < nt; ++t) {
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int e = MC_TRIS[cc][t * 3 + v];
        const int c0 = MC_EDGE_CORNERS[e][0], c1 = MC_EDGE_CORNERS[e][1];
        const double mu = (0.0 - val[c0]) / (val[c1] - val[c0]);
        const double px = (double)((c0 >> 2) & 1) + mu * (double)(((c1 >> 2) & 1) - ((c0 >> 2) & 1));
        const double py = (double)((c0 >> 1) & 1) + mu * (double)(((c1 >> 1) & 1) - ((c0 >> 1) & 1));
        const double pz = (double)(c0 & 1) + mu * (double)((c1 & 1) - (c0 & 1));
        out[t * 9 + v * 3 + 0] = (px + (double)i) / den * 2.0 + (-1.0);
        out[t * 9 + v * 3 + 1] = (py + (double)j) / den * 2.0 + (-1.0);
        out[t * 9 + v * 3 + 2] = (pz + (double)k) / den * 2.0 + (-1.0);
      }
    }
    out += nt * 9;
  }
}

int cap_units(int64_t ncell) { return (int)((ncell + CAP_UNIT - 1) / CAP_UNIT); }

size_t cap_scan_temp_bytes(int64_t nblocks) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const long long*)nullptr, (long long*)nullptr, (int)nblocks);
  return b + 256;
}

int cap_classify(const float* df, const float* vecs, int N, float thr, unsigned char* code, long long* block_count, long long* block_offset,
                 void* temp, size_t temp_bytes, cudaStream_t st) {
  const int64_t M = N - 1, ncell = M * M * M;
  const int64_t nblocks = cap_units(ncell);
  const int64_t tiles = ((M + CAP_TI - 1) / CAP_TI) * ((M + CAP_TJ - 1) / CAP_TJ) * ((M + CAP_TK - 1) / CAP_TK);
  cap_classify_kernel<<<(unsigned)tiles, CAP_BLOCK, 0, st>>>(df, vecs, N, thr, code);
  DUDF_LAUNCH_OK();
  cap_count_kernel<<<(unsigned)nblocks, CAP_BLOCK, 0, st>>>(code, ncell, block_count);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, block_count, block_offset, (int)nblocks, st));
  dudf_count_launch();
  return 0;
}

int cap_emit(const float* df, const unsigned char* code, int N, const long long* block_offset, double* tris, cudaStream_t st) {
  const int64_t M = N - 1, ncell = M * M * M;
  const int64_t nblocks = cap_units(ncell);
  cap_emit_kernel<<<(unsigned)nblocks, CAP_BLOCK, 0, st>>>(df, code, N, ncell, block_offset, tris);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

// CAP-UDF marching cubes on the device: extract_mesh_CAP of src/render_mc.py:201-256, the step right after the dense grid
// query.  The reference walks the (N-1)^3 cells in a Python triple loop (hours at 512^3); cells are independent:
//   * a cell is skipped when the smallest of its 8 distances exceeds the threshold (0.008);
//   * corner c is NEGATIVE when dot(grad[corner 0], grad[corner c]) < 0, its value is -ndf, otherwise +ndf;
//   * if any value is negative the signed 2x2x2 block goes through marching cubes at iso-value 0: vertices at the linear zero
//     crossing of the sign-changing edges (float64, like the float64 block the reference builds), shifted by the cell index and
//     mapped to [-1,1]^3 by v / (N-1) * 2 - 1.
// Two passes, both HBM-bound (one thread per cell, k fastest -> coalesced reads, neighbours share lines through L1/L2):
//   classify  reads the 8 distances (and the gradients of near-surface cells only), writes one case byte per cell and one
//             triangle count per 256-cell block;  an exclusive scan of the block counts gives every block its output offset;
//   emit      re-reads the case bytes, scans the per-cell counts inside the block and writes the triangles of cell (i,j,k)
//             in the reference's order (cells in i, j, k lexicographic order, triangles in table order).
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

// case index of one cell (0: nothing to emit)
__device__ __forceinline__ unsigned cap_cell_case(const float* __restrict__ df, const float* __restrict__ vecs, int N, float thr, int i, int j, int k) {
  const int64_t base = ((int64_t)i * N + j) * N + k;
  float v[8];
  float vmin = 3.4e38f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    v[c] = df[base + ((c >> 2) & 1) * (int64_t)N * N + ((c >> 1) & 1) * N + (c & 1)];
    vmin = fminf(vmin, v[c]);
  }
  if (vmin > thr) return 0u;
  const float gx = vecs[base * 3], gy = vecs[base * 3 + 1], gz = vecs[base * 3 + 2];
  unsigned code = 0;
#pragma unroll
  for (int c = 1; c < 8; ++c) {
    const int64_t p = base + ((c >> 2) & 1) * (int64_t)N * N + ((c >> 1) & 1) * N + (c & 1);
    const float d = __fadd_rn(__fadd_rn(__fmul_rn(gx, vecs[p * 3]), __fmul_rn(gy, vecs[p * 3 + 1])), __fmul_rn(gz, vecs[p * 3 + 2]));
    if (d < 0.f && v[c] > 0.f) code |= 1u << c;       // res = -val is negative only for val > 0 (res.min() < 0)
  }
  return code;
}

__global__ void __launch_bounds__(CAP_BLOCK) cap_classify_kernel(const float* __restrict__ df, const float* __restrict__ vecs, int N, float thr,
                                                                 int64_t ncell, unsigned char* __restrict__ code, long long* __restrict__ block_count) {
  const int M = N - 1;
  const int64_t cell = (int64_t)blockIdx.x * CAP_BLOCK + threadIdx.x;
  unsigned cc = 0;
  if (cell < ncell) {
    int i, j, k;
    cap_cell_ijk(cell, M, i, j, k);
    cc = cap_cell_case(df, vecs, N, thr, i, j, k);
    code[cell] = (unsigned char)cc;
  }
  const int n = MC_NTRI[cc];
  typedef cub::BlockReduce<int, CAP_BLOCK> Reduce;
  __shared__ Reduce::TempStorage tmp;
  const int sum = Reduce(tmp).Sum(n);
  if (threadIdx.x == 0) block_count[blockIdx.x] = sum;
}

__global__ void __launch_bounds__(CAP_BLOCK) cap_emit_kernel(const float* __restrict__ df, const unsigned char* __restrict__ code, int N, int64_t ncell,
                                                             const long long* __restrict__ block_offset, double* __restrict__ tris) {
  const int M = N - 1;
  const int64_t cell = (int64_t)blockIdx.x * CAP_BLOCK + threadIdx.x;
  const unsigned cc = (cell < ncell) ? code[cell] : 0u;
  const int n = MC_NTRI[cc];
  typedef cub::BlockScan<int, CAP_BLOCK> Scan;
  __shared__ Scan::TempStorage tmp;
  int off;
  Scan(tmp).ExclusiveSum(n, off);
  if (n == 0) return;
  int i, j, k;
  cap_cell_ijk(cell, M, i, j, k);
  const int64_t base = ((int64_t)i * N + j) * N + k;
  double val[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double a = (double)df[base + ((c >> 2) & 1) * (int64_t)N * N + ((c >> 1) & 1) * N + (c & 1)];
    val[c] = ((cc >> c) & 1u) ? -a : a;
  }
  const double den = (double)(N - 1);             // v / (N-1) * 2 + (-1), in the order the reference evaluates it
  double* out = tris + (block_offset[blockIdx.x] + off) * 9;
  for (int t = 0; t < n; ++t) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int e = MC_TRIS[cc][t * 3 + q];
      const int c0 = MC_EDGE_CORNERS[e][0], c1 = MC_EDGE_CORNERS[e][1];
      const double mu = (0.0 - val[c0]) / (val[c1] - val[c0]);
      const double px = (double)((c0 >> 2) & 1) + mu * (double)(((c1 >> 2) & 1) - ((c0 >> 2) & 1));
      const double py = (double)((c0 >> 1) & 1) + mu * (double)(((c1 >> 1) & 1) - ((c0 >> 1) & 1));
      const double pz = (double)(c0 & 1) + mu * (double)((c1 & 1) - (c0 & 1));
      out[t * 9 + q * 3 + 0] = (px + (double)i) / den * 2.0 + (-1.0);
      out[t * 9 + q * 3 + 1] = (py + (double)j) / den * 2.0 + (-1.0);
      out[t * 9 + q * 3 + 2] = (pz + (double)k) / den * 2.0 + (-1.0);
    }
  }
}

size_t cap_scan_temp_bytes(int64_t nblocks) {
  size_t b = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const long long*)nullptr, (long long*)nullptr, (int)nblocks);
  return b + 256;
}

int cap_classify(const float* df, const float* vecs, int N, float thr, unsigned char* code, long long* block_count, long long* block_offset,
                 void* temp, size_t temp_bytes, cudaStream_t st) {
  const int64_t M = N - 1, ncell = M * M * M;
  const int64_t nblocks = (ncell + CAP_BLOCK - 1) / CAP_BLOCK;
  cap_classify_kernel<<<(unsigned)nblocks, CAP_BLOCK, 0, st>>>(df, vecs, N, thr, ncell, code, block_count);
  DUDF_LAUNCH_OK();
  DUDF_CUDA_OK(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, block_count, block_offset, (int)nblocks, st));
  dudf_count_launch();
  return 0;
}

int cap_emit(const float* df, const unsigned char* code, int N, const long long* block_offset, double* tris, cudaStream_t st) {
  const int64_t M = N - 1, ncell = M * M * M;
  const int64_t nblocks = (ncell + CAP_BLOCK - 1) / CAP_BLOCK;
  cap_emit_kernel<<<(unsigned)nblocks, CAP_BLOCK, 0, st>>>(df, code, N, ncell, block_offset, tris);
  DUDF_LAUNCH_OK();
  return 0;
}

}  // namespace dudf

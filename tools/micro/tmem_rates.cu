// TMEM read rate of the epilogue's access pattern (tcgen05.ld.32x32b, each warp its own 32 lanes) at 4 / 8 / 16 warps per CTA
// and x16 / x32 column loads: bytes per clock per SM, with the tensor pipe idle.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../diffudf_b200/csrc -o tmem_rates tmem_rates.cu && ./tmem_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "dudf_umma.cuh"
using namespace umma;

template <int X>
__global__ void k(float* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 512; c += 2 * X) {          // two loads in flight
      uint32_t r[32], q[32];
      if (X == 32) { tmem_ld_x32(base + c, r); tmem_ld_x32(base + c + 32, q); }
      else { tmem_ld_x16(base + c, r); tmem_ld_x16(base + c + 16, q); }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < X; ++j) acc ^= r[j] + q[j];
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(slot);
}

template <int X>
void run() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 200;
  for (int warps : {4, 8, 16}) {
    k<X><<<148, warps * 32>>>(out, clk, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return; }
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double bytes = (double)iters * 512 * 128 * warps;      // each warp reads 32 lanes x 4 B x 512 columns per iteration
    printf("x%-2d loads, %2d warps: %7.1f B/clk/SM  (%.0f clk per 128 KB sub-tile accumulator)\n", X, warps, bytes / avg, 131072.0 / (bytes / avg));
  }
  cudaFree(out); cudaFree(clk);
}
int main() { run<32>(); run<16>(); return 0; }

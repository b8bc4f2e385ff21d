// Issue-rate micro-benchmark of the instructions the sine-jet epilogues are made of (sm_100a): warp-instructions per clock per
// SM sub-partition for MUFU.SIN/COS, F2FP pack, half->float unpack, FFMA, FMNMX, at 1/2/4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
template <int OP>
__global__ void k(float* out, long long* clk, float seed) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.01f + threadIdx.x * 1e-3f; h[i] = 0x3c003c00u + i; }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(seed));
      if (OP == 1) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                 // FMUL + MUFU.SIN
      if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));                 // MUFU.EX2 alone
      if (OP == 3) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h[i]) : "f"(a[i])); a[i] = __uint_as_float(h[i]); }     // F2FP.PACK_AB (chained)
      if (OP == 4) { asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(a[i]) : "r"(h[i])); h[i] = __float_as_uint(a[i]); }   // half -> float (chained)
      if (OP == 5) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
      if (OP == 6) { asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %1;" : "=r"(h[i]) : "f"(a[i])); a[i] = __uint_as_float(h[i]); }
      if (OP == 7) asm volatile("fma.rn.f16x2 %0, %0, %1, %0;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
      if (OP == 8) asm volatile("cvt.rni.f32.f32 %0, %0;" : "+f"(a[i]));                     // FRND
      if (OP == 10) { asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(a[i]) : "r"(h[i])); h[i] = __float_as_uint(a[i]); }   // high half -> float
      if (OP == 11) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h[i]) : "f"(a[i])); asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(__uint_as_float(h[i]))); }   // F2FP + MUFU: same pipe?
      if (OP == 9) { asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[(i + 1) & 7])); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[(i + 2) & 7]) : "f"(seed));
                     asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[(i + 3) & 7]) : "f"(seed)); }    // MUFU interleaved with 2 FFMA
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  printf("%-44s", name);
  for (int warps : {4, 8, 16}) {
    k<OP><<<148, warps * 32>>>(out, clk, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double instr_per_smsp = (double)ITERS * 8 * per_iter * (warps / 4);
    printf("  %dw/sched: %6.2f clk/warp-instr", warps / 4, avg / instr_per_smsp);
  }
  printf("\n");
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("sin.approx (FMUL + MUFU.SIN), per pair", 1);
  run<2>("ex2.approx (MUFU.EX2)", 1);
  run<3>("cvt.rn.f16x2.f32 (F2FP.PACK_AB)", 1);
  run<6>("cvt.rn.satfinite.f16x2.f32", 1);
  run<4>("cvt.f32.f16 (half -> float)", 1);
  run<5>("min.f32 (FMNMX)", 1);
  run<7>("fma.rn.f16x2 (HFMA2)", 1);
  run<8>("cvt.rni.f32.f32 (FRND)", 1);
  run<9>("sin.approx + 2 FFMA, per triple", 1);
  run<10>("cvt.f32.f16 high half", 1);
  run<11>("F2FP + sin.approx, per pair", 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}

// Micro-benchmark: issue throughput of packed fp32 (FFMA2) vs scalar FFMA on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ffma2_bench tools/micro/ffma2_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float x) {
  float a[16];
  unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
  const unsigned long long xx = ((unsigned long long)__float_as_uint(x) << 32) | __float_as_uint(x);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(x));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], xx, xx);
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i];
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)148 * 8 * 256 * iters * 16;
      if (rep) printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s), fma/clk/SM @1.9GHz = %.1f\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9,
                      2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}

// Ceiling of the cost volume's OUTPUT pattern alone (no taps, no math): the same grid decomposition as cost_volume_kernel
// (CTA = 4 warps = 128 consecutive pixels of one ERP row x a chunk of depths) writing D*C*H*W floats either
//   mode 0: channels-last  (B,D,H,W,C): a warp writes 32 pixels x 128 B = 4 KB contiguous per depth (128-bit streaming stores)
//   mode 1: planar         (B,D,C,H,W): a warp writes 32 channel rows of 128 B, 512 KB apart, per depth (32-bit streaming stores)
//   mode 2: planar, 128-bit stores (lane = 4 pixels of one of 8 channels per instruction)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o planar_write planar_write.cu && ./planar_write
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128, 5) write_kernel(float* out, int H, int W, int D, int d_chunk) {
  const int C = 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = W / 128;
  const int y = blockIdx.x / tiles_x;
  const int x_warp = (blockIdx.x % tiles_x) * 128 + warp * 32;
  const int d0 = blockIdx.y * d_chunk, d1 = min(D, d0 + d_chunk);
  const size_t HW = (size_t)H * W;
  for (int d = d0; d < d1; ++d) {
    const float v = (float)d + (float)lane;
    if (MODE == 0) {
      float4* o = reinterpret_cast<float4*>(out) + (((size_t)d * H + y) * W + x_warp) * 8 + lane;
#pragma unroll
      for (int j = 0; j < 8; ++j) __stcs(o + j * 32, make_float4(v, v, v, v));
    } else if (MODE == 1) {
      float* o = out + (size_t)d * C * HW + (size_t)y * W + x_warp + lane;
#pragma unroll 8
      for (int c = 0; c < C; ++c) __stcs(o + (size_t)c * HW, v);
    } else {
      float4* o = reinterpret_cast<float4*>(out + (size_t)d * C * HW + (size_t)(lane >> 3) * HW + (size_t)y * W + x_warp) + (lane & 7);
#pragma unroll
      for (int c = 0; c < C; c += 4) __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(o) + (size_t)c * HW), make_float4(v, v, v, v));
    }
  }
}

int main() {
  const int H = 256, W = 512, D = 64, C = 32;
  const size_t n = (size_t)D * C * H * W;
  float *out, *flush;
  cudaMalloc(&out, n * 4);
  cudaMalloc(&flush, 256u << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int d_chunk : {4, 8, 16, 64}) {
    dim3 grid((W / 128) * H, (D + d_chunk - 1) / d_chunk);
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      for (int it = 0; it < 6; ++it) {
        cudaMemset(flush, 0, 256u << 20);
        cudaEventRecord(e0);
        if (mode == 0) write_kernel<0><<<grid, 128>>>(out, H, W, D, d_chunk);
        else if (mode == 1) write_kernel<1><<<grid, 128>>>(out, H, W, D, d_chunk);
        else write_kernel<2><<<grid, 128>>>(out, H, W, D, d_chunk);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it > 0 && ms < best) best = ms;
      }
      printf("d_chunk %2d mode %d: %.4f ms  %.1f GB/s\n", d_chunk, mode, best, n * 4 / best / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

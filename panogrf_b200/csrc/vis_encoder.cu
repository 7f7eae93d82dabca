// Per-call CNN on the render path: DefaultVisEncoder (network/vis_encoder.py:6-33) =
//   cat(resize(img_feats), ray_feats) -> conv3x3 -> 2 x ResidualBlock(InstanceNorm, ReLU, conv3x3, InstanceNorm, ReLU, conv3x3; + skip)
//   -> conv1x1   (network/ops.py:6-29, 61-115).
// The convolutions run on the tensor cores through csrc/conv3d.cu (a (N,1,h,w,C) volume is a 2-D feature map; residual add in the
// epilogue; wrap or zero padding along the width).  This file holds the rest: the fused resize + concatenation + bf16 channels-last
// conversion of the two NCHW inputs, and InstanceNorm2d(affine) + ReLU as a statistics pass (fp64 accumulation) and an apply pass.
// Numerics: bf16 activations between layers, fp32 / fp64 arithmetic inside each kernel (bf16 render mode, rtol 1e-2).
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace pgrf {

// thread = (pixel, 8-channel chunk).  img (N,Ci,hi,wi) fp32 NCHW is resized to (h,w) like F.interpolate(size, 'bilinear',
// align_corners=False) (ATen area_pixel_compute_source_index) when the sizes differ; ray (N,Cr,h,w) fp32 NCHW follows it.
__global__ void __launch_bounds__(256) feats_to_bf16_cl_kernel(const float* __restrict__ img, int Ci, int hi, int wi,
                                                               const float* __restrict__ ray, int Cr, int N, int h, int w,
                                                               __nv_bfloat16* __restrict__ out) {
  const int C = Ci + Cr, C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)N * h * w * C8) return;
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int x = (int)(t % w); t /= w;
  const int y = (int)(t % h);
  const int n = (int)(t / h);
  float v[8];
  if (8 * c8 < Ci) {
    const float sy = fmaxf(((float)y + 0.5f) * ((float)hi / (float)h) - 0.5f, 0.f), sx = fmaxf(((float)x + 0.5f) * ((float)wi / (float)w) - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const bool same = hi == h && wi == w;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* pl = img + ((size_t)n * Ci + 8 * c8 + k) * hi * wi;
      if (same) v[k] = __ldg(pl + (size_t)y * wi + x);
      else
        v[k] = (1.f - ly) * ((1.f - lx) * __ldg(pl + (size_t)y0 * wi + x0) + lx * __ldg(pl + (size_t)y0 * wi + x1)) +
               ly * ((1.f - lx) * __ldg(pl + (size_t)y1 * wi + x0) + lx * __ldg(pl + (size_t)y1 * wi + x1));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(ray + (((size_t)n * Cr + (8 * c8 - Ci) + k) * h + y) * w + x);
  }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(out + (size_t)(i / C8) * C)[c8] = o;
}

// InstanceNorm2d statistics of a bf16 channels-last map (N, HW, C): stats[n][c] = (sum, sum of squares) in fp64.
// block = 256 threads = (32 pixel lanes) x (C/8 <= 8 channel chunks); grid.x strides over the pixels of image blockIdx.y
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, double* __restrict__ stats) {
  const int C8 = C / 8, lanes = 256 / C8;
  const int c8 = threadIdx.x % C8, pl = threadIdx.x / C8;
  const int n = blockIdx.y;
  float s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    for (int p = blockIdx.x * lanes + pl; p < HW; p += gridDim.x * lanes) {
      const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)n * HW + p) * C) + c8);
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(hh[j]);
        s1[2 * j] += f.x; s1[2 * j + 1] += f.y;
        s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]); s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
      }
    }
  }
  __shared__ float sm[2][64][33];     // [sum | sumsq][channel][pixel lane]
  if (pl < lanes && pl < 32)
    for (int k = 0; k < 8; ++k) { sm[0][8 * c8 + k][pl] = s1[k]; sm[1][8 * c8 + k][pl] = s2[k]; }
  // lanes > 32 only when C8 < 8: fold the upper pixel lanes in first
  __syncthreads();
  if (pl >= 32 && pl < lanes)
    for (int k = 0; k < 8; ++k) { atomicAdd(&sm[0][8 * c8 + k][pl & 31], s1[k]); atomicAdd(&sm[1][8 * c8 + k][pl & 31], s2[k]); }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    const int which = threadIdx.x / C, c = threadIdx.x % C;
    double acc = 0.0;
    const int used = lanes < 32 ? lanes : 32;
    for (int l = 0; l < used; ++l) acc += (double)sm[which][c][l];
    atomicAdd(stats + ((size_t)n * C + c) * 2 + which, acc);
  }
}

// y = relu((x - mean) * rstd * gamma + beta), biased variance, eps inside the square root (nn.InstanceNorm2d); thread = (pixel, 8 channels)
__global__ void __launch_bounds__(256) instnorm_relu_kernel(const __nv_bfloat16* __restrict__ x, int N, int HW, int C,
                                                            const double* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y) {
  const int C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)N * HW * C8) return;
  const int c8 = (int)(i % C8);
  const int n = (int)(i / ((unsigned)HW * C8));
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (size_t)(i / C8) * C) + c8);
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
  float v[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = 8 * c8 + k;
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean = s1 / HW, var = fmax(s2 / HW - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float o = (v[k] - (float)mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    v[k] = o > 0.f ? o : 0.f;
  }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(y + (size_t)(i / C8) * C)[c8] = o;
}

}  // namespace pgrf

using namespace pgrf;

static inline unsigned vblocks(long long n) { return (unsigned)((n + 255) / 256); }

extern "C" int pgrf_feats_to_bf16_cl(const float* img, int Ci, int hi, int wi, const float* ray, int Cr, int N, int h, int w, void* out,
                                     void* stream) {
  PGRF_REQUIRE(img && ray && out && N >= 1 && h >= 1 && w >= 1 && hi >= 1 && wi >= 1, "feats_to_bf16_cl: bad arguments");
  PGRF_REQUIRE(Ci % 8 == 0 && Cr % 8 == 0 && (Ci + Cr) % 16 == 0, "feats_to_bf16_cl: Ci=%d Cr=%d (multiples of 8, sum multiple of 16)", Ci, Cr);
  const long long n = (long long)N * h * w * ((Ci + Cr) / 8);
  PGRF_REQUIRE(n < 4294967040LL, "feats_to_bf16_cl: %lld work items exceed the 32-bit index range", n);
  feats_to_bf16_cl_kernel<<<vblocks(n), 256, 0, (cudaStream_t)stream>>>(img, Ci, hi, wi, ray, Cr, N, h, w, (__nv_bfloat16*)out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_instnorm_relu_fwd(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                                      void* y, void* stream) {
  PGRF_REQUIRE(x && gamma && beta && stats_ws && y && N >= 1 && HW >= 1, "instnorm_relu: bad arguments");
  PGRF_REQUIRE(C % 8 == 0 && C >= 8 && C <= 64, "instnorm_relu: C=%d (multiple of 8, at most 64)", C);
  const long long n = (long long)N * HW * (C / 8);
  PGRF_REQUIRE(n < 4294967040LL, "instnorm_relu: %lld work items exceed the 32-bit index range", n);
  cudaStream_t st = (cudaStream_t)stream;
  PGRF_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * C, st));
  const int lanes = 256 / (C / 8);
  int gx = (HW + lanes * 8 - 1) / (lanes * 8);
  gx = gx < 1 ? 1 : (gx > 592 ? 592 : gx);
  instnorm_stats_kernel<<<dim3(gx, N), 256, 0, st>>>((const __nv_bfloat16*)x, HW, C, stats_ws);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  instnorm_relu_kernel<<<vblocks(n), 256, 0, st>>>((const __nv_bfloat16*)x, N, HW, C, stats_ws, gamma, beta, eps, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

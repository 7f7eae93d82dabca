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
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, int cs,
                                                             double* __restrict__ stats) {
  // x points at the first channel of a slab of `cs` <= 64 channels inside rows of C channels
  const int C8 = cs / 8, lanes = 256 / C8;
  const int c8 = threadIdx.x % C8, pl = threadIdx.x / C8;
  const int n = blockIdx.y;
  float s1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, s2[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < lanes) {
    // four independent 16-byte loads in flight per thread (the kernel is a latency chain otherwise); same summation order as a plain loop
    const int step = gridDim.x * lanes;
    for (int p0 = blockIdx.x * lanes + pl; p0 < HW; p0 += 4 * step) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * step;
        q[u] = p < HW ? __ldg(reinterpret_cast<const uint4*>(x + ((size_t)n * HW + p) * C) + c8) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(hh[j]);
          s1[2 * j] += f.x; s1[2 * j + 1] += f.y;
          s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]); s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
        }
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
  if (threadIdx.x < 2 * cs) {
    const int which = threadIdx.x / cs, c = threadIdx.x % cs;
    double acc = 0.0;
    const int used = lanes < 32 ? lanes : 32;
    for (int l = 0; l < used; ++l) acc += (double)sm[which][c][l];
    atomicAdd(stats + ((size_t)n * C + c) * 2 + which, acc);
  }
}

// y = act((x - mean) * rstd * gamma + beta [+ res]), biased variance, eps inside the square root (nn.InstanceNorm2d);
// act: 0 none, 1 ReLU, 2 ELU; thread = (pixel, 8 channels)
__global__ void __launch_bounds__(256) instnorm_act_kernel(const __nv_bfloat16* __restrict__ x, int N, int HW, int C,
                                                           const double* __restrict__ stats, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ res,
                                                           int act, __nv_bfloat16* __restrict__ y) {
  // grid = (blocks over HW * C/8, N): a block stays inside one image, so mean / rstd / gamma / beta of its channels are computed once
  // per block (one fp64 division + square root per channel instead of one per element) and read from shared memory
  __shared__ float4 s_par[256];
  const int C8 = C / 8;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean = s1 / HW, var = fmax(s2 / HW - mean * mean, 0.0);
    s_par[c] = make_float4((float)mean, (float)(1.0 / sqrt(var + (double)eps)), __ldg(gamma + c), __ldg(beta + c));
  }
  __syncthreads();
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;          // (pixel, channel chunk) inside the image
  if (j >= (unsigned)HW * C8) return;
  const int c8 = (int)(j % C8);
  const size_t row = (size_t)n * HW + j / C8;
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + row * C) + c8);
  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
  float v[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(hh[k]); v[2 * k] = f.x; v[2 * k + 1] = f.y; }
  if (res) {
    const uint4 qr = __ldg(reinterpret_cast<const uint4*>(res + row * C) + c8);
    const __nv_bfloat162* hr = reinterpret_cast<const __nv_bfloat162*>(&qr);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(hr[k]); rv[2 * k] = f.x; rv[2 * k + 1] = f.y; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 pr = s_par[8 * c8 + k];
    float o = (v[k] - pr.x) * pr.y * pr.z + pr.w + rv[k];
    if (act == 1) o = o > 0.f ? o : 0.f;
    else if (act == 2) o = o > 0.f ? o : expm1f(o);
    v[k] = o;
  }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(y + row * C)[c8] = o;
}

// ---- ResUNetLight parts (network/ops.py:235-455) --------------------------------------------------------------------------------
// conv1 = WrapPadding(3) + Conv2d(k = 7, stride 2): the 7x7xCin patch of every OUTPUT pixel as one bf16 channels-last row
// (k = c * 49 + ky * 7 + kx, the flattening of the PyTorch weight, zero padded to Kpad), so the convolution is a pointwise GEMM.
// x: fp32 NCHW (N,Cin,H,W); wrap: zeros along height + wrap along width, else zeros on every side; thread = (output pixel, 8 k's)
__global__ void __launch_bounds__(256) patch7x7_s2_kernel(const float* __restrict__ x, int N, int Cin, int H, int W, int ho, int wo, int Kpad,
                                                          int wrap, __nv_bfloat16* __restrict__ out) {
  const int K8 = Kpad / 8, K = Cin * 49;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)N * ho * wo * K8) return;
  unsigned t = i;
  const int k8 = (int)(t % K8); t /= K8;
  const int xo = (int)(t % wo); t /= wo;
  const int yo = (int)(t % ho);
  const int n = (int)(t / ho);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = 8 * k8 + j;
    float val = 0.f;
    if (k < K) {
      const int c = k / 49, r = k - c * 49, ky = r / 7, kx = r - ky * 7;
      const int yy = 2 * yo + ky - 3;
      int xx = 2 * xo + kx - 3;
      bool ok = yy >= 0 && yy < H;
      if (wrap) xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
      else ok = ok && xx >= 0 && xx < W;
      if (ok) val = __ldg(x + (((size_t)n * Cin + c) * H + yy) * W + xx);
    }
    v[j] = val;
  }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(out + (size_t)(i / K8) * Kpad)[k8] = o;
}

// stride 2 of a stride-1 convolution output / of the input of a 1x1 stride-2 convolution: y[yo][xo] = x[2 yo][2 xo]
__global__ void __launch_bounds__(256) subsample2_kernel(const __nv_bfloat16* __restrict__ x, int N, int h, int w, int C,
                                                         __nv_bfloat16* __restrict__ y) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2, C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)N * ho * wo * C8) return;
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % wo); t /= wo;
  const int yo = (int)(t % ho);
  const int n = (int)(t / ho);
  reinterpret_cast<uint4*>(y + (size_t)(i / C8) * C)[c8] =
      __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * h + 2 * yo) * w + 2 * xo) * C) + c8);
}

// F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) on bf16 channels-last (upconv, ops.py:226-233)
__global__ void __launch_bounds__(256) upsample2d2_ac_kernel(const __nv_bfloat16* __restrict__ x, int N, int h, int w, int C,
                                                             __nv_bfloat16* __restrict__ y) {
  const int ho = 2 * h, wo = 2 * w, C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)N * ho * wo * C8) return;
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % wo); t /= wo;
  const int yo = (int)(t % ho);
  const int n = (int)(t / ho);
  const float sy = ho > 1 ? (float)yo * ((float)(h - 1) / (float)(ho - 1)) : 0.f;
  const float sx = wo > 1 ? (float)xo * ((float)(w - 1) / (float)(wo - 1)) : 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
  const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * h + ys[k]) * w + xs[k]) * C) + c8);
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); acc[2 * j] = fmaf(wgt[k], f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(wgt[k], f.y, acc[2 * j + 1]); }
  }
  uint4 o;
  o.x = umma::pack2(acc[0], acc[1]); o.y = umma::pack2(acc[2], acc[3]); o.z = umma::pack2(acc[4], acc[5]); o.w = umma::pack2(acc[6], acc[7]);
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

static int instnorm_launch(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                           const void* res, int act, void* y, void* stream) {
  PGRF_REQUIRE(x && gamma && beta && stats_ws && y && N >= 1 && HW >= 1, "instnorm: bad arguments");
  PGRF_REQUIRE(C % 8 == 0 && C >= 8 && C <= 64 * 4, "instnorm: C=%d (multiple of 8, at most 256)", C);
  PGRF_REQUIRE(act >= 0 && act <= 2, "instnorm: act=%d (0 none, 1 relu, 2 elu)", act);
  const long long n = (long long)N * HW * (C / 8);
  PGRF_REQUIRE(n < 4294967040LL, "instnorm: %lld work items exceed the 32-bit index range", n);
  cudaStream_t st = (cudaStream_t)stream;
  PGRF_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * C, st));
  // the statistics kernel reduces at most 64 channels per launch: wider maps are handled in channel slabs through the pointer offset
  for (int c0 = 0; c0 < C; c0 += 64) {
    const int cs = C - c0 < 64 ? C - c0 : 64;
    const int lanes = 256 / (cs / 8);
    int gx = (HW + lanes * 8 - 1) / (lanes * 8);
    gx = gx < 1 ? 1 : (gx > 592 ? 592 : gx);
    instnorm_stats_kernel<<<dim3(gx, N), 256, 0, st>>>((const __nv_bfloat16*)x + c0, HW, C, cs, stats_ws + 2 * c0);
    count_launch();
    PGRF_CUDA(cudaGetLastError());
  }
  const long long per_image = (long long)HW * (C / 8);
  PGRF_REQUIRE(N <= 65535, "instnorm: N=%d exceeds the grid's y range", N);
  instnorm_act_kernel<<<dim3((unsigned)((per_image + 255) / 256), (unsigned)N), 256, 0, st>>>(
      (const __nv_bfloat16*)x, N, HW, C, stats_ws, gamma, beta, eps, (const __nv_bfloat16*)res, act, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_instnorm_relu_fwd(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                                      void* y, void* stream) {
  return instnorm_launch(x, N, HW, C, gamma, beta, eps, stats_ws, nullptr, 1, y, stream);
}

extern "C" int pgrf_instnorm_act_fwd(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                                     const void* res, int act, void* y, void* stream) {
  return instnorm_launch(x, N, HW, C, gamma, beta, eps, stats_ws, res, act, y, stream);
}

extern "C" int pgrf_patch7x7_s2_fwd(const float* x, int N, int Cin, int H, int W, int Kpad, int wrap, void* out, void* stream) {
  PGRF_REQUIRE(x && out && N >= 1 && Cin >= 1 && H >= 1 && W >= 1, "patch7x7_s2: bad arguments");
  PGRF_REQUIRE(Kpad % 16 == 0 && Kpad >= Cin * 49, "patch7x7_s2: Kpad=%d must be a multiple of 16 and >= %d", Kpad, Cin * 49);
  const int ho = (H - 1) / 2 + 1, wo = (W - 1) / 2 + 1;
  const long long n = (long long)N * ho * wo * (Kpad / 8);
  PGRF_REQUIRE(n < 4294967040LL, "patch7x7_s2: %lld work items exceed the 32-bit index range", n);
  patch7x7_s2_kernel<<<vblocks(n), 256, 0, (cudaStream_t)stream>>>(x, N, Cin, H, W, ho, wo, Kpad, wrap, (__nv_bfloat16*)out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_subsample2_fwd(const void* x, int N, int h, int w, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && N >= 1 && h >= 1 && w >= 1 && C % 8 == 0, "subsample2: bad arguments");
  const long long n = (long long)N * ((h + 1) / 2) * ((w + 1) / 2) * (C / 8);
  PGRF_REQUIRE(n < 4294967040LL, "subsample2: %lld work items exceed the 32-bit index range", n);
  subsample2_kernel<<<vblocks(n), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, N, h, w, C, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_upsample2d2_ac_fwd(const void* x, int N, int h, int w, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && N >= 1 && h >= 1 && w >= 1 && C % 8 == 0, "upsample2d2_ac: bad arguments");
  const long long n = (long long)N * (2 * h) * (2 * w) * (C / 8);
  PGRF_REQUIRE(n < 4294967040LL, "upsample2d2_ac: %lld work items exceed the 32-bit index range", n);
  upsample2d2_ac_kernel<<<vblocks(n), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, N, h, w, C, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

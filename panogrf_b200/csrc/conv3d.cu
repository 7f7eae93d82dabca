// 3-D cost regulariser: the Conv3DBlockv2 / UNet2 stack of models/common_blocks.py:187-242, 366-503 as consumed by
// network/omni_mvsnet/pipeline3_model.py:847-855, on the 5th-generation tensor cores.
//
//   conv3d_igemm_kernel  3x3x3 convolution (+ bias + LeakyReLU 0.01) as an implicit GEMM on tcgen05:
//        M = 128 output voxels per CTA (thread <-> voxel <-> TMEM lane), N = a tile of output channels (16..128, fp32 accumulators in
//        tensor memory), K = 27 taps x input channels, walked in stages of one (tap, <=64-channel chunk).
//        A operand: activations are bf16 CHANNELS-LAST (B,D,H,W,C), so a voxel's channel chunk of one tap is one contiguous run:
//        each thread gathers its row with 128-bit loads — WrapPadding3D (common_blocks.py:448-503: zeros along depth / height, wrap
//        along width) is folded into the gather — and writes it k-chunk-major into the stage buffer.  B operand: the layer's weights
//        pre-packed per (n-tile, tap, chunk) as a ready k-chunk-major block, fetched by ONE bulk copy (TMA) per stage.  3-stage ring,
//        completion through tcgen05.commit -> mbarrier; the gather of stage i+1 overlaps the MMAs of stage i.  A second input pointer
//        makes the U-Net's torch.cat((upsampled, skip), 1) free.
//   conv3d_cout1_kernel  the two single-output-channel layers of the last decoder (128 -> 1, 1 -> 1): fp32 SIMT, fp32 output.
//   avgpool / trilinear  AvgPool3d(2) and the x2 trilinear upsampling (align_corners=False) of UNet2.forward on bf16 channels-last.
// Numerics: bf16 operands, fp32 accumulation and activation (the reference's cuDNN convolutions run in TF32 on the same GPUs).
#include <cuda_bf16.h>

#include "render_device.cuh"
#include "umma.cuh"

namespace pgrf {

constexpr int kConvStages = 3;
constexpr int kConvRows = 128;

struct ConvParams {
  const __nv_bfloat16* xa; int Ca;     // first input, channels-last, channel count (multiple of 16)
  const __nv_bfloat16* xb; int Cb;     // second input of a concatenation (or null / 0)
  const unsigned char* wpk;            // packed weights [n_tiles][27][n_cc][KC/8][NT][8] bf16
  const float* bias;                   // [Cout] (padded)
  __nv_bfloat16* y; int Cout;          // output, channels-last, Cout (multiple of 16)
  int B, D, H, W, KC, n_cc, act;
  long long n_vox;
};

template <int NT>
__global__ void __launch_bounds__(kConvRows, 2) conv3d_igemm_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_full[kConvStages], bar_free[kConvStages], bar_done;
  const int r = threadIdx.x, warp = r >> 5;
  const int a_bytes = p.KC * kConvRows * 2, b_bytes = p.KC * NT * 2;
  unsigned char* As = smem;
  unsigned char* Bs = smem + kConvStages * a_bytes;
  if (warp == 0) umma::tmem_alloc(&tmem_base_s, NT < 32 ? 32 : NT);
  if (r == 0) {
    for (int s = 0; s < kConvStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_free[s], 1); }
    mbar_init(&bar_done, 1);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s;

  // this thread's output voxel
  const long long v = (long long)blockIdx.x * kConvRows + r;
  const bool row_ok = v < p.n_vox;
  long long t = row_ok ? v : 0;
  const int x = (int)(t % p.W); t /= p.W;
  const int y = (int)(t % p.H); t /= p.H;
  const int d = (int)(t % p.D);
  const int b = (int)(t / p.D);
  const int nt = blockIdx.y;
  const unsigned char* wt = p.wpk + (size_t)nt * 27 * p.n_cc * b_bytes;
  const int n_it = 27 * p.n_cc;
  const int kch = p.KC >> 3;                         // 16-byte chunks per row and stage

#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    const int s = it % kConvStages, use = it / kConvStages;
    if (it >= kConvStages) mbar_wait(&bar_free[s], (use - 1) & 1);      // the MMAs that read this stage have completed
    if (r == 0) {
      mbar_expect_tx(&bar_full[s], b_bytes);
      bulk_g2s(Bs + s * b_bytes, wt + (size_t)it * b_bytes, b_bytes, &bar_full[s]);
    }
    const int tap = it / p.n_cc, cc = it - tap * p.n_cc;
    const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int dd = d + kd - 1, yy = y + kh - 1;
    int xx = x + kw - 1;
    xx = xx < 0 ? xx + p.W : (xx >= p.W ? xx - p.W : xx);             // wrap along width
    const bool ok = row_ok && dd >= 0 && dd < p.D && yy >= 0 && yy < p.H;   // zeros along depth / height
    const int c0 = cc * p.KC;
    const __nv_bfloat16* src;
    if (c0 < p.Ca) src = p.xa + ((((size_t)b * p.D + dd) * p.H + yy) * p.W + xx) * p.Ca + c0;
    else src = p.xb + ((((size_t)b * p.D + dd) * p.H + yy) * p.W + xx) * p.Cb + (c0 - p.Ca);
    unsigned char* arow = As + s * a_bytes + (size_t)r * 16;
#pragma unroll 4
    for (int j = 0; j < kch; ++j) {
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      if (ok) q = __ldg(reinterpret_cast<const uint4*>(src) + j);
      *reinterpret_cast<uint4*>(arow + (size_t)j * kConvRows * 16) = q;
    }
    umma::fence_smem_to_async();
    __syncthreads();
    if (r == 0) {
      mbar_wait(&bar_full[s], use & 1);
      umma::fence_after_sync();
      umma::gemm_issue(tb, As + s * a_bytes, kConvRows, Bs + s * b_bytes, NT, NT, p.KC, it > 0);
      umma::commit(&bar_free[s]);
    }
  }
  if (r == 0) umma::commit(&bar_done);
  mbar_wait(&bar_done, 0);
  umma::fence_after_sync();

  // epilogue: + bias, LeakyReLU, bf16, 16-byte channel runs (the TMEM loads are warp-collective: only the stores are predicated)
  {
    const uint32_t tq = tb + ((uint32_t)(warp * 32) << 16);
    __nv_bfloat16* dst = p.y + (size_t)(row_ok ? v : 0) * p.Cout + (size_t)nt * NT;
    const float* bias = p.bias + nt * NT;
#pragma unroll
    for (int c = 0; c < NT; c += 16) {
      float a[16];
      umma::ld16(tq + c, a);
      uint32_t q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float u0 = a[2 * i] + __ldg(bias + c + 2 * i), u1 = a[2 * i + 1] + __ldg(bias + c + 2 * i + 1);
        if (p.act) { u0 = u0 > 0.f ? u0 : 0.01f * u0; u1 = u1 > 0.f ? u1 : 0.01f * u1; }
        q[i] = umma::pack2(u0, u1);
      }
      if (row_ok) {
        reinterpret_cast<uint4*>(dst + c)[0] = make_uint4(q[0], q[1], q[2], q[3]);
        reinterpret_cast<uint4*>(dst + c)[1] = make_uint4(q[4], q[5], q[6], q[7]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tb, NT < 32 ? 32 : NT);
}

// single output channel (last decoder): fp32 SIMT.  in: bf16 channels-last (two concatenated inputs) or fp32 single channel
__global__ void __launch_bounds__(256) conv3d_cout1_kernel(const __nv_bfloat16* __restrict__ xa, int Ca, const __nv_bfloat16* __restrict__ xb,
                                                           int Cb, const float* __restrict__ xf, const float* __restrict__ w /* [27][Cin] */,
                                                           float bias, int B, int D, int H, int W, int act, float* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (long long)B * D * H * W) return;
  long long t = v;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const int b = (int)(t / D);
  const int Cin = xf ? 1 : Ca + Cb;
  float acc = bias;
  for (int tap = 0; tap < 27; ++tap) {
    const int dd = d + tap / 9 - 1, yy = y + (tap / 3) % 3 - 1;
    int xx = x + tap % 3 - 1;
    xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
    if (dd < 0 || dd >= D || yy < 0 || yy >= H) continue;
    const size_t nv = (((size_t)b * D + dd) * H + yy) * W + xx;
    const float* wt = w + tap * Cin;
    if (xf) { acc = fmaf(__ldg(xf + nv), wt[0], acc); continue; }
    const uint4* pa = reinterpret_cast<const uint4*>(xa + nv * Ca);
    for (int j = 0; j < Ca / 8; ++j) {
      const uint4 q = __ldg(pa + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc = fmaf(f.x, wt[8 * j + 2 * i], acc); acc = fmaf(f.y, wt[8 * j + 2 * i + 1], acc); }
    }
    if (xb) {
      const uint4* pb = reinterpret_cast<const uint4*>(xb + nv * Cb);
      for (int j = 0; j < Cb / 8; ++j) {
        const uint4 q = __ldg(pb + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc = fmaf(f.x, wt[Ca + 8 * j + 2 * i], acc); acc = fmaf(f.y, wt[Ca + 8 * j + 2 * i + 1], acc); }
      }
    }
  }
  if (act) acc = acc > 0.f ? acc : 0.01f * acc;
  out[v] = acc;
}

// AvgPool3d(2) on bf16 channels-last: thread = (output voxel, 8-channel chunk)
__global__ void __launch_bounds__(256) avgpool3d2_kernel(const __nv_bfloat16* __restrict__ x, int B, int D, int H, int W, int C,
                                                         __nv_bfloat16* __restrict__ y) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Do * Ho * Wo * C8) return;
  long long t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int dO = (int)(t % Do);
  const int b = (int)(t / Do);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; ++k) {
    const int dd = 2 * dO + (k >> 2), yy = 2 * yo + ((k >> 1) & 1), xx = 2 * xo + (k & 1);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + ((((size_t)b * D + dd) * H + yy) * W + xx) * C) + c8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
  }
  uint4 o;
  o.x = umma::pack2(acc[0] * 0.125f, acc[1] * 0.125f); o.y = umma::pack2(acc[2] * 0.125f, acc[3] * 0.125f);
  o.z = umma::pack2(acc[4] * 0.125f, acc[5] * 0.125f); o.w = umma::pack2(acc[6] * 0.125f, acc[7] * 0.125f);
  reinterpret_cast<uint4*>(y + i / C8 * C)[c8] = o;
}

// x2 trilinear upsampling, align_corners=False (ATen area_pixel_compute_source_index): thread = (output voxel, 8-channel chunk)
__device__ __forceinline__ void up_src(int o, int n, int& i0, int& i1, float& l) {
  float s = 0.5f * ((float)o + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l = s - (float)i0;
}
__global__ void __launch_bounds__(256) upsample3d2_kernel(const __nv_bfloat16* __restrict__ x, int B, int D, int H, int W, int C,
                                                          __nv_bfloat16* __restrict__ y) {
  const int Do = 2 * D, Ho = 2 * H, Wo = 2 * W, C8 = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Do * Ho * Wo * C8) return;
  long long t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int dO = (int)(t % Do);
  const int b = (int)(t / Do);
  int d0, d1, y0, y1, x0, x1;
  float ld, ly, lx;
  up_src(dO, D, d0, d1, ld); up_src(yo, H, y0, y1, ly); up_src(xo, W, x0, x1, lx);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 8; ++k) {
    const int dd = (k & 4) ? d1 : d0, yy = (k & 2) ? y1 : y0, xx = (k & 1) ? x1 : x0;
    const float wgt = ((k & 4) ? ld : 1.f - ld) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + ((((size_t)b * D + dd) * H + yy) * W + xx) * C) + c8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] = fmaf(wgt, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(wgt, f.y, acc[2 * j + 1]); }
  }
  uint4 o;
  o.x = umma::pack2(acc[0], acc[1]); o.y = umma::pack2(acc[2], acc[3]); o.z = umma::pack2(acc[4], acc[5]); o.w = umma::pack2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(y + i / C8 * C)[c8] = o;
}

// fp32 (B,C,D,H,W) with arbitrary element strides -> bf16 channels-last (B,D,H,W,Cpad), zero padded channels
__global__ void __launch_bounds__(256) to_bf16_cl_kernel(const float* __restrict__ x, long long sb, long long sc, long long sd, long long sh,
                                                         long long sw, int B, int C, int D, int H, int W, int Cpad,
                                                         __nv_bfloat16* __restrict__ y) {
  const int C8 = Cpad / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * D * H * W * C8) return;
  long long t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xx = (int)(t % W); t /= W;
  const int yy = (int)(t % H); t /= H;
  const int dd = (int)(t % D);
  const int b = (int)(t / D);
  const float* src = x + b * sb + dd * sd + yy * sh + xx * sw;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { const int c = 8 * c8 + k; v[k] = c < C ? __ldg(src + c * sc) : 0.f; }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(y + i / C8 * Cpad)[c8] = o;
}

}  // namespace pgrf

using namespace pgrf;

static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" int pgrf_conv3d_to_bf16_cl(const float* x, long long sb, long long sc, long long sd, long long sh, long long sw, int B, int C,
                                      int D, int H, int W, int Cpad, void* y, void* stream) {
  PGRF_REQUIRE(x && y && B >= 1 && C >= 1 && Cpad >= C && Cpad % 8 == 0, "conv3d_to_bf16_cl: bad arguments");
  const long long n = (long long)B * D * H * W * (Cpad / 8);
  to_bf16_cl_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, sb, sc, sd, sh, sw, B, C, D, H, W, Cpad, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_igemm_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, int Cout,
                                     int B, int D, int H, int W, int act, void* stream) {
  PGRF_REQUIRE(xa && wpk && bias && y, "conv3d: null pointer argument");
  PGRF_REQUIRE(Ca >= 16 && Ca % 16 == 0 && Cb % 16 == 0 && (Cb == 0 || xb) && Cout >= 16 && Cout % 16 == 0, "conv3d: channel counts must "
               "be multiples of 16 (Ca=%d Cb=%d Cout=%d)", Ca, Cb, Cout);
  ConvParams p;
  p.xa = (const __nv_bfloat16*)xa; p.Ca = Ca; p.xb = (const __nv_bfloat16*)xb; p.Cb = Cb;
  p.wpk = (const unsigned char*)wpk; p.bias = bias; p.y = (__nv_bfloat16*)y; p.Cout = Cout;
  p.B = B; p.D = D; p.H = H; p.W = W; p.act = act;
  const int Cin = Ca + Cb;
  int KC = Cin < 64 ? Cin : 64;
  while (Ca % KC || Cb % KC) KC >>= 1;           // a stage never straddles the two inputs of a concatenation
  PGRF_REQUIRE(KC >= 16, "conv3d: no common chunk size for Ca=%d Cb=%d", Ca, Cb);
  p.KC = KC; p.n_cc = Cin / KC;
  p.n_vox = (long long)B * D * H * W;
  const int NT = Cout >= 128 ? 128 : Cout;
  PGRF_REQUIRE(NT == 16 || NT == 32 || NT == 64 || NT == 128, "conv3d: Cout=%d unsupported", Cout);
  PGRF_REQUIRE(Cout % NT == 0, "conv3d: Cout=%d not a multiple of the channel tile", Cout);
  dim3 grid(blocks_for(p.n_vox, kConvRows), (unsigned)(Cout / NT));
  const size_t smem = (size_t)kConvStages * (KC * kConvRows * 2 + KC * NT * 2);
  cudaStream_t st = (cudaStream_t)stream;
#define PGRF_CONV(N)                                                                                                   \
  case N:                                                                                                              \
    PGRF_CUDA(cudaFuncSetAttribute(conv3d_igemm_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    conv3d_igemm_kernel<N><<<grid, kConvRows, smem, st>>>(p);                                                          \
    break;
  switch (NT) { PGRF_CONV(16) PGRF_CONV(32) PGRF_CONV(64) PGRF_CONV(128) }
#undef PGRF_CONV
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_cout1_fwd(const void* xa, int Ca, const void* xb, int Cb, const float* xf, const float* w, float bias, int B,
                                     int D, int H, int W, int act, float* out, void* stream) {
  PGRF_REQUIRE(w && out && (xf || (xa && Ca % 8 == 0 && Cb % 8 == 0)), "conv3d_cout1: bad arguments");
  const long long n = (long long)B * D * H * W;
  conv3d_cout1_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)xa, Ca, (const __nv_bfloat16*)xb, Cb, xf, w,
                                                                            bias, B, D, H, W, act, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_avgpool3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "avgpool3d: sizes must be even, C %% 8 == 0");
  const long long n = (long long)B * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  avgpool3d2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, D, H, W, C, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_upsample3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && C % 8 == 0, "upsample3d: C %% 8 == 0");
  const long long n = (long long)B * (2 * D) * (2 * H) * (2 * W) * (C / 8);
  upsample3d2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, D, H, W, C, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

// K1 backward — gradient of the fused spherical-sweep cost volume w.r.t. the feature maps.
//
// The reference obtains it from autograd through grid_sample + abs/mul
// (models/spherical_cost_volume.py:135-230); depth hypotheses and poses carry no gradient
// (built under no_grad / detached, pipeline3_model.py:647,671).  Same decomposition as the forward kernel:
//   CTA = 4 warps = 128 pixels of one ERP row x a chunk of depth hypotheses
//   phase A (lane <-> pixel): recompute the bilinear footprint of every swept view (bit-identical to the forward)
//   phase B (lane <-> (pixel, float4 channel group)): read the upstream gradient (coalesced 128-bit), recompute the
//            warped feature where the cost needs it, scatter 4 x 128-bit vector atomics into the source view, keep the
//            reference-view gradient in registers over the depth chunk and add it once at the end.
// grad_images must be zeroed by the caller; float atomics make the summation order run-dependent (as in ATen's
// grid_sampler_2d_backward on CUDA).
#include "cost_volume.cuh"

namespace pgrf {

struct CvBwdParams {
  CvParams f;
  const float* grad_out;   // (B,D,H,W,C) contiguous
  float* grad_images;      // (B,S,H,W,C)
};

__device__ __forceinline__ void red_add4(float4* dst, const float4& v) { atomicAdd(dst, v); }

template <int C>
__global__ void __launch_bounds__(kCvThreads) cost_volume_bwd_kernel(const CvBwdParams q) {
  const CvParams& p = q.f;
  constexpr int CG = C / 4;
  constexpr int PPS = 32 / CG;
  constexpr int NSUB = 32 / PPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_A[kMaxSrc][12];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (p.W + kCvThreads - 1) / kCvThreads;
  const int y = blockIdx.x / tiles_x;
  const int x_warp = (blockIdx.x % tiles_x) * kCvThreads + warp * 32;
  const int b = blockIdx.z;
  const int d_begin = blockIdx.y * p.d_chunk;
  const int d_end = min(p.D, d_begin + p.d_chunk);
  TapRec* rec = reinterpret_cast<TapRec*>(smem_raw) + warp * (p.n_src * 32);

  if (threadIdx.x < p.n_src) relative_pose(p, b, threadIdx.x, s_A[threadIdx.x]);
  __syncthreads();

  const int x = x_warp + lane;
  float rx, ry, rz;
  pixel_ray(p, min(x, p.W - 1), y, rx, ry, rz);

  const int pp = lane / CG, cg = lane % CG;
  const size_t view_f4 = (size_t)p.H * p.W * CG;
  const float4* img4 = reinterpret_cast<const float4*>(p.images) + (size_t)b * p.S * view_f4;
  float4* gimg4 = reinterpret_cast<float4*>(q.grad_images) + (size_t)b * p.S * view_f4;
  const int row_f4 = p.W * CG;
  const size_t plane = (size_t)p.H * p.W;
  const float inv_div = p.divisor != 0.f ? 1.f / p.divisor : 1.f;

  float4 ref[NSUB], gref[NSUB];
#pragma unroll
  for (int j = 0; j < NSUB; ++j) {
    const int px = x_warp + j * PPS + pp;
    ref[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    gref[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < p.W) ref[j] = ldg4(img4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + px) * CG + cg);
  }

  for (int d = d_begin; d < d_end; ++d) {
    float depth;
    if (p.depth_volume) depth = (x < p.W) ? __ldg(p.depth_volume + ((size_t)b * p.D + d) * plane + (size_t)y * p.W + x) : 1.f;
    else depth = __ldg(p.depths + d);
    for (int s = 0; s < p.n_src; ++s) {
      const float* A = s_A[s];
      const float ax = A[0] * rx + A[1] * ry + A[2] * rz;
      const float ay = A[3] * rx + A[4] * ry + A[5] * rz;
      const float az = A[6] * rx + A[7] * ry + A[8] * rz;
      const float cx = fmaf(depth, ax, A[9]), cy = fmaf(depth, ay, A[10]), cz = fmaf(depth, az, A[11]);
      float u, v;
      point_uv(p.dataset, cx, cy, cz, u, v);
      const float ix = ((u + 1.f) / 2.f) * (float)(p.W - 1);
      const float iy = ((v + 1.f) / 2.f) * (float)(p.H - 1);
      float x0f = floorf(ix), y0f = floorf(iy);
      x0f = fminf(fmaxf(x0f, 0.f), (float)(p.W - 2));
      y0f = fminf(fmaxf(y0f, 0.f), (float)(p.H - 2));
      TapRec r;
      r.tx = ix - x0f;
      r.ty = iy - y0f;
      r.off4 = ((int)y0f * p.W + (int)x0f) * CG;
      r.pad = 0;
      rec[s * 32 + lane] = r;
    }
    __syncwarp();

    const float4* g_cl = reinterpret_cast<const float4*>(q.grad_out) + ((((size_t)b * p.D + d) * p.H + y) * p.W + x_warp) * CG + lane;
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      const int pi = j * PPS + pp;
      if (x_warp + pi >= p.W) continue;
      float4 g = __ldcs(g_cl + j * 32);
      g.x *= inv_div; g.y *= inv_div; g.z *= inv_div; g.w *= inv_div;
      for (int s = 0; s < p.n_src; ++s) {
        const TapRec r = rec[s * 32 + pi];
        const size_t voff = (size_t)p.src_views[s] * view_f4 + cg + r.off4;
        const float tx1 = 1.f - r.tx, ty1 = 1.f - r.ty;
        const float wnw = tx1 * ty1, wne = r.tx * ty1, wsw = tx1 * r.ty, wse = r.tx * r.ty;
        float4 dw = g;                                     // d loss / d warped feature
        if (p.cost_type != PGRF_COST_NONE) {
          const float4* row0 = img4 + voff;
          const float4* row1 = row0 + row_f4;
          const float4 nw = ldg4(row0), ne = ldg4(row0 + CG), sw = ldg4(row1), se = ldg4(row1 + CG);
          float4 val;
          val.x = nw.x * wnw; val.y = nw.y * wnw; val.z = nw.z * wnw; val.w = nw.w * wnw;
          val.x = fmaf(ne.x, wne, val.x); val.y = fmaf(ne.y, wne, val.y); val.z = fmaf(ne.z, wne, val.z); val.w = fmaf(ne.w, wne, val.w);
          val.x = fmaf(sw.x, wsw, val.x); val.y = fmaf(sw.y, wsw, val.y); val.z = fmaf(sw.z, wsw, val.z); val.w = fmaf(sw.w, wsw, val.w);
          val.x = fmaf(se.x, wse, val.x); val.y = fmaf(se.y, wse, val.y); val.z = fmaf(se.z, wse, val.z); val.w = fmaf(se.w, wse, val.w);
          const float4 rf = ref[j];
          if (p.cost_type == PGRF_COST_ABS_DIFF) {
            // torch.abs backward: grad * sign(x), sign(0) = 0
            const float sx = (float)((val.x > rf.x) - (val.x < rf.x)), sy = (float)((val.y > rf.y) - (val.y < rf.y));
            const float sz = (float)((val.z > rf.z) - (val.z < rf.z)), sw_ = (float)((val.w > rf.w) - (val.w < rf.w));
            dw.x = g.x * sx; dw.y = g.y * sy; dw.z = g.z * sz; dw.w = g.w * sw_;
            gref[j].x -= dw.x; gref[j].y -= dw.y; gref[j].z -= dw.z; gref[j].w -= dw.w;
          } else {  // dot
            dw.x = g.x * rf.x; dw.y = g.y * rf.y; dw.z = g.z * rf.z; dw.w = g.w * rf.w;
            gref[j].x = fmaf(g.x, val.x, gref[j].x); gref[j].y = fmaf(g.y, val.y, gref[j].y);
            gref[j].z = fmaf(g.z, val.z, gref[j].z); gref[j].w = fmaf(g.w, val.w, gref[j].w);
          }
        }
        float4* grow0 = gimg4 + voff;
        float4* grow1 = grow0 + row_f4;
        red_add4(grow0, make_float4(dw.x * wnw, dw.y * wnw, dw.z * wnw, dw.w * wnw));
        red_add4(grow0 + CG, make_float4(dw.x * wne, dw.y * wne, dw.z * wne, dw.w * wne));
        red_add4(grow1, make_float4(dw.x * wsw, dw.y * wsw, dw.z * wsw, dw.w * wsw));
        red_add4(grow1 + CG, make_float4(dw.x * wse, dw.y * wse, dw.z * wse, dw.w * wse));
      }
    }
    __syncwarp();
  }
  if (p.cost_type != PGRF_COST_NONE) {
#pragma unroll
    for (int j = 0; j < NSUB; ++j) {
      const int px = x_warp + j * PPS + pp;
      if (px < p.W) red_add4(gimg4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + px) * CG + cg, gref[j]);
    }
  }
}

template <int C>
static int launch_bwd(const CvBwdParams& q, cudaStream_t st) {
  const CvParams& p = q.f;
  const size_t smem = (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec);
  dim3 grid((unsigned)(((p.W + kCvThreads - 1) / kCvThreads) * p.H), (unsigned)((p.D + p.d_chunk - 1) / p.d_chunk), (unsigned)p.B);
  cost_volume_bwd_kernel<C><<<grid, kCvThreads, smem, st>>>(q);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_cost_volume_bwd(const float* grad_out, const float* images, int B, int S, int H, int W, int C,
                                    const float* depths, const float* depth_volume, int D, const float* rots, const float* trans,
                                    int ref_idx, const int* src_views, int n_src, float divisor, int dataset, int cost_type,
                                    float* grad_images, void* stream) {
  PGRF_REQUIRE(grad_out && grad_images, "cost_volume_bwd: null pointer argument");
  PGRF_REQUIRE((((uintptr_t)grad_out | (uintptr_t)grad_images) & 15) == 0, "cost_volume_bwd: grad_out/grad_images must be 16-byte aligned");
  CvBwdParams q;
  const int frc = cv_fill_params(q.f, images, B, S, H, W, C, depths, depth_volume, D, rots, trans, ref_idx, src_views, n_src, divisor,
                                 dataset, cost_type);
  if (frc != PGRF_OK) return frc;
  q.grad_out = grad_out;
  q.grad_images = grad_images;
  // reference-view gradients are added once per (CTA, pixel): prefer long depth chunks, but keep >= ~2 waves of CTAs
  const long long ctas_per_chunk = (long long)((W + kCvThreads - 1) / kCvThreads) * H * B;
  int n_chunks = (int)((148LL * 8 * 2 + ctas_per_chunk - 1) / ctas_per_chunk);
  if (n_chunks < 1) n_chunks = 1;
  int d_chunk = (D + n_chunks - 1) / n_chunks;
  if (d_chunk < 4) d_chunk = D < 4 ? D : 4;
  q.f.d_chunk = d_chunk;
  PGRF_REQUIRE((D + d_chunk - 1) / d_chunk <= 65535, "cost_volume_bwd: too many depth chunks");
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 4: return launch_bwd<4>(q, st);
    case 8: return launch_bwd<8>(q, st);
    case 16: return launch_bwd<16>(q, st);
    case 32: return launch_bwd<32>(q, st);
    default: return launch_bwd<64>(q, st);
  }
}

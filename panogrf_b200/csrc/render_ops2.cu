// More stand-alone operators of the reference's functional / module API (not used by the fused renderer):
//   compute_prob            — MixtureLogisticsDistDecoder.compute_prob (dist_decoder.py:113-140) with
//                             get_near_far_points(is_ref=True) (:6-51)
//   interpolate_feature_map — render_ops.py:126-143 -> ops.py:32-52: bilinear, border padding, any channel count, NCHW maps
//   depth2points_spherical  — render_ops.py:76-106 (+ ray_utils.py:4-16): pixel -> unit ray -> world points and directions
#include "render_device.cuh"

namespace pgrf {

__global__ void __launch_bounds__(256) compute_prob_kernel(const float* __restrict__ depth, const float* __restrict__ interval,
                                                           long long interval_view_stride, const float* __restrict__ mean,
                                                           const float* __restrict__ var, const float* __restrict__ vis,
                                                           const float* __restrict__ aw, const float* __restrict__ depth_range,
                                                           int rfn, long long n, int dn, int is_ref, float* __restrict__ alpha,
                                                           float* __restrict__ visibility, float* __restrict__ hit_prob) {
  const long long total = (long long)rfn * n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / n);
    const long long g = i % n;
    const int s = (int)(g % dn);
    const float* iv = interval + (size_t)v * interval_view_stride;
    const float d_s = __ldg(iv + g);
    const float d_prev = s > 0 ? __ldg(iv + g - 1) : d_s;
    const float rnear = __ldg(depth_range + 2 * v), rfar = __ldg(depth_range + 2 * v + 1);
    const float dv = inv_norm(fmaxf(__ldg(depth + i), 1e-5f), rnear, rfar);
    float nearp = dv - d_prev / 2.f, farp = dv + d_s / 2.f;
    if (!is_ref) {
      // the rays' own distribution (dist_decoder.py:37-45): bin edges are the midpoints between consecutive normalised inverse depths,
      // half an interval beyond the first / last sample
      nearp = s > 0 ? (inv_norm(fmaxf(__ldg(depth + i - 1), 1e-5f), rnear, rfar) + dv) / 2.f : dv - d_s / 2.f;
      farp = s + 1 < dn ? (dv + inv_norm(fmaxf(__ldg(depth + i + 1), 1e-5f), rnear, rfar)) / 2.f : dv + d_s / 2.f;
    }
    const float a = __ldg(aw + i);
    const float mix[2] = {a, 1.f - a};
    float vsum = 0.f, hp = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float mu = __ldg(mean + 2 * i + j), va = __ldg(var + 2 * i + j);
      float cdf0 = 0.5f + 0.5f * tanhf((nearp - mu) * va);
      float cdf1 = 0.5f + 0.5f * tanhf((farp - mu) * va);
      if (vis) { const float vv = __ldg(vis + i); cdf0 *= vv; cdf1 *= vv; }
      vsum += (1.f - cdf0) * mix[j];
      hp += (cdf1 - cdf0) * mix[j];
    }
    alpha[i] = logf(hp / (vsum - hp + 1e-5f) + 1e-5f);
    visibility[i] = vsum;
    hit_prob[i] = hp;
  }
}

// thread = (view, point); channels looped (planar NCHW taps: 4 scalar loads per channel, L1/L2 resident maps)
__global__ void __launch_bounds__(256) interpolate_kernel(const float* __restrict__ feats, int rfn, int C, int fh, int fw,
                                                          const float* __restrict__ pix, long long pn, int h, int w,
                                                          float* __restrict__ out) {
  const size_t plane = (size_t)fh * fw;
  const long long total = (long long)rfn * pn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / pn);
    const Footprint f = border_footprint(__ldg(pix + 2 * i), __ldg(pix + 2 * i + 1), h, w, fh, fw);
    const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
    const float wnw = tx1 * ty1, wne = f.tx * ty1, wsw = tx1 * f.ty, wse = f.tx * f.ty;
    const float* base = feats + (size_t)v * C * plane + f.off;
    for (int c = 0; c < C; ++c) {
      const float* m = base + c * plane;
      float o = __ldg(m) * wnw;
      o = fmaf(__ldg(m + f.dx), wne, o);
      o = fmaf(__ldg(m + f.dy * fw), wsw, o);
      o = fmaf(__ldg(m + f.dy * fw + f.dx), wse, o);
      out[(size_t)i * C + c] = o;
    }
  }
}

__global__ void __launch_bounds__(256) depth2points_kernel(const float* __restrict__ coords, const float* __restrict__ depth,
                                                           long long depth_ray_stride, const float* __restrict__ c2w, int dataset,
                                                           int H, int W, long long rn, int dn, float* __restrict__ pts,
                                                           float* __restrict__ dir) {
  const long long total = rn * dn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ray = i / dn;
    const int s = (int)(i % dn);
    const float cx = __ldg(coords + 2 * ray), cy = __ldg(coords + 2 * ray + 1);
    float dx, dy, dz;
    equi_unit_dir(dataset, (float)(long long)cx, (float)(long long)cy, H, W, dx, dy, dz);
    const float r0 = c2w[0] * dx + c2w[1] * dy + c2w[2] * dz;
    const float r1 = c2w[4] * dx + c2w[5] * dy + c2w[6] * dz;
    const float r2 = c2w[8] * dx + c2w[9] * dy + c2w[10] * dz;
    const float t = __ldg(depth + ray * depth_ray_stride + s);
    pts[3 * i] = c2w[3] + r0 * t; pts[3 * i + 1] = c2w[7] + r1 * t; pts[3 * i + 2] = c2w[11] + r2 * t;
    const float nrm = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
    dir[3 * i] = -r0 / nrm; dir[3 * i + 1] = -r1 / nrm; dir[3 * i + 2] = -r2 / nrm;
  }
}

static int grid_for(long long total) {
  const long long b = (total + 255) / 256;
  return (int)(b < 148LL * 16 ? b : 148LL * 16);
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_compute_prob_fwd(const float* depth, const float* interval, int interval_per_view, const float* mean,
                                     const float* var, const float* vis, const float* aw, const float* depth_range, int rfn,
                                     long long n, int dn, float* alpha, float* visibility, float* hit_prob, void* stream) {
  PGRF_REQUIRE(depth && interval && mean && var && aw && depth_range && alpha && visibility && hit_prob, "compute_prob: null pointer argument");
  PGRF_REQUIRE(rfn >= 1 && n >= 1 && dn >= 1 && n % dn == 0, "compute_prob: rfn=%d n=%lld dn=%d (n must be a multiple of dn)", rfn, n, dn);
  compute_prob_kernel<<<grid_for((long long)rfn * n), 256, 0, (cudaStream_t)stream>>>(depth, interval, interval_per_view ? n : 0, mean, var, vis,
                                                                                    aw, depth_range, rfn, n, dn, 1, alpha, visibility, hit_prob);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_compute_prob_que_fwd(const float* depth, const float* interval, const float* mean, const float* var, const float* vis,
                                         const float* aw, const float* depth_range, int qn, long long n, int dn, float* alpha,
                                         float* visibility, float* hit_prob, void* stream) {
  PGRF_REQUIRE(depth && interval && mean && var && aw && depth_range && alpha && visibility && hit_prob, "compute_prob_que: null pointer argument");
  PGRF_REQUIRE(qn >= 1 && n >= 1 && dn >= 1 && n % dn == 0, "compute_prob_que: qn=%d n=%lld dn=%d (n must be a multiple of dn)", qn, n, dn);
  compute_prob_kernel<<<grid_for((long long)qn * n), 256, 0, (cudaStream_t)stream>>>(depth, interval, n, mean, var, vis, aw, depth_range, qn, n, dn,
                                                                                   0, alpha, visibility, hit_prob);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_interpolate_feature_map_fwd(const float* feats, int rfn, int C, int fh, int fw, const float* pix, long long pn, int h,
                                                int w, float* out, void* stream) {
  PGRF_REQUIRE(feats && pix && out, "interpolate_feature_map: null pointer argument");
  PGRF_REQUIRE(rfn >= 1 && C >= 1 && fh >= 1 && fw >= 1 && pn >= 1 && h >= 2 && w >= 2, "interpolate_feature_map: bad sizes");
  interpolate_kernel<<<grid_for((long long)rfn * pn), 256, 0, (cudaStream_t)stream>>>(feats, rfn, C, fh, fw, pix, pn, h, w, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_depth2points_fwd(const float* coords, const float* depth, int depth_ray_stride, const float* c2w, int dataset, int H,
                                     int W, long long rn, int dn, float* pts, float* dir, void* stream) {
  PGRF_REQUIRE(coords && depth && c2w && pts && dir, "depth2points: null pointer argument");
  PGRF_REQUIRE(dataset >= 0 && dataset <= 3, "depth2points: unknown dataset id %d", dataset);
  PGRF_REQUIRE(rn >= 1 && dn >= 1 && H >= 2 && W >= 2 && (depth_ray_stride == 0 || depth_ray_stride == dn), "depth2points: bad sizes");
  depth2points_kernel<<<grid_for(rn * dn), 256, 0, (cudaStream_t)stream>>>(coords, depth, depth_ray_stride, c2w, dataset, H, W, rn, dn, pts, dir);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward passes of the two cheap differentiable stages (the reference gets them from autograd):
//   composite_bwd            — d/d(density|alpha), d/d(colors) of hit_prob / pixel_colors / render_depth
//   interpolate_feature_map_bwd — d/d(feats) of the bilinear border gather (scatter-add of the 4 tap weights)
// ------------------------------------------------------------------------------------------------
namespace pgrf {

// one warp per ray.  h_s = a_s T_s, T_s = prod_{j<s} x_j, x_j = 1 - a_j + 1e-10:
//   dL/da_s = G_s T_s - (sum_{k>s} G_k h_k) / x_s,   G_s = g_hit[s] + <g_pix, c_s> + g_depth z_s
__global__ void __launch_bounds__(256) composite_bwd_kernel(const float* __restrict__ density, const float* __restrict__ alpha_in,
                                                            const float* __restrict__ colors, const float* __restrict__ depth,
                                                            int depth_ray_stride, const float* __restrict__ g_hit,
                                                            const float* __restrict__ g_pix, const float* __restrict__ g_depth,
                                                            float* __restrict__ grad_in, float* __restrict__ grad_colors, int rn, int dn) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* alpha = sm + warp * 4 * dn;
  float* trans = alpha + dn;
  float* gh = trans + dn;      // G_s * h_s, then its exclusive suffix sum
  float* G = gh + dn;
  for (long long ray = (long long)blockIdx.x * 8 + warp; ray < rn; ray += (long long)gridDim.x * 8) {
    for (int s = lane; s < dn; s += 32)
      alpha[s] = alpha_in ? __ldg(alpha_in + ray * dn + s) : 1.f - expf(-fmaxf(__ldg(density + ray * dn + s), 0.f));
    __syncwarp();
    if (lane == 0) {
      float t = 1.f;
      for (int s = 0; s < dn; ++s) { trans[s] = t; t = t * (1.f - alpha[s] + 1e-10f); }
    }
    __syncwarp();
    const float gp0 = g_pix ? __ldg(g_pix + ray * 3) : 0.f, gp1 = g_pix ? __ldg(g_pix + ray * 3 + 1) : 0.f,
                gp2 = g_pix ? __ldg(g_pix + ray * 3 + 2) : 0.f;
    const float gd = g_depth ? __ldg(g_depth + ray) : 0.f;
    for (int s = lane; s < dn; s += 32) {
      float g = g_hit ? __ldg(g_hit + ray * dn + s) : 0.f;
      const float h = alpha[s] * trans[s];
      if (colors && g_pix) {
        const float* c = colors + (ray * dn + s) * 3;
        g += gp0 * __ldg(c) + gp1 * __ldg(c + 1) + gp2 * __ldg(c + 2);
        if (grad_colors) {
          float* gc = grad_colors + (ray * dn + s) * 3;
          gc[0] = gp0 * h; gc[1] = gp1 * h; gc[2] = gp2 * h;
        }
      } else if (grad_colors) {
        float* gc = grad_colors + (ray * dn + s) * 3;
        gc[0] = gc[1] = gc[2] = 0.f;
      }
      if (depth && g_depth) g += gd * __ldg(depth + ray * depth_ray_stride + s);
      G[s] = g;
      gh[s] = g * h;
    }
    __syncwarp();
    if (lane == 0) {   // exclusive suffix sum
      float acc = 0.f;
      for (int s = dn - 1; s >= 0; --s) { const float v = gh[s]; gh[s] = acc; acc += v; }
    }
    __syncwarp();
    for (int s = lane; s < dn; s += 32) {
      const float x = 1.f - alpha[s] + 1e-10f;
      float ga = G[s] * trans[s] - gh[s] / x;
      if (!alpha_in) {
        const float sg = __ldg(density + ray * dn + s);
        ga = sg > 0.f ? ga * expf(-sg) : 0.f;
      }
      grad_in[ray * dn + s] = ga;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) interpolate_bwd_kernel(const float* __restrict__ grad_out, int rfn, int C, int fh, int fw,
                                                              const float* __restrict__ pix, long long pn, int h, int w,
                                                              float* __restrict__ grad_feats) {
  const size_t plane = (size_t)fh * fw;
  const long long total = (long long)rfn * pn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / pn);
    const Footprint f = border_footprint(__ldg(pix + 2 * i), __ldg(pix + 2 * i + 1), h, w, fh, fw);
    const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
    const float wnw = tx1 * ty1, wne = f.tx * ty1, wsw = tx1 * f.ty, wse = f.tx * f.ty;
    float* base = grad_feats + (size_t)v * C * plane + f.off;
    for (int c = 0; c < C; ++c) {
      const float g = __ldg(grad_out + (size_t)i * C + c);
      float* m = base + c * plane;
      atomicAdd(m, g * wnw);
      if (f.dx) atomicAdd(m + 1, g * wne);
      if (f.dy) atomicAdd(m + fw, g * wsw);
      if (f.dx && f.dy) atomicAdd(m + fw + 1, g * wse);
    }
  }
}

}  // namespace pgrf

extern "C" int pgrf_composite_bwd(const float* density, const float* alpha, const float* colors, const float* depth, int depth_ray_stride,
                                  int rn, int dn, const float* g_hit_prob, const float* g_pixel_colors, const float* g_render_depth,
                                  float* grad_density_or_alpha, float* grad_colors, void* stream) {
  PGRF_REQUIRE((density != nullptr) != (alpha != nullptr), "composite_bwd: pass exactly one of density / alpha");
  PGRF_REQUIRE(grad_density_or_alpha != nullptr, "composite_bwd: null pointer argument");
  PGRF_REQUIRE(rn >= 1 && dn >= 1 && dn <= 1024, "composite_bwd: rn=%d dn=%d", rn, dn);
  const int grid = (rn + 7) / 8 < 148 * 8 ? (rn + 7) / 8 : 148 * 8;
  const size_t smem = 8 * 4 * (size_t)dn * sizeof(float);
  if (smem > 48 * 1024) PGRF_CUDA(cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  composite_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(density, alpha, colors, depth, depth_ray_stride, g_hit_prob, g_pixel_colors,
                                                                g_render_depth, grad_density_or_alpha, grad_colors, rn, dn);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_interpolate_feature_map_bwd(const float* grad_out, int rfn, int C, int fh, int fw, const float* pix, long long pn, int h,
                                                int w, float* grad_feats, void* stream) {
  PGRF_REQUIRE(grad_out && pix && grad_feats, "interpolate_feature_map_bwd: null pointer argument");
  PGRF_REQUIRE(rfn >= 1 && C >= 1 && fh >= 1 && fw >= 1 && pn >= 1 && h >= 2 && w >= 2, "interpolate_feature_map_bwd: bad sizes");
  interpolate_bwd_kernel<<<grid_for((long long)rfn * pn), 256, 0, (cudaStream_t)stream>>>(grad_out, rfn, C, fh, fw, pix, pn, h, w, grad_feats);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

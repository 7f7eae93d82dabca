// Stand-alone operators of the render path, for callers that use the reference's functional API
// (network/render_ops.py) instead of the fused renderer, and for per-kernel roofline measurements:
//   project_gather  — project_points_dict (render_ops.py:234-257) + get_img_feats (renderer.py:180-188):
//                     world points -> every source panorama -> pixel, depth, direction + 3 bilinear gathers
//   composite       — alpha_values2hit_prob (render_ops.py:145-153) + renderer.py:216-218,302-304
//   fine_sample     — sample_fine_depth (render_ops.py:413-473), deterministic u table
//   depth_hypotheses— pipeline3_model.py:723-733,774-815: 5 mono-guided + 59 linear hypotheses, per-pixel sorted
#include "render_device.cuh"

namespace pgrf {

// ------------------------------------------------------------------------------------------------
// K2: projection + gathers.  CTA = 256 threads = 64 points x up to 4 views per tile.
// ------------------------------------------------------------------------------------------------
constexpr int kPgThreads = 256;
constexpr int kPgMaxViews = 4;

struct PgRec {
  int off_rf, off_if, off_im;   // north-west texel index inside the stacked (rfn*h*w) map
  int dxy;                      // bits 0,1: ray_feats  2,3: img_feats  4,5: imgs
  float tx_rf, ty_rf, tx_if, ty_if, tx_im, ty_im;
};

struct PgParams {
  const float* pts;        // (pn,3) world points
  const float* w2c;        // (rfn,3,4)
  const float* imgs_cl; const float* img_feats_cl; const float* ray_feats_cl;
  float* out_pix;          // (rfn,pn,2)
  float* out_depth;        // (rfn,pn)
  float* out_dir;          // (rfn,pn,3)
  float* out_rf;           // (rfn,pn,32)
  float* out_rgb;          // (rfn,pn,3)
  float* out_if;           // (rfn,pn,32) or null
  long long pn;
  int rfn, dataset, H, W, img_h, img_w, if_h, if_w, rf_h, rf_w;
};

template <bool PACKED = true>
__device__ __forceinline__ float4 blend4(const float4* __restrict__ base, int sx, int sy, float tx, float ty) {
  const float4 nw = ldg4(base), ne = ldg4(base + sx), sw = ldg4(base + sy), se = ldg4(base + sy + sx);
  const float tx1 = 1.f - tx, ty1 = 1.f - ty;
  const float wnw = tx1 * ty1, wne = tx * ty1, wsw = tx1 * ty, wse = tx * ty;
  if (!PACKED) {
    float4 o;
    o.x = nw.x * wnw; o.y = nw.y * wnw; o.z = nw.z * wnw; o.w = nw.w * wnw;
    o.x = fmaf(ne.x, wne, o.x); o.y = fmaf(ne.y, wne, o.y); o.z = fmaf(ne.z, wne, o.z); o.w = fmaf(ne.w, wne, o.w);
    o.x = fmaf(sw.x, wsw, o.x); o.y = fmaf(sw.y, wsw, o.y); o.z = fmaf(sw.z, wsw, o.z); o.w = fmaf(sw.w, wsw, o.w);
    o.x = fmaf(se.x, wse, o.x); o.y = fmaf(se.y, wse, o.y); o.z = fmaf(se.z, wse, o.z); o.w = fmaf(se.w, wse, o.w);
    return o;
  }
  // ATen's order (nw, ne, sw, se) on packed fp32 pairs: the same IEEE mul / fma per channel, half the issue slots
  const float2 w0 = make_float2(wnw, wnw), w1 = make_float2(wne, wne), w2 = make_float2(wsw, wsw), w3 = make_float2(wse, wse);
  float2 lo = fmul2(make_float2(nw.x, nw.y), w0), hi = fmul2(make_float2(nw.z, nw.w), w0);
  lo = ffma2(make_float2(ne.x, ne.y), w1, lo); hi = ffma2(make_float2(ne.z, ne.w), w1, hi);
  lo = ffma2(make_float2(sw.x, sw.y), w2, lo); hi = ffma2(make_float2(sw.z, sw.w), w2, hi);
  lo = ffma2(make_float2(se.x, se.y), w3, lo); hi = ffma2(make_float2(se.z, se.w), w3, hi);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// STAGE: the 12-byte-per-row outputs (unit direction, colour) leave through shared memory as fully coalesced 4-byte stores instead of
// three stride-12 scalar stores per row (each of those touches three 128-byte lines per warp).
template <bool PACKED, int UNR, bool STAGE = false>
__global__ void __launch_bounds__(kPgThreads, 5) project_gather_kernel(const PgParams p) {
  __shared__ PgRec rec[kPgThreads];
  __shared__ float s_dir[STAGE ? 3 * kPgThreads : 1], s_rgb[STAGE ? 3 * kPgThreads : 1];
  const int tid = threadIdx.x;
  const int kPgPoints = kPgThreads / p.rfn;                // points per tile: all 256 threads own a (view, point) row
  const int rows = kPgPoints * p.rfn;
  for (long long tile = blockIdx.x; tile * kPgPoints < p.pn; tile += gridDim.x) {
    const long long p0 = tile * kPgPoints;
    // ---- phase 1: thread = (view, point): w2c, spherical, pixel, direction, three footprints
    if (tid < rows) {
      const int v = (tid >= kPgPoints) + (tid >= 2 * kPgPoints) + (tid >= 3 * kPgPoints);
      const long long pi = p0 + (tid - v * kPgPoints);
      PgRec r;
      r.dxy = 0; r.off_rf = r.off_if = r.off_im = 0;
      r.tx_rf = r.ty_rf = r.tx_if = r.ty_if = r.tx_im = r.ty_im = 0.f;
      if (pi < p.pn) {
        const float x = __ldg(p.pts + 3 * pi), y = __ldg(p.pts + 3 * pi + 1), z = __ldg(p.pts + 3 * pi + 2);
        const float* w = p.w2c + 12 * v;
        const float c0 = w[0] * x + w[1] * y + w[2] * z + w[3];
        const float c1 = w[4] * x + w[5] * y + w[6] * z + w[7];
        const float c2 = w[8] * x + w[9] * y + w[10] * z + w[11];
        float radius, px, py;
        cam_to_equi(p.dataset, c0, c1, c2, p.H, p.W, radius, px, py);
        const float cam0 = -(w[0] * w[3] + w[4] * w[7] + w[8] * w[11]);
        const float cam1 = -(w[1] * w[3] + w[5] * w[7] + w[9] * w[11]);
        const float cam2 = -(w[2] * w[3] + w[6] * w[7] + w[10] * w[11]);
        const float e0 = x - cam0, e1 = y - cam1, e2 = z - cam2;
        const float en = fmaxf(sqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e-5f);
        const size_t row = (size_t)v * p.pn + pi;
        __stcs(reinterpret_cast<float2*>(p.out_pix) + row, make_float2(px, py));
        p.out_depth[row] = radius;
        if (STAGE) { s_dir[3 * tid] = -e0 / en; s_dir[3 * tid + 1] = -e1 / en; s_dir[3 * tid + 2] = -e2 / en; }
        else { p.out_dir[3 * row] = -e0 / en; p.out_dir[3 * row + 1] = -e1 / en; p.out_dir[3 * row + 2] = -e2 / en; }
        Footprint f = border_footprint(px, py, p.img_h, p.img_w, p.rf_h, p.rf_w);
        r.off_rf = v * p.rf_h * p.rf_w + f.off; r.dxy |= f.dx | (f.dy << 1); r.tx_rf = f.tx; r.ty_rf = f.ty;
        f = border_footprint(px, py, p.img_h, p.img_w, p.if_h, p.if_w);
        r.off_if = v * p.if_h * p.if_w + f.off; r.dxy |= (f.dx << 2) | (f.dy << 3); r.tx_if = f.tx; r.ty_if = f.ty;
        f = border_footprint(px, py, p.img_h, p.img_w, p.img_h, p.img_w);
        r.off_im = v * p.img_h * p.img_w + f.off; r.dxy |= (f.dx << 4) | (f.dy << 5); r.tx_im = f.tx; r.ty_im = f.ty;
        // colours: one 16-byte tap each, done by the row's own thread
        const float4 c = blend4(reinterpret_cast<const float4*>(p.imgs_cl) + r.off_im, (r.dxy >> 4) & 1, ((r.dxy >> 5) & 1) * p.img_w,
                                r.tx_im, r.ty_im);
        if (STAGE) { s_rgb[3 * tid] = c.x; s_rgb[3 * tid + 1] = c.y; s_rgb[3 * tid + 2] = c.z; }
        else { p.out_rgb[3 * row] = c.x; p.out_rgb[3 * row + 1] = c.y; p.out_rgb[3 * row + 2] = c.z; }
      }
      rec[tid] = r;
    }
    __syncthreads();
    if (STAGE) {
      const long long left = p.pn - p0;
      const int n3 = 3 * (int)(left < kPgPoints ? left : kPgPoints);
      for (int v = 0; v < p.rfn; ++v) {
        const size_t base = 3 * ((size_t)v * p.pn + p0);
        for (int j = tid; j < n3; j += kPgThreads) {
          __stcs(p.out_dir + base + j, s_dir[3 * v * kPgPoints + j]);
          __stcs(p.out_rgb + base + j, s_rgb[3 * v * kPgPoints + j]);
        }
      }
    }
    // ---- phase 2: lane = (row, float4 channel group): 8 lanes fetch one 128-byte texel line per tap and write one
    //      128-byte output row (coalesced both ways)
#pragma unroll UNR
    for (int it = tid; it < rows * 8; it += kPgThreads) {
      const int rrow = it >> 3, cg = it & 7;
      // rfn <= 4: the view of a row by comparisons (an integer division by the run-time tile size costs ~35 instructions per lane)
      const int v = (rrow >= kPgPoints) + (rrow >= 2 * kPgPoints) + (rrow >= 3 * kPgPoints);
      const long long pi = p0 + (rrow - v * kPgPoints);
      if (pi >= p.pn) continue;
      const PgRec r = rec[rrow];
      const size_t row = (size_t)v * p.pn + pi;
      const float4 a = blend4<PACKED>(reinterpret_cast<const float4*>(p.ray_feats_cl) + (size_t)r.off_rf * 8 + cg, (r.dxy & 1) * 8,
                              ((r.dxy >> 1) & 1) * p.rf_w * 8, r.tx_rf, r.ty_rf);
      __stcs(reinterpret_cast<float4*>(p.out_rf) + row * 8 + cg, a);
      if (p.out_if) {
        const float4 b = blend4<PACKED>(reinterpret_cast<const float4*>(p.img_feats_cl) + (size_t)r.off_if * 8 + cg, ((r.dxy >> 2) & 1) * 8,
                                ((r.dxy >> 3) & 1) * p.if_w * 8, r.tx_if, r.ty_if);
        __stcs(reinterpret_cast<float4*>(p.out_if) + row * 8 + cg, b);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K4: compositing, one warp per ray (sequential fp32 cumprod = the stated accumulation order)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) composite_kernel(const float* __restrict__ density, const float* __restrict__ alpha_in,
                                                        const float* __restrict__ colors, const float* __restrict__ depth,
                                                        int depth_ray_stride, float* __restrict__ hit_prob,
                                                        float* __restrict__ pixel_colors, float* __restrict__ render_depth,
                                                        int rn, int dn) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* alpha = sm + warp * 2 * dn;
  float* hit = alpha + dn;
  for (long long ray = (long long)blockIdx.x * 8 + warp; ray < rn; ray += (long long)gridDim.x * 8) {
    for (int s = lane; s < dn; s += 32)
      alpha[s] = alpha_in ? __ldg(alpha_in + ray * dn + s) : 1.f - expf(-fmaxf(__ldg(density + ray * dn + s), 0.f));
    __syncwarp();
    if (lane == 0) {
      float trans = 1.f;
      for (int s = 0; s < dn; ++s) { hit[s] = alpha[s] * trans; trans = trans * (1.f - alpha[s] + 1e-10f); }
    }
    __syncwarp();
    float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
    for (int s = lane; s < dn; s += 32) {
      const float h = hit[s];
      if (hit_prob) hit_prob[ray * dn + s] = h;
      if (colors) {
        const float* c = colors + (ray * dn + s) * 3;
        cr = fmaf(h, __ldg(c), cr); cg = fmaf(h, __ldg(c + 1), cg); cb = fmaf(h, __ldg(c + 2), cb);
      }
      if (depth) cd = fmaf(h, __ldg(depth + ray * depth_ray_stride + s), cd);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cr += __shfl_xor_sync(0xffffffffu, cr, o); cg += __shfl_xor_sync(0xffffffffu, cg, o);
      cb += __shfl_xor_sync(0xffffffffu, cb, o); cd += __shfl_xor_sync(0xffffffffu, cd, o);
    }
    if (lane == 0) {
      if (pixel_colors && colors) { pixel_colors[ray * 3] = cr; pixel_colors[ray * 3 + 1] = cg; pixel_colors[ray * 3 + 2] = cb; }
      if (render_depth && depth) render_depth[ray] = cd;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// K4b: inverse-CDF fine resampling, one warp per ray
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fine_sample_kernel(const float* __restrict__ depth, int depth_ray_stride,
                                                          const float* __restrict__ hit_prob, const float* __restrict__ u_table,
                                                          float near_d, float far_d, int inv_mode, int rn, int dn, int fdn,
                                                          int sort_out, int use_all, float* __restrict__ fine_out,
                                                          int* __restrict__ inds_out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = 2 * (dn + 1) + (fdn + dn);
  float* cdf = sm + warp * per;
  float* center = cdf + dn + 1;
  float* fine = center + dn + 1;
  const float nn = -1.f / near_d, ff = -1.f / far_d;
  for (long long ray = (long long)blockIdx.x * 8 + warp; ray < rn; ray += (long long)gridDim.x * 8) {
    const float* dp = depth + ray * depth_ray_stride;
    const float* hp = hit_prob + ray * dn;
    for (int s = lane; s <= dn; s += 32) {
      float d1 = __ldg(dp + min(s, dn - 1)), d0 = __ldg(dp + max(s - 1, 0));
      if (inv_mode) { d1 = (-1.f / d1 - nn) / (ff - nn); d0 = (-1.f / d0 - nn) / (ff - nn); }
      center[s] = (s == 0 || s == dn) ? d1 : (d1 + d0) / 2.f;
    }
    float tot = 0.f;
    if (lane == 0) for (int s = 0; s < dn; ++s) tot += __ldg(hp + s) + 1e-5f;
    tot = __shfl_sync(0xffffffffu, tot, 0);
    for (int s = lane; s < dn; s += 32) cdf[s + 1] = (__ldg(hp + s) + 1e-5f) / tot;
    __syncwarp();
    if (lane == 0) {
      float c = 0.f;
      cdf[0] = 0.f;
      for (int s = 0; s < dn; ++s) { c += cdf[s + 1]; cdf[s + 1] = c; }
    }
    __syncwarp();
    for (int k = lane; k < fdn; k += 32) {
      const float u = __ldg(u_table + k);
      int lo = 0, hi = dn + 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
      if (inds_out) inds_out[ray * fdn + k] = lo;
      const int below = max(lo - 1, 0), above = min(dn, lo);
      const float cb = cdf[below], ca = cdf[above];
      float denom = ca - cb;
      if (denom < 1e-5f) denom = 1.f;
      const float t = (u - cb) / denom;
      float fd = __fadd_rn(center[below], __fmul_rn(t, center[above] - center[below]));
      if (inv_mode) { fd = __fadd_rn(__fmul_rn(fd, ff - nn), nn); fd = -1.f / fd; }
      fine[k] = fd;
    }
    int total = fdn;
    if (use_all) { for (int s = lane; s < dn; s += 32) fine[fdn + s] = __ldg(dp + s); total = fdn + dn; }
    __syncwarp();
    if (sort_out) {
      for (int k = lane; k < total; k += 32) {
        const float x = fine[k];
        int rank = 0;
        for (int j = 0; j < total; ++j) { const float y = fine[j]; rank += (y < x || (y == x && j < k)) ? 1 : 0; }
        fine_out[ray * total + rank] = x;
      }
    } else {
      for (int k = lane; k < total; k += 32) fine_out[ray * total + k] = fine[k];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// sample_3sigma + sample_pdf (network/sample_utils.py:6-60, det = True) and the selection / sort of fine_render_impl
// (network/renderer.py:438-470): one warp per ray.  Rays whose marker ft[ray][0] >= min_valid get n depths between ft[ray][1] and
// ft[ray][2] (bin edges clamped to [near, far], bins weighted by a unit Gaussian over +-3 sigma, inverse CDF at u = t_table);
// the others keep the row they have.  t_table = linspace(0,1,n), gauss = 1/sqrt(2 pi) exp(-x^2/2) at linspace(-3,3,n-1): built by the
// host with the reference's torch ops.  Accumulation order of the normaliser and the cdf: sequential fp32 (the stated order).
// io (rn, n [+ dn]): with `coarse` the row becomes sort(coarse depths ++ samples), else sort(samples); `select` = 0 writes the
// unsorted samples of EVERY ray (the functional drop-in).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_3sigma_kernel(const float* __restrict__ ft, int ft_stride, float min_valid,
                                                            const float* __restrict__ t_table, const float* __restrict__ gauss, int n,
                                                            float near_d, float far_d, const float* __restrict__ coarse,
                                                            int coarse_stride, int dn, int select, int rn, float* __restrict__ io) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = n + (coarse ? dn : 0);
  float* edges = sm + warp * (2 * n + total);
  float* cdf = edges + n;
  float* z = cdf + n;
  for (long long ray = (long long)blockIdx.x * 8 + warp; ray < rn; ray += (long long)gridDim.x * 8) {
    const float* f = ft + ray * ft_stride;
    if (select && !(__ldg(f) >= min_valid)) continue;                       // warp-uniform
    const float low = __ldg(f + (select ? 1 : 0)), high = __ldg(f + (select ? 2 : 1));
    const float step = (high - low) / (float)(n - 1);
    for (int s = lane; s < n; s += 32) {
      const float t = __ldg(t_table + s);
      const float e = __fadd_rn(__fmul_rn(low, 1.f - t), __fmul_rn(high, t));
      edges[s] = fminf(fmaxf(e, near_d), far_d);
    }
    __syncwarp();
    for (int s = lane; s < n - 1; s += 32) cdf[s + 1] = __fadd_rn(__fmul_rn((edges[s + 1] - edges[s]) / step, __ldg(gauss + s)), 1e-5f);
    __syncwarp();
    if (lane == 0) {
      float tot = 0.f;
      for (int s = 1; s < n; ++s) tot += cdf[s];
      float c = 0.f;
      cdf[0] = 0.f;
      for (int s = 1; s < n; ++s) { c += cdf[s] / tot; cdf[s] = c; }
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {
      const float u = __ldg(t_table + k);
      int lo = 0, hi = n;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
      const int below = max(lo - 1, 0), above = min(n - 1, lo);
      const float cb = cdf[below], ca = cdf[above];
      float denom = ca - cb;
      if (denom < 1e-5f) denom = 1.f;
      const float t = (u - cb) / denom;
      z[k] = __fadd_rn(edges[below], __fmul_rn(t, edges[above] - edges[below]));
    }
    if (coarse) for (int s = lane; s < dn; s += 32) z[n + s] = __ldg(coarse + ray * coarse_stride + s);
    __syncwarp();
    float* out = io + ray * total;
    if (select) {
      for (int k = lane; k < total; k += 32) {
        const float x = z[k];
        int rank = 0;
        for (int j = 0; j < total; ++j) { const float y = z[j]; rank += (y < x || (y == x && j < k)) ? 1 : 0; }
        out[rank] = x;
      }
    } else {
      for (int k = lane; k < total; k += 32) out[k] = z[k];
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// a5: depth hypotheses (pipeline3_model.py:717-733, 774-815).  thread = pixel.
//   mono-guided list: clamp(mu + s * k_i, min, max), i < n_mono, with
//       mode 0  s * k_i = k_table[i]                      ("fixed_sigma": the host passes float32(k_i * fixed_sigma))
//       mode 1  s = max(sigma, basic_sigma), product rounded, then added   (mono_uncertainty, :731)
//       mode 2  (sigma * k_i) * relaxation_factor                          (:729)
//   centres: mode 0 a table shared by all pixels (linear :802 / inverse-linear :804, built by the host with the reference's torch
//       ops), mode 1 per-pixel `revise_range` (:784-799: [d_min, d_min + interval * j], j < n-1, interval = (d_max - d_min) / (n-1)),
//       mode 2 none (`wo_hdh`).
//   output = the two ascending lists merged (== torch.sort of their concatenation, values only); with `sort_out` = 0 (wo_hdh) the
//   mono list is written in k order, as the reference leaves it.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxMono = 16;
__global__ void __launch_bounds__(256) depth_hypotheses_kernel(const float* __restrict__ ref_mu, const float* __restrict__ ref_sigma,
                                                               const float* __restrict__ k_table, int n_mono, int mono_mode,
                                                               float basic_sigma, float relax, const float* __restrict__ centers,
                                                               int n_cen, int cen_mode, float fixed_dist, float min_d, float max_d,
                                                               int sort_out, long long hw, long long total_px,
                                                               float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_px) return;
  const long long b = i / hw, px = i % hw;
  const float mu = __ldg(ref_mu + i);
  const int D = n_mono + (cen_mode == 2 ? 0 : n_cen);
  float* o = out + b * D * hw + px;
  float mono[kMaxMono];
  const float sg = ref_sigma ? __ldg(ref_sigma + i) : 0.f;
#pragma unroll
  for (int j = 0; j < kMaxMono; ++j) {
    if (j >= n_mono) { mono[j] = INFINITY; continue; }
    const float k = __ldg(k_table + j);
    float add;
    if (mono_mode == 0) add = k;
    else if (mono_mode == 1) add = __fmul_rn(fmaxf(sg, basic_sigma), k);
    else add = __fmul_rn(__fmul_rn(sg, k), relax);
    mono[j] = fminf(fmaxf(__fadd_rn(mu, add), min_d), max_d);
  }
  if (!sort_out) {
    for (int j = 0; j < n_mono; ++j) o[(long long)j * hw] = mono[j];
    return;
  }
  // ascending k and a positive sigma give an ascending list; a predicted sigma may be negative (mode 2): insertion sort
#pragma unroll
  for (int a2 = 1; a2 < kMaxMono; ++a2) {
#pragma unroll
    for (int c = a2; c > 0; --c) {
      const float lo = fminf(mono[c - 1], mono[c]), hi = fmaxf(mono[c - 1], mono[c]);
      mono[c - 1] = lo; mono[c] = hi;
    }
  }
  float dmin = 0.f, interval = 0.f;
  if (cen_mode == 1) {
    dmin = fmaxf(__fsub_rn(mu, fixed_dist), min_d);
    const float dmax = fminf(__fadd_rn(mu, fixed_dist), max_d);
    interval = __fdiv_rn(__fsub_rn(dmax, dmin), (float)(n_cen - 1));
  }
  auto centre = [&](int j) -> float {
    if (cen_mode == 0) return __ldg(centers + j);
    return j == 0 ? dmin : __fadd_rn(dmin, __fmul_rn(interval, (float)(j - 1)));
  };
  int im = 0, il = 0;
  const int n_lin = cen_mode == 2 ? 0 : n_cen;
  float vm = mono[0];
  for (int d = 0; d < D; ++d) {
    const float vl = il < n_lin ? centre(il) : INFINITY;
    float v;
    if (vm <= vl) {
      v = vm; ++im;
      vm = INFINITY;
#pragma unroll
      for (int j = 1; j < kMaxMono; ++j) if (j == im) vm = mono[j];
    } else { v = vl; ++il; }
    o[(long long)d * hw] = v;
  }
}

}  // namespace pgrf

using namespace pgrf;
namespace pgrf { int g_pg_variant = 5; int g_pg_grid = 64; }   // B200 sweep (tools/time_pg_variants.py): staged 12-byte outputs, 8 gathers in flight per lane, 64 CTAs per SM-slot

extern "C" int pgrf_project_gather_fwd(const float* pts, long long pn, const float* w2c, int rfn, int dataset, int H, int W,
                                       const float* imgs_cl, int img_h, int img_w, const float* img_feats_cl, int if_h, int if_w,
                                       const float* ray_feats_cl, int rf_h, int rf_w, float* out_pix, float* out_depth,
                                       float* out_dir, float* out_ray_feats, float* out_rgb, float* out_img_feats, void* stream) {
  PGRF_REQUIRE(pts && w2c && imgs_cl && ray_feats_cl && out_pix && out_depth && out_dir && out_ray_feats && out_rgb,
               "project_gather: null pointer argument");
  PGRF_REQUIRE(!out_img_feats || img_feats_cl, "project_gather: out_img_feats needs img_feats_cl");
  PGRF_REQUIRE(pn >= 1 && rfn >= 1 && rfn <= 4, "project_gather: pn=%lld rfn=%d (1..4 views)", pn, rfn);
  PGRF_REQUIRE(dataset >= 0 && dataset <= 3, "project_gather: unknown dataset id %d", dataset);
  PgParams p;
  p.pts = pts; p.w2c = w2c; p.imgs_cl = imgs_cl; p.img_feats_cl = img_feats_cl; p.ray_feats_cl = ray_feats_cl;
  p.out_pix = out_pix; p.out_depth = out_depth; p.out_dir = out_dir; p.out_rf = out_ray_feats; p.out_rgb = out_rgb;
  p.out_if = out_img_feats;
  p.pn = pn; p.rfn = rfn; p.dataset = dataset; p.H = H; p.W = W; p.img_h = img_h; p.img_w = img_w;
  p.if_h = if_h; p.if_w = if_w; p.rf_h = rf_h; p.rf_w = rf_w;
  const int pts_per_tile = kPgThreads / rfn;
  const long long tiles = (pn + pts_per_tile - 1) / pts_per_tile;
  const int grid = (int)(tiles < 148 * g_pg_grid ? tiles : 148 * g_pg_grid);
  switch (g_pg_variant) {
    case 0: project_gather_kernel<false, 2><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    case 1: project_gather_kernel<true, 2><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    case 2: project_gather_kernel<true, 4><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    case 3: project_gather_kernel<false, 4><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    case 4: project_gather_kernel<false, 8><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    case 5: project_gather_kernel<false, 4, true><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
    default: project_gather_kernel<true, 4, true><<<grid, kPgThreads, 0, (cudaStream_t)stream>>>(p); break;
  }
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_composite_fwd(const float* density, const float* alpha, const float* colors, const float* depth,
                                  int depth_ray_stride, int rn, int dn, float* hit_prob, float* pixel_colors,
                                  float* render_depth, void* stream) {
  PGRF_REQUIRE((density != nullptr) != (alpha != nullptr), "composite: pass exactly one of density / alpha");
  // 8 rays per CTA x 2 vectors of dn floats in shared memory: opt in above 48 KB, reject what no SM can hold (227 KB)
  const size_t smem = 8 * 2 * (size_t)dn * sizeof(float);
  PGRF_REQUIRE(rn >= 1 && dn >= 1 && smem <= 200 * 1024, "composite: rn=%d dn=%d (dn <= 3200)", rn, dn);
  if (smem > 48 * 1024) PGRF_CUDA(cudaFuncSetAttribute(composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (rn + 7) / 8 < 148 * 8 ? (rn + 7) / 8 : 148 * 8;
  composite_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(density, alpha, colors, depth, depth_ray_stride,
                                                                                 hit_prob, pixel_colors, render_depth, rn, dn);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_fine_sample_fwd(const float* depth, int depth_ray_stride, const float* hit_prob, const float* u_table,
                                    float near_depth, float far_depth, int inv_mode, int rn, int dn, int fine_dn, int sort_out,
                                    int use_all, float* fine_out, int* inds_out, void* stream) {
  PGRF_REQUIRE(depth && hit_prob && u_table && fine_out, "fine_sample: null pointer argument");
  PGRF_REQUIRE(rn >= 1 && dn >= 2 && fine_dn >= 1 && dn <= 1024 && fine_dn <= 1024, "fine_sample: rn=%d dn=%d fine_dn=%d", rn, dn, fine_dn);
  const size_t smem = 8 * (size_t)(2 * (dn + 1) + fine_dn + dn) * sizeof(float);
  PGRF_REQUIRE(smem <= 200 * 1024, "fine_sample: dn/fine_dn too large for shared memory");
  if (smem > 48 * 1024) PGRF_CUDA(cudaFuncSetAttribute(fine_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (rn + 7) / 8 < 148 * 8 ? (rn + 7) / 8 : 148 * 8;
  fine_sample_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(depth, depth_ray_stride, hit_prob, u_table, near_depth, far_depth,
                                                               inv_mode, rn, dn, fine_dn, sort_out, use_all, fine_out, inds_out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_sample_3sigma_fwd(const float* ft, int ft_stride, float min_valid, const float* t_table, const float* gauss, int n,
                                      float near_depth, float far_depth, const float* coarse_depth, int coarse_ray_stride, int dn,
                                      int select, int rn, float* io, void* stream) {
  PGRF_REQUIRE(ft && t_table && gauss && io, "sample_3sigma: null pointer argument");
  PGRF_REQUIRE(rn >= 1 && n >= 2 && n <= 1024 && dn >= 0 && dn <= 1024 && ft_stride >= (select ? 3 : 2), "sample_3sigma: rn=%d n=%d dn=%d", rn, n, dn);
  PGRF_REQUIRE(select || !coarse_depth, "sample_3sigma: the coarse depths are merged in selection mode only");
  const size_t smem = 8 * (size_t)(3 * n + (coarse_depth ? dn : 0)) * sizeof(float);
  PGRF_REQUIRE(smem <= 200 * 1024, "sample_3sigma: n / dn too large for shared memory");
  if (smem > 48 * 1024) PGRF_CUDA(cudaFuncSetAttribute(sample_3sigma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (rn + 7) / 8 < 148 * 8 ? (rn + 7) / 8 : 148 * 8;
  sample_3sigma_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(ft, ft_stride, min_valid, t_table, gauss, n, near_depth, far_depth,
                                                                 coarse_depth, coarse_ray_stride, dn, select, rn, io);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_depth_hypotheses2_fwd(const float* ref_mu, const float* ref_sigma, int B, int h, int w, const float* k_table,
                                          int n_mono, int mono_mode, float basic_sigma, float relaxation, const float* centers,
                                          int n_centers, int centers_mode, float fixed_dist, float min_depth, float max_depth,
                                          int sort_out, float* out, void* stream) {
  PGRF_REQUIRE(ref_mu && out && (n_mono == 0 || k_table), "depth_hypotheses: null pointer argument");
  PGRF_REQUIRE(B >= 1 && h >= 1 && w >= 1 && n_mono >= 0 && n_mono <= kMaxMono && n_centers >= 0, "depth_hypotheses: bad sizes");
  PGRF_REQUIRE(mono_mode >= 0 && mono_mode <= 2 && (mono_mode == 0 || ref_sigma || n_mono == 0), "depth_hypotheses: sigma modes need ref_sigma");
  PGRF_REQUIRE(centers_mode >= 0 && centers_mode <= 2 && (centers_mode != 0 || centers || n_centers == 0) &&
                   (centers_mode != 1 || n_centers >= 2), "depth_hypotheses: bad centres");
  PGRF_REQUIRE(n_mono + (centers_mode == 2 ? 0 : n_centers) >= 1, "depth_hypotheses: empty output");
  const long long hw = (long long)h * w, total = hw * B;
  depth_hypotheses_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      ref_mu, ref_sigma, k_table, n_mono, mono_mode, basic_sigma, relaxation, centers, n_centers, centers_mode, fixed_dist, min_depth,
      max_depth, sort_out, hw, total, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_depth_hypotheses_fwd(const float* ref_mu, int B, int h, int w, const float* k_sigma, int n_mono, const float* linear,
                                         int n_linear, float min_depth, float max_depth, float* out, void* stream) {
  PGRF_REQUIRE(ref_mu && k_sigma && linear && out, "depth_hypotheses: null pointer argument");
  return pgrf_depth_hypotheses2_fwd(ref_mu, nullptr, B, h, w, k_sigma, n_mono, 0, 0.f, 1.f, linear, n_linear, 0, 0.f, min_depth, max_depth, 1,
                                    out, stream);
}

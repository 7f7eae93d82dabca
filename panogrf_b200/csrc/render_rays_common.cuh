// Per-ray tail of the bf16 rays kernels (render_rays16.cu, render_rays_tc.cu), one warp per ray: alpha compositing
// (renderer.py:210-219, render_ops.py:145-153, renderer.py:302-304) and, for the coarse pass, the inverse-CDF fine resampling
// (render_ops.py:413-473) + sort (renderer.py:472).  Scans are warp scans over per-lane segments (the fp32 parity path keeps the
// sequential order).
#pragma once
#include "render_device.cuh"

namespace pgrf {

constexpr int RRV = 256;     // floats per per-ray scratch vector (rows of a tile x 2 for the cdf / fine buffers)
enum { RV2_SIGMA = 0, RV2_ALPHA, RV2_HIT, RV2_CDF, RV2_CENTER, RV2_FINE };

__device__ __forceinline__ float warp_sum16(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// `r` = index of the ray inside the tile (its samples are rows r*dn .. r*dn+dn-1), `RV` = 6 x [256] floats, `RGB` = 3 x [rrows]
__device__ __forceinline__ void composite_ray(const pgrf_render_args& a, long long ray, int r, int dn, int lane, float* RV,
                                              const float* RGB, int rrows) {
  const int RROWS = rrows;
  const int m0 = r * dn;
  float* alpha = RV + RV2_ALPHA * RRV + m0;
  float* hit = RV + RV2_HIT * RRV + m0;
  const float* sigma = RV + RV2_SIGMA * RRV + m0;
  const float* dp = a.depth + (size_t)ray * a.depth_ray_stride;
  for (int s = lane; s < dn; s += 32) alpha[s] = 1.f - fast_exp(-sigma[s]);
  __syncwarp();
  {   // transmittance = exclusive prefix product of (1 - alpha + 1e-10): warp scan over contiguous per-lane segments
      // (the fp32 parity path keeps the sequential order; here only the association of the products differs)
    const int epl = (dn + 31) >> 5;                    // <= 4 elements per lane
    const int b0 = lane * epl;
    float loc[4], p = 1.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      loc[e] = p;
      if (e < epl && b0 + e < dn) p *= 1.f - alpha[b0 + e] + 1e-10f;
    }
    float incl = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= t; }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (e < epl && b0 + e < dn) hit[b0 + e] = alpha[b0 + e] * (excl * loc[e]);
  }
  __syncwarp();
  float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
  for (int s = lane; s < dn; s += 32) {
    const float hs = hit[s];
    const float r_ = RGB[0 * RROWS + m0 + s], g_ = RGB[1 * RROWS + m0 + s], b_ = RGB[2 * RROWS + m0 + s];
    cr = fmaf(hs, r_, cr); cg = fmaf(hs, g_, cg); cb = fmaf(hs, b_, cb);
    cd = fmaf(hs, __ldg(dp + s), cd);
    if (a.hit_prob) a.hit_prob[(size_t)ray * dn + s] = hs;
    if (a.density) a.density[(size_t)ray * dn + s] = sigma[s];
    if (a.colors) {
      float* c = a.colors + ((size_t)ray * dn + s) * 3;
      c[0] = r_; c[1] = g_; c[2] = b_;
    }
  }
  cr = warp_sum16(cr); cg = warp_sum16(cg); cb = warp_sum16(cb); cd = warp_sum16(cd);
  if (lane == 0) {
    a.pixel_colors[(size_t)ray * 3 + 0] = cr; a.pixel_colors[(size_t)ray * 3 + 1] = cg; a.pixel_colors[(size_t)ray * 3 + 2] = cb;
    if (a.render_depth) a.render_depth[ray] = cd;
  }
  if (a.fine_depth) {
    // ---- sample_fine_depth (render_ops.py:413-473), deterministic u-table ----
    float* cdf = RV + RV2_CDF * RRV + 2 * m0;
    float* center = RV + RV2_CENTER * RRV + 2 * m0;
    float* fine = RV + RV2_FINE * RRV + 2 * m0;
    const bool inv = a.use_disp != 0;
    const float nn = -1.f / a.que_near, ff = -1.f / a.que_far;
    for (int s = lane; s <= dn; s += 32) {
      float d1 = __ldg(dp + min(s, dn - 1));
      float d0 = __ldg(dp + max(s - 1, 0));
      if (inv) { d1 = (-1.f / d1 - nn) / (ff - nn); d0 = (-1.f / d0 - nn) / (ff - nn); }
      center[s] = (s == 0 || s == dn) ? d1 : (d1 + d0) / 2.f;
    }
    {   // pdf = (hit + 1e-5) / sum, cdf = inclusive prefix sum: warp reduce + warp scan over per-lane segments
      const int epl = (dn + 31) >> 5;
      const int b0 = lane * epl;
      float w[4], part = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { w[e] = (e < epl && b0 + e < dn) ? hit[b0 + e] + 1e-5f : 0.f; part += w[e]; }
      float tot = part;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
      const float itot = 1.f / tot;
      float run = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { w[e] *= itot; run += w[e]; w[e] = run; }   // inclusive inside the lane
      float incl = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) { excl = 0.f; cdf[0] = 0.f; }
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (e < epl && b0 + e < dn) cdf[b0 + e + 1] = excl + w[e];
    }
    __syncwarp();
    const int fdn = a.fine_dn;
    for (int k = lane; k < fdn; k += 32) {
      const float u = __ldg(a.fine_u + k);
      int lo = 0, hi = dn + 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
      const int inds = lo;
      if (a.fine_inds) a.fine_inds[(size_t)ray * fdn + k] = inds;
      const int below = max(inds - 1, 0), above = min(dn, inds);
      const float cb_ = cdf[below], ca_ = cdf[above];
      float denom = ca_ - cb_;
      if (denom < 1e-5f) denom = 1.f;
      const float t = (u - cb_) / denom;
      float fd = __fadd_rn(center[below], __fmul_rn(t, center[above] - center[below]));
      if (inv) { fd = __fadd_rn(__fmul_rn(fd, ff - nn), nn); fd = -1.f / fd; }
      fine[k] = fd;
    }
    int total_out = fdn;
    if (a.fine_use_all) {
      for (int s = lane; s < dn; s += 32) fine[fdn + s] = __ldg(dp + s);
      total_out = fdn + dn;
    }
    __syncwarp();
    bool sorted = !a.fine_use_all;                 // inverse-CDF samples of increasing u are almost always already ordered
    if (sorted)
      for (int k = lane; k + 1 < total_out; k += 32) sorted = sorted && (fine[k] <= fine[k + 1]);
    if (__all_sync(0xffffffffu, sorted)) {
      for (int k = lane; k < total_out; k += 32) a.fine_depth[(size_t)ray * total_out + k] = fine[k];
    } else
    for (int k = lane; k < total_out; k += 32) {   // rank sort (value-only result == torch.sort)
      const float x = fine[k];
      int rank = 0;
      for (int j = 0; j < total_out; ++j) {
        const float y = fine[j];
        rank += (y < x || (y == x && j < k)) ? 1 : 0;
      }
      a.fine_depth[(size_t)ray * total_out + rank] = x;
    }
  }
}

}  // namespace pgrf

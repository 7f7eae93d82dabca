// Depth-prior sample placement of the render path ("diner" branch, network/renderer.py:570-600 and :318-355):
//   project_points_dict_diner                     network/render_ops.py:260-290
//   sample_depthguided / fill_up_uniform_samples  network/original_depth_guided_sample.py:45-297 / 333-366
// One WARP per ray (4 independent rays per CTA, no block-wide barrier anywhere) walks the ray's candidate depths (typically 1000): project into every source panorama, gather the
// MVS depth / variance / normal priors (bilinear, border padding), surface likelihood = max over views, compact the
// few candidates with non-zero likelihood, keep the most likely ones (shared-memory bitonic sort only when more survive
// than there are slots), draw the Gaussian samples, fill empty slots, append the uniform samples, sort, write.
// Nothing but the (rn, n_samples) result reaches HBM; the reference materialises (rfn, rn, n_candidates, 7) floats.
#include "render_device.cuh"

namespace pgrf {

constexpr int kDgWarps = 4;                // rays per CTA (one warp each)
constexpr int kDgThreads = 32 * kDgWarps;
constexpr int kDgMaxCand = 4096;
constexpr int kDgMaxOut = 512;
constexpr int kDgQueue = 256;             // (candidate, view) pairs waiting for the exact evaluation, per warp

int g_dg_regsort = 1;     // 0 = final sorts in shared memory for every size (debug knob dg_regsort)
int g_dg_prefilter = 1;   // 0 = exact projection for every (candidate, view) pair (debug knob dg_prefilter)

__device__ __forceinline__ float tap1(const float* __restrict__ m, const Footprint& f, int fw) {
  const float nw = __ldg(m + f.off), ne = __ldg(m + f.off + f.dx);
  const float sw = __ldg(m + f.off + f.dy * fw), se = __ldg(m + f.off + f.dy * fw + f.dx);
  const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
  float o = nw * (tx1 * ty1);
  o = fmaf(ne, f.tx * ty1, o);
  o = fmaf(sw, tx1 * f.ty, o);
  o = fmaf(se, f.tx * f.ty, o);
  return o;
}

// original_depth_guided_sample.py:80-90,180-196 for one (view, candidate)
__device__ __forceinline__ float surface_likelihood(const pgrf_diner_args& a, float mu, float uncert, float pd, float cosv) {
  const bool ok = fabsf(mu - pd) < a.depth_diff_max && (!a.include_norm || cosv <= 0.f);
  if (!ok) return 0.f;
  const float sigma = a.diner_sigma > 0.f ? a.diner_sigma : (a.sigma_is_var ? sqrtf(uncert) : uncert);
  const float half = a.cand_step / 2.f;
  const float s2 = __fmul_rn(sigma, 1.41421356237309504880f);
  const float x1 = (pd + half - mu) / s2;
  const float x0 = (pd - half - mu) / s2;
  return 0.5f * fabsf(erff(x1) - erff(x0));
}

// ---- conservative pre-filter of phase 1 --------------------------------------------------------------------
// A (candidate, view) pair contributes only when |mu - pd| < depth_diff_max, mu = the bilinear prior depth at the projected
// pixel.  The exact projection (atan2f / acosf / IEEE divisions, ~380 instructions) is what the oracle is compared with bit for
// bit, and it is needed for the handful of candidates next to the prior surface only.  The filter projects with polynomial
// atan / acos (|error| < 2e-6 rad, checked in tests/test_diner_gpu.py through the bit-exact results) and is SOUND: when the
// approximate map coordinate is further than `eps` (4x the error bound + the fp32 rounding of the coordinate chain) from a
// texel boundary, the exact footprint covers the same four texels, the exact mu is a convex combination of them, and
// pd outside [min - thr, max + thr] (with slack for rounding) proves the pair contributes 0.  Everything else — next to a
// texel boundary, next to the prior surface, NaNs — is queued and evaluated exactly.
__device__ __forceinline__ float fast_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
struct DgFilter {
  float sx, bx, sy, by;        // angle -> map coordinate
  float eps_x, eps_y, eps_ys;  // safety margins in texels (eps_ys scales with 1/sin(phi): conditioning of acos)
  float xmax, ymax;
  int dataset;
};
__device__ __forceinline__ DgFilter dg_filter(const pgrf_diner_args& a) {
  DgFilter f;
  const bool align = (a.map_h == a.img_h && a.map_w == a.img_w);
  const float fw = (float)a.map_w, fh = (float)a.map_h;
  // pixel = angle / (2 pi | pi) * (W-1 | H-1) (cam_to_equi), map coordinate = pixel / (img-1) * (f-1 | f) [- 0.5] (border_footprint)
  f.sx = (align ? fw - 1.f : fw) * (1.f / PGRF_TWO_PI_F) * ((float)(a.W - 1) / (float)(a.img_w - 1));
  f.sy = (align ? fh - 1.f : fh) * (1.f / PGRF_PI_F) * ((float)(a.H - 1) / (float)(a.img_h - 1));
  f.bx = f.by = align ? 0.f : -0.5f;
  f.eps_x = fw * 2.5e-6f;
  f.eps_y = fh * 3.6e-6f;
  f.eps_ys = fh * (8e-7f / PGRF_PI_F);
  f.xmax = fw - 1.f;
  f.ymax = fh - 1.f;
  f.dataset = a.dataset;
  return f;
}
// c = camera-frame point (approximate: A + B t), dpos = bound of its distance to the exactly computed point
__device__ __forceinline__ bool dg_certainly_far(const DgFilter& f, const float* __restrict__ dmap, int fw, float thr, float c0, float c1,
                                                 float c2, float dpos) {
  const float r2 = c0 * c0 + c1 * c1 + c2 * c2;
  const float rpd = fast_rsqrt(r2);                 // r2 == 0 -> NaN below -> queued
  const float pd = r2 * rpd;
  // the four conventions of cam_to_equi (render_device.cuh) differ in the components that feed atan2 / acos and in how the angle
  // becomes a pixel: px = t' / (2 pi) * (W - 1), py = phi' / pi * (H - 1) with
  //   m3d          t = atan2(c2, c0)   t' = wrap(t + pi/2)                  phi' = acos(c1 / (r + 1e-5))
  //   replica      t = atan2(c0, c2)   t' = t + pi                          phi' = pi - acos(c1 / r)     (= asin + pi/2)
  //   residential  t = atan2(c2, c0)   t' = t + 3pi/2 (- 2pi if t > pi/2)   phi' = acos(c1 / r)          (= pi/2 - asin)
  //   CoffeeArea   t = atan2(c1, c0)   t' = 2pi - wrap(t)                   phi' = acos(c2 / r)
  // a branch taken differently than the exact code moves t' by 2 pi, i.e. the map coordinate from one end of the row to the
  // other: both ends are within the margin of a texel boundary (or clamped), so such pairs are queued like every boundary case.
  const int ds = f.dataset;
  const float num = ds == PGRF_DS_REPLICA_TEST ? c0 : (ds == PGRF_DS_COFFEEAREA ? c1 : c2);
  const float den = ds == PGRF_DS_REPLICA_TEST ? c2 : c0;
  const float pc = ds == PGRF_DS_COFFEEAREA ? c2 : c1;
  const float ax = fabsf(den), az = fabsf(num);
  const float rmx = fast_rcp(fmaxf(ax, az));
  const float q = fminf(ax, az) * rmx;
  const float s = q * q;
  float t = fmaf(s, 0.00782548263669014f, -0.03689862787723541f);
  t = fmaf(t, s, 0.08374155312776566f);
  t = fmaf(t, s, -0.13480405509471893f);
  t = fmaf(t, s, 0.19879871606826782f);
  t = fmaf(t, s, -0.3332637548446655f);
  t = fmaf(t, s, 0.9999993443489075f);
  t *= q;
  if (az > ax) t = PGRF_HALF_PI_F - t;
  if (den < 0.f) t = PGRF_PI_F - t;
  if (num < 0.f) t = -t;
  if (ds == PGRF_DS_M3D) { t += PGRF_HALF_PI_F; if (t < 0.f) t += PGRF_TWO_PI_F; }
  else if (ds == PGRF_DS_REPLICA_TEST) t += PGRF_PI_F;
  else if (ds == PGRF_DS_RESIDENTIAL) t += (t > PGRF_HALF_PI_F) ? -PGRF_HALF_PI_F : 4.71238898038468985769f;
  else t = (t < 0.f) ? -t : PGRF_TWO_PI_F - t;
  const float xq = pc * fast_rcp(pd + (ds == PGRF_DS_M3D ? 1e-5f : 0.f));
  const float aq = fabsf(xq), om = 1.f - aq;
  float ph = fmaf(aq, 0.002251368248835206f, -0.011012386530637741f);
  ph = fmaf(ph, aq, 0.02674933150410652f);
  ph = fmaf(ph, aq, -0.048724401742219925f);
  ph = fmaf(ph, aq, 0.08873733133077621f);
  ph = fmaf(ph, aq, -0.21458369493484497f);
  ph = fmaf(ph, aq, 1.5707961320877075f);
  const float rs = fast_rsqrt(om);                 // om <= 0 -> inf / NaN: the margin below is not finite and the pair is queued
  ph *= om * rs;
  if ((xq < 0.f) != (ds == PGRF_DS_REPLICA_TEST)) ph = PGRF_PI_F - ph;      // replica: phi' = pi - acos
  float ix = fmaf(t, f.sx, f.bx), iy = fmaf(ph, f.sy, f.by);
  ix = fminf(fmaxf(ix, 0.f), f.xmax);
  iy = fminf(fmaxf(iy, 0.f), f.ymax);
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float tx = ix - x0, ty = iy - y0;
  // margins: approximation + rounding (constant part), position uncertainty dpos seen under the angle's lever arm,
  // conditioning of acos (1 / sin(phi) = rs * rsqrt(1 + |x|))
  const float ex = fmaf(f.sx * dpos, rmx, f.eps_x);
  const float ey = fmaf(fmaf(2.f * f.sy * dpos, rpd, f.eps_ys), rs * fast_rsqrt(1.f + aq), f.eps_y);
  const bool inside = fminf(fminf(tx, 1.f - tx) - ex, fminf(ty, 1.f - ty) - ey) > 0.f;   // false on NaN: the pair is queued
  // branch-free: the four texels are read in any case (indices clamped for the read only), so that two candidates' chains interleave
  const int xi = min((int)x0, (int)f.xmax - 1), yi = min((int)y0, (int)f.ymax - 1);
  const int idx = max(yi, 0) * fw + max(xi, 0);
  const float nw = __ldg(dmap + idx), ne = __ldg(dmap + idx + 1), sw = __ldg(dmap + idx + fw), se = __ldg(dmap + idx + fw + 1);
  const float lo = fminf(fminf(nw, ne), fminf(sw, se)), hi = fmaxf(fmaxf(nw, ne), fmaxf(sw, se));
  const float slack = fmaf(1e-5f, pd + 2.f * fmaxf(fabsf(lo), fabsf(hi)), fmaf(2.f, dpos, thr * 1.00001f));   // thr + rounding slack
  return inside && ((pd - hi > slack) || (lo - pd > slack));      // false on NaN texels
}

template <typename T, typename Less>
__device__ __forceinline__ void bitonic_sort(T* a, int n, Less less, int lane) {   // n = power of two, one warp, `a` in shared memory
  for (int k = 2; k <= n; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < n; i += 32) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const T x = a[i], y = a[ixj];
          if (less(y, x) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncwarp();
    }
}

// Ascending bitonic sort of 32*E floats held one column per lane (element index = r*32 + lane): compare-exchange partners
// 32 or more apart live in the same lane, closer ones are one shuffle away.  Same exchange rule as bitonic_sort above.
template <int E>
__device__ __forceinline__ void warp_sort_regs(float (&x)[E], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          if ((r ^ jr) > r) {
            const bool up = (((r << 5) | lane) & k) == 0;
            const float lo = x[r], hi = x[r ^ jr];
            if ((hi < lo) == up) { x[r] = hi; x[r ^ jr] = lo; }
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const bool up = (((r << 5) | lane) & k) == 0;
          const float mine = x[r], other = __shfl_xor_sync(0xffffffffu, mine, j);
          const bool take = lower ? ((other < mine) == up) : ((mine < other) == up);
          x[r] = take ? other : mine;
        }
      }
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ int next_pow2(int n) { int p = 1; while (p < n) p <<= 1; return p; }

// Range of every prior depth map, [min, max] per view (NaNs ignored): lets phase 1 drop candidates by their distance to the source
// camera alone, before any angle is computed.  One CTA per view.
__global__ void __launch_bounds__(1024) depth_range_kernel(const float* __restrict__ maps, long long px, float* __restrict__ out) {
  __shared__ float s_lo[32], s_hi[32];
  const float* m = maps + (size_t)blockIdx.x * px;
  float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
  for (long long i = threadIdx.x; i < px; i += blockDim.x) { const float v = __ldg(m + i); lo = fminf(lo, v); hi = fmaxf(hi, v); }
  for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = s_lo[threadIdx.x]; hi = s_hi[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi; }
  }
}

// E > 0: the two final sorts run in registers (32*E >= n_samples + n_uniform slots); E == 0: in shared memory
template <int E>
__global__ void __launch_bounds__(kDgThreads) depth_guided_kernel(const pgrf_diner_args a, int nc_pad, int nc_pow2, int g_prefilter_on,
                                                                 const float* __restrict__ prior_range) {
  extern __shared__ __align__(16) unsigned char dsm[];
  const int tid = threadIdx.x & 31, warp = threadIdx.x >> 5;                             // tid = lane: the warp owns the ray
  const int nc = a.n_candidates, ns = a.n_samples, ng = a.n_gaussian, nu = a.n_uniform;
  const int keep = ns - ng;
  const int n_out = ns + nu;
  const int out_pow2 = next_pow2(n_out), ns_pow2 = next_pow2(ns);
  const bool from_dict = a.prj_mu != nullptr;
  // per-warp shared memory: keys [nc_pow2] | lik (later the opacity weights, in place) [nc_pad] | z [out_pow2]
  const size_t warp_bytes = (size_t)nc_pow2 * 8 + (size_t)nc_pad * 4 + (size_t)out_pow2 * 4 + kDgQueue * 4;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dsm + warp * warp_bytes);
  float* lik = reinterpret_cast<float*>(keys + nc_pow2);
  float* opq = lik;
  float* z = lik + nc_pad;
  int* queue = reinterpret_cast<int*>(z + out_pow2);
  const bool prefilter = !from_dict && g_prefilter_on && a.map_h >= 2 && a.map_w >= 2;
  const DgFilter flt = dg_filter(a);

  for (long long ray = (long long)blockIdx.x * kDgWarps + warp; ray < a.rn; ray += (long long)gridDim.x * kDgWarps) {
    int count = 0;                                                                       // survivors so far (warp-uniform)
    const float* cand = a.cand_depth + ray * a.cand_ray_stride;
    // ---- ray in world space (render_ops.py:76-106)
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (!from_dict) {
      const float cx = __ldg(a.coords + 2 * ray), cy = __ldg(a.coords + 2 * ray + 1);
      float dx, dy, dz;
      equi_unit_dir(a.dataset, (float)(long long)cx, (float)(long long)cy, a.H, a.W, dx, dy, dz);
      const float* c = a.que_c2w;
      r0 = c[0] * dx + c[1] * dy + c[2] * dz;
      r1 = c[4] * dx + c[5] * dy + c[6] * dz;
      r2 = c[8] * dx + c[9] * dy + c[10] * dz;
      o0 = c[3]; o1 = c[7]; o2 = c[11];
      const float rn = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
      u0 = r0 / rn; u1 = r1 / rn; u2 = r2 / rn;             // = -que_dir (original_depth_guided_sample.py:112-114)
    }
    const size_t map_px = (size_t)a.map_h * a.map_w;
    // exact likelihood of candidate i in view v (the code path the oracle is compared with)
    auto exact_lik = [&](int i, int v) -> float {
      const float t = __ldg(cand + i);
      const float p0 = o0 + r0 * t, p1 = o1 + r1 * t, p2 = o2 + r2 * t;
      const float* w = a.ref_w2c + 12 * v;
      const float pc0 = w[0] * p0 + w[1] * p1 + w[2] * p2 + w[3];
      const float pc1 = w[4] * p0 + w[5] * p1 + w[6] * p2 + w[7];
      const float pc2 = w[8] * p0 + w[9] * p1 + w[10] * p2 + w[11];
      float pd, px, py;
      cam_to_equi(a.dataset, pc0, pc1, pc2, a.H, a.W, pd, px, py);
      const Footprint f = border_footprint(px, py, a.img_h, a.img_w, a.map_h, a.map_w);
      const float mu = tap1(a.mvs_depth + v * map_px, f, a.map_w);
      if (!(fabsf(mu - pd) < a.depth_diff_max)) return 0.f;            // the common case: far from the prior surface
      const float uncert = tap1(a.mvs_uncert + v * map_px, f, a.map_w);
      float cosv = 0.f;
      if (a.include_norm) {
        const float d0 = w[0] * u0 + w[1] * u1 + w[2] * u2;
        const float d1 = w[4] * u0 + w[5] * u1 + w[6] * u2;
        const float d2 = w[8] * u0 + w[9] * u1 + w[10] * u2;
        const float* nm = a.mvs_normal + (size_t)v * 3 * map_px;
        cosv = d0 * tap1(nm, f, a.map_w) + d1 * tap1(nm + map_px, f, a.map_w) + d2 * tap1(nm + 2 * map_px, f, a.map_w);
      }
      return surface_likelihood(a, mu, uncert, pd, cosv);
    };
    if (prefilter) {
      // ---- phase 1 (filtered): reject with the approximate projection, queue the rest, evaluate the queue exactly
      for (int i = tid; i < nc; i += 32) lik[i] = 0.f;
      int n_q = 0;                                                                        // warp-uniform
      auto drain = [&]() {
        __syncwarp();
        for (int e = tid; e < n_q; e += 32) {
          const int code = queue[e];
          const float l = exact_lik(code & 0xFFFF, code >> 16);
          if (l > 0.f) atomicMax(reinterpret_cast<unsigned*>(lik) + (code & 0xFFFF), __float_as_uint(l));   // l > 0: bit order = value order
        }
        __syncwarp();
        n_q = 0;
      };
      for (int v = 0; v < a.rfn; ++v) {
        // camera-frame point of candidate depth t: (W o + T) + (W r) t, with a bound of its distance to the exact evaluation
        // (8 ulp of the largest intermediate magnitude of either evaluation order)
        const float* w = a.ref_w2c + 12 * v;
        const float A0 = w[0] * o0 + w[1] * o1 + w[2] * o2 + w[3], B0 = w[0] * r0 + w[1] * r1 + w[2] * r2;
        const float A1 = w[4] * o0 + w[5] * o1 + w[6] * o2 + w[7], B1 = w[4] * r0 + w[5] * r1 + w[6] * r2;
        const float A2 = w[8] * o0 + w[9] * o1 + w[10] * o2 + w[11], B2 = w[8] * r0 + w[9] * r1 + w[10] * r2;
        const float on = fabsf(o0) + fabsf(o1) + fabsf(o2), rnm = fabsf(r0) + fabsf(r1) + fabsf(r2);
        float magA = 0.f, magB = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float wr = fmaxf(fmaxf(fabsf(w[4 * k]), fabsf(w[4 * k + 1])), fabsf(w[4 * k + 2]));
          magA = fmaxf(magA, wr * on + fabsf(w[4 * k + 3]));
          magB = fmaxf(magB, wr * rnm);
        }
        magA *= 1e-6f; magB *= 1e-6f;
        const float* dmap = a.mvs_depth + v * map_px;
        // distance-only rejection: |mu - pd| >= thr for every mu of the view when pd is outside [min - thr, max + thr] (with slack for
        // rounding and the position bound); compared on squared distances.  An all-NaN map gives [+inf, -inf] -> NaN bounds -> nothing is dropped here.
        const float g_lo = __ldg(prior_range + 2 * v), g_hi = __ldg(prior_range + 2 * v + 1);
        const float b_hi = fmaf(fabsf(g_hi) + a.depth_diff_max, 1.00002f, g_hi > 0.f ? 0.f : 2.f * g_hi);        // (g_hi + thr) + slack
        const float b_lo = fmaf(fabsf(g_lo) + a.depth_diff_max, -1.00002f, g_lo > 0.f ? 2.f * g_lo : 0.f);       // (g_lo - thr) - slack
        // two candidates per lane and iteration: two independent dependency chains for the scheduler
        const float* cp = cand + tid;
        for (int i0 = 0; i0 < nc; i0 += 64, cp += 64) {
          const int ia = i0 + tid, ib = ia + 32;
          bool need_a = false, need_b = false;
          {
            const float ta = ia < nc ? __ldg(cp) : 0.f, tb = ib < nc ? __ldg(cp + 32) : 0.f;
            const float ca0 = fmaf(B0, ta, A0), ca1 = fmaf(B1, ta, A1), ca2 = fmaf(B2, ta, A2), da = fmaf(magB, fabsf(ta), magA);
            const float cb0 = fmaf(B0, tb, A0), cb1 = fmaf(B1, tb, A1), cb2 = fmaf(B2, tb, A2), db = fmaf(magB, fabsf(tb), magA);
            const float ha = fmaf(2.f, da, b_hi), la = fmaf(-2.f, da, b_lo), hb = fmaf(2.f, db, b_hi), lb = fmaf(-2.f, db, b_lo);
            const float ra = ca0 * ca0 + ca1 * ca1 + ca2 * ca2, rb = cb0 * cb0 + cb1 * cb1 + cb2 * cb2;
            const bool out_a = ra > ha * ha * 1.000001f || (la > 0.f && ra * 1.000001f < la * la);
            const bool out_b = rb > hb * hb * 1.000001f || (lb > 0.f && rb * 1.000001f < lb * lb);
            if (!__any_sync(0xffffffffu, (ia < nc && !out_a) || (ib < nc && !out_b))) continue;   // the whole batch is off the prior's range
            const bool far_a = out_a || dg_certainly_far(flt, dmap, a.map_w, a.depth_diff_max, ca0, ca1, ca2, da);
            const bool far_b = out_b || dg_certainly_far(flt, dmap, a.map_w, a.depth_diff_max, cb0, cb1, cb2, db);
            need_a = ia < nc && !far_a;
            need_b = ib < nc && !far_b;
          }
          const unsigned ma = __ballot_sync(0xffffffffu, need_a), mb = __ballot_sync(0xffffffffu, need_b);
          if (ma | mb) {
            if (n_q + __popc(ma) + __popc(mb) > kDgQueue) drain();
            if (need_a) queue[n_q + __popc(ma & ((1u << tid) - 1u))] = ia | (v << 16);
            n_q += __popc(ma);
            if (need_b) queue[n_q + __popc(mb & ((1u << tid) - 1u))] = ib | (v << 16);
            n_q += __popc(mb);
          }
        }
      }
      drain();
      for (int i0 = 0; i0 < nc; i0 += 32) {
        const int i = i0 + tid;
        const float best = i < nc ? lik[i] : 0.f;
        if (i < nc && a.likelihood) a.likelihood[(size_t)ray * nc + i] = best;
        const unsigned alive = __ballot_sync(0xffffffffu, best > 0.f);
        if (best > 0.f) {
          const int pos = count + __popc(alive & ((1u << tid) - 1u));
          keys[pos] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
        }
        count += __popc(alive);
      }
    } else
    // ---- phase 1: likelihood of every candidate, max over views; survivors are compacted as sortable keys
    for (int i0 = 0; i0 < nc; i0 += 32) {
      const int i = i0 + tid;
      float best = 0.f;
      if (i < nc) {
      if (from_dict) {
        const float* qd = a.que_dir + ((size_t)ray * nc + i) * 3;
        const float q0 = -__ldg(qd), q1 = -__ldg(qd + 1), q2 = -__ldg(qd + 2);
        for (int v = 0; v < a.rfn; ++v) {
          const size_t row = ((size_t)v * a.rn + ray) * nc + i;
          float cosv = 0.f;
          if (a.include_norm) {
            const float* w = a.ref_w2c + 12 * v;
            const float d0 = w[0] * q0 + w[1] * q1 + w[2] * q2;
            const float d1 = w[4] * q0 + w[5] * q1 + w[6] * q2;
            const float d2 = w[8] * q0 + w[9] * q1 + w[10] * q2;
            const float* n = a.prj_normal + 3 * row;
            cosv = d0 * __ldg(n) + d1 * __ldg(n + 1) + d2 * __ldg(n + 2);
          }
          best = fmaxf(best, surface_likelihood(a, __ldg(a.prj_mu + row), __ldg(a.prj_uncert + row), __ldg(a.prj_depth + row), cosv));
        }
      } else {
        for (int v = 0; v < a.rfn; ++v) best = fmaxf(best, exact_lik(i, v));
      }
      lik[i] = best;
      if (a.likelihood) a.likelihood[(size_t)ray * nc + i] = best;
      }
      // survivors are compacted in candidate order (ballot prefix): deterministic, no atomics
      const unsigned alive = __ballot_sync(0xffffffffu, best > 0.f);
      if (best > 0.f) {
        // descending key order = descending likelihood, ties: lower candidate index first (stable descending sort)
        const int pos = count + __popc(alive & ((1u << tid) - 1u));
        keys[pos] = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
      }
      count += __popc(alive);
    }
    __syncwarp();
    // ---- phase 2: more survivors than slots -> order them
    if (count > keep) {
      const int P = next_pow2(count);
      for (int i = count + tid; i < P; i += 32) keys[i] = 0ull;
      __syncwarp();
      bitonic_sort(keys, P, [](unsigned long long x, unsigned long long y) { return x > y; }, tid);
    }
    // ---- phase 3: Gaussian samples around the occlusion-aware mean (original_depth_guided_sample.py:199-202,257-275)
    float g_mean = 0.f, g_std = 0.f;
    bool g_on = false;
    if (ng > 0) {
      {
        float T = 1.f;                                                    // prod_{j<i} (1 - lik_j), sequential fp32
        for (int c0 = 0; c0 < nc; c0 += 32) {
          const int i = c0 + tid;
          const float l = i < nc ? lik[i] : 0.f;
          unsigned m = __ballot_sync(0xffffffffu, l > 0.f);
          float mine = 0.f;
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float lb = __shfl_sync(0xffffffffu, l, b);
            if (tid == b) mine = lb * T;
            T = T * (1.f - lb);
          }
          if (i < nc) opq[i] = mine;
        }
      }
      __syncwarp();
      float part = 0.f;
      for (int i = tid; i < nc; i += 32) part += opq[i];
      const float S = warp_sum(part);
      g_on = S != 0.f;
      if (g_on) {
        part = 0.f;
        // entries with zero opacity add +-0 to a partial sum: skipping them is bit-identical and saves the division
        for (int i = tid; i < nc; i += 32) { const float o = opq[i]; if (o != 0.f) part += __ldg(cand + i) * (o / S); }
        g_mean = warp_sum(part);
        part = 0.f;
        for (int i = tid; i < nc; i += 32) {
          const float o = opq[i];
          if (o == 0.f) continue;
          const float d = __ldg(cand + i) - g_mean;
          part += d * d * (o / S);
        }
        g_std = sqrtf(warp_sum(part));
      }
    }
    // ---- phase 4: slots = [most likely candidates | Gaussian samples], 0 = empty
    const int n_sel = count < keep ? count : keep;
    auto slot_value = [&](int s) -> float {
      if (s < n_sel) return __ldg(cand + (0xFFFFFFFFu - (unsigned)(keys[s] & 0xFFFFFFFFull)));
      if (s < keep) return 0.f;
      if (s < ns) return g_on ? __fadd_rn(__fmul_rn(__ldg(a.gauss + ray * ng + (s - keep)), g_std), g_mean) : 0.f;
      return __int_as_float(0x7f800000);
    };
    if constexpr (E > 0) {
      float x[E];
#pragma unroll
      for (int r = 0; r < E; ++r) x[r] = slot_value(r * 32 + tid);
      warp_sort_regs<E>(x, tid);
      // ---- fill_up_uniform_samples (original_depth_guided_sample.py:333-366)
      int mine = 0;
#pragma unroll
      for (int r = 0; r < E; ++r) mine += (r * 32 + tid < ns && x[r] == 0.f) ? 1 : 0;
      const float n_miss_f = warp_sum((float)mine);
      if (n_miss_f > 0.f) {
        const float step = (a.max_depth - a.min_depth) / n_miss_f;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int s = r * 32 + tid;
          if (s < ns && x[r] == 0.f) {
            float zf = __fadd_rn(a.min_depth, __fmul_rn((float)s, step));
            x[r] = __fadd_rn(zf, __fmul_rn(__ldg(a.fill_rand + ray * ns + s), step));
          }
        }
      }
      // ---- optional uniform samples (renderer.py:346-349), final sort
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int s = r * 32 + tid;
        if (s >= ns) x[r] = (s < n_out) ? __ldg(a.uniform_depth + (s - ns)) : __int_as_float(0x7f800000);
      }
      warp_sort_regs<E>(x, tid);
#pragma unroll
      for (int r = 0; r < E; ++r)
        if (r * 32 + tid < n_out) a.out_depth[ray * n_out + r * 32 + tid] = x[r];
      __syncwarp();
    } else {
    for (int s = tid; s < ns_pow2; s += 32) z[s] = slot_value(s);
    __syncwarp();
    bitonic_sort(z, ns_pow2, [](float x, float y) { return x < y; }, tid);
    // ---- fill_up_uniform_samples (original_depth_guided_sample.py:333-366)
    int mine = 0;
    for (int s = tid; s < ns; s += 32) mine += (z[s] == 0.f) ? 1 : 0;
    const float n_miss_f = warp_sum((float)mine);
    if (n_miss_f > 0.f) {
      const float step = (a.max_depth - a.min_depth) / n_miss_f;
      for (int s = tid; s < ns; s += 32)
        if (z[s] == 0.f) {
          float zf = __fadd_rn(a.min_depth, __fmul_rn((float)s, step));
          zf = __fadd_rn(zf, __fmul_rn(__ldg(a.fill_rand + ray * ns + s), step));
          z[s] = zf;
        }
    }
    __syncwarp();
    // ---- optional uniform samples (renderer.py:346-349), final sort
    for (int s = ns + tid; s < out_pow2; s += 32) z[s] = (s < n_out) ? __ldg(a.uniform_depth + (s - ns)) : __int_as_float(0x7f800000);
    __syncwarp();
    bitonic_sort(z, out_pow2, [](float x, float y) { return x < y; }, tid);
    for (int s = tid; s < n_out; s += 32) a.out_depth[ray * n_out + s] = z[s];
    __syncwarp();
    }
  }
}

// project_points_dict_diner: thread = (view, point)
__global__ void __launch_bounds__(256) project_gather_diner_kernel(const float* __restrict__ pts, long long pn, const float* __restrict__ w2c,
                                                                   int rfn, int dataset, int H, int W, const float* __restrict__ mvs_depth,
                                                                   const float* __restrict__ mvs_uncert, const float* __restrict__ mvs_normal,
                                                                   int map_h, int map_w, int img_h, int img_w, float* __restrict__ out_pix,
                                                                   float* __restrict__ out_depth, float* __restrict__ out_mu,
                                                                   float* __restrict__ out_uncert, float* __restrict__ out_normal) {
  const size_t map_px = (size_t)map_h * map_w;
  const long long total = pn * rfn;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(g / pn);
    const long long pi = g % pn;
    const float x = __ldg(pts + 3 * pi), y = __ldg(pts + 3 * pi + 1), zc = __ldg(pts + 3 * pi + 2);
    const float* w = w2c + 12 * v;
    const float c0 = w[0] * x + w[1] * y + w[2] * zc + w[3];
    const float c1 = w[4] * x + w[5] * y + w[6] * zc + w[7];
    const float c2 = w[8] * x + w[9] * y + w[10] * zc + w[11];
    float radius, px, py;
    cam_to_equi(dataset, c0, c1, c2, H, W, radius, px, py);
    const Footprint f = border_footprint(px, py, img_h, img_w, map_h, map_w);
    out_pix[2 * g] = px; out_pix[2 * g + 1] = py;
    out_depth[g] = radius;
    out_mu[g] = tap1(mvs_depth + v * map_px, f, map_w);
    out_uncert[g] = tap1(mvs_uncert + v * map_px, f, map_w);
    if (out_normal) {
      const float* nm = mvs_normal + (size_t)v * 3 * map_px;
      out_normal[3 * g] = tap1(nm, f, map_w);
      out_normal[3 * g + 1] = tap1(nm + map_px, f, map_w);
      out_normal[3 * g + 2] = tap1(nm + 2 * map_px, f, map_w);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// depth2normal (network/orig_diner_depth2normal.py:7-110): prior normals for `backface_culling`.
// Pass 1 (thread = pixel): back-project the 4-neighbourhood (zero rows above / below the panorama, longitude wrap), cross product
// of the vertical and horizontal central differences, normalise; record the reference's "cleaning" offset (a neighbour whose x
// coordinate is exactly 0 pushes the lookup one pixel to the opposite side).  Pass 2: gather the RAW normal at the offset pixel,
// zero where the depth is 0, write (N,3,H,W).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth2normal_raw_kernel(const float* __restrict__ dmap, int N, int H, int W, int dataset,
                                                               float* __restrict__ raw, signed char* __restrict__ offs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const float* d = dmap + (i / ((long long)H * W)) * H * W;
  auto point = [&](int yy, int xx, float (&o)[3]) {
    if (yy < 0 || yy >= H) { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; return; }
    float dx, dy, dz;
    equi_unit_dir<false>(dataset, (float)xx, (float)yy, H, W, dx, dy, dz);
    const float dd = __ldg(d + (size_t)yy * W + xx);
    o[0] = __fmul_rn(dx, dd); o[1] = __fmul_rn(dy, dd); o[2] = __fmul_rn(dz, dd);
  };
  float dn[3], up[3], rt[3], lf[3];
  point(y + 1, x, dn); point(y - 1, x, up);
  point(y, x + 1 == W ? 0 : x + 1, rt); point(y, x == 0 ? W - 1 : x - 1, lf);
  float v[3], h[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { v[k] = __fsub_rn(dn[k], up[k]); h[k] = __fsub_rn(rt[k], lf[k]); }
  float n[3] = {__fsub_rn(__fmul_rn(v[1], h[2]), __fmul_rn(v[2], h[1])), __fsub_rn(__fmul_rn(v[2], h[0]), __fmul_rn(v[0], h[2])),
                __fsub_rn(__fmul_rn(v[0], h[1]), __fmul_rn(v[1], h[0]))};
  const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(n[0], n[0]), __fmul_rn(n[1], n[1])), __fmul_rn(n[2], n[2])));
  raw[3 * i] = __fdiv_rn(n[0], len); raw[3 * i + 1] = __fdiv_rn(n[1], len); raw[3 * i + 2] = __fdiv_rn(n[2], len);
  offs[2 * i] = (signed char)((up[0] == 0.f ? 1 : 0) - (dn[0] == 0.f ? 1 : 0));
  offs[2 * i + 1] = (signed char)((lf[0] == 0.f ? 1 : 0) - (rt[0] == 0.f ? 1 : 0));
}
__global__ void __launch_bounds__(256) depth2normal_clean_kernel(const float* __restrict__ dmap, const float* __restrict__ raw,
                                                                 const signed char* __restrict__ offs, int N, int H, int W,
                                                                 float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * H * W) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const long long n = i / ((long long)H * W);
  const int yy = min(max(y + offs[2 * i], 0), H - 1), xx = min(max(x + offs[2 * i + 1], 0), W - 1);
  const long long j = (n * H + yy) * W + xx;
  const bool hole = __ldg(dmap + i) == 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) out[((n * 3 + k) * H + y) * W + x] = hole ? 0.f : raw[3 * j + k];
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_depth_guided_sample_fwd(const pgrf_diner_args* args, void* stream) {
  PGRF_REQUIRE(args != nullptr, "depth_guided_sample: null args");
  const pgrf_diner_args& a = *args;
  PGRF_REQUIRE(a.rfn >= 1 && a.rn >= 1, "depth_guided_sample: rfn=%d rn=%lld", a.rfn, a.rn);
  PGRF_REQUIRE(a.n_candidates >= 2 && a.n_candidates <= kDgMaxCand, "depth_guided_sample: n_candidates=%d not in [2,%d]",
               a.n_candidates, kDgMaxCand);
  PGRF_REQUIRE(a.n_samples >= 1 && a.n_uniform >= 0 && a.n_samples + a.n_uniform <= kDgMaxOut,
               "depth_guided_sample: n_samples=%d n_uniform=%d (at most %d together)", a.n_samples, a.n_uniform, kDgMaxOut);
  PGRF_REQUIRE(a.n_gaussian >= 0 && a.n_samples >= a.n_gaussian, "depth_guided_sample: n_samples >= n_gaussian required");
  PGRF_REQUIRE(a.n_samples <= a.n_candidates, "depth_guided_sample: n_samples > n_candidates");
  PGRF_REQUIRE(a.cand_depth && a.ref_w2c && a.fill_rand && a.out_depth, "depth_guided_sample: null pointer argument");
  PGRF_REQUIRE(a.n_gaussian == 0 || a.gauss, "depth_guided_sample: n_gaussian > 0 needs the gauss table");
  PGRF_REQUIRE(a.n_uniform == 0 || a.uniform_depth, "depth_guided_sample: n_uniform > 0 needs uniform_depth");
  if (a.prj_mu) {
    PGRF_REQUIRE(a.prj_uncert && a.prj_depth && a.que_dir && (!a.include_norm || a.prj_normal),
                 "depth_guided_sample: dict variant needs prj_uncert, prj_depth, que_dir (and prj_normal with include_norm)");
  } else {
    PGRF_REQUIRE(a.dataset >= 0 && a.dataset <= 3, "Unknown dataset id %d", a.dataset);
    PGRF_REQUIRE(a.coords && a.que_c2w && a.mvs_depth && a.mvs_uncert && (!a.include_norm || a.mvs_normal),
                 "depth_guided_sample: fused variant needs coords, que_c2w, mvs_depth, mvs_uncert (and mvs_normal with include_norm)");
    PGRF_REQUIRE(a.H >= 2 && a.W >= 2 && a.map_h >= 1 && a.map_w >= 1 && a.img_h >= 2 && a.img_w >= 2, "depth_guided_sample: bad map sizes");
  }
  const int nc_pad = (a.n_candidates + 3) & ~3;
  int nc_pow2 = 1; while (nc_pow2 < a.n_candidates) nc_pow2 <<= 1;
  int out_pow2 = 1; while (out_pow2 < a.n_samples + a.n_uniform) out_pow2 <<= 1;
  const size_t smem = kDgWarps * ((size_t)nc_pow2 * 8 + (size_t)nc_pad * 4 + (size_t)out_pow2 * 4 + kDgQueue * 4);
  PGRF_REQUIRE(smem <= 227 * 1024, "depth_guided_sample: %zu bytes of shared memory", smem);
  const long long max_grid = 148LL * 16, want = (a.rn + kDgWarps - 1) / kDgWarps;
  const int grid = (int)(want < max_grid ? want : max_grid);
  const int E = (g_dg_regsort && out_pow2 <= 128) ? (out_pow2 <= 32 ? 1 : out_pow2 / 32) : 0;
  // range of the prior maps (fused variant only), in stream order
  float* prior_range = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (!a.prj_mu) {
    PGRF_CUDA(cudaMallocAsync((void**)&prior_range, sizeof(float) * 2 * a.rfn, st));
    depth_range_kernel<<<a.rfn, 1024, 0, st>>>(a.mvs_depth, (long long)a.map_h * a.map_w, prior_range);
    count_launch();
  }
#define PGRF_DG_LAUNCH(EE)                                                                                                            \
  do {                                                                                                                                \
    if (smem > 48 * 1024) PGRF_CUDA(cudaFuncSetAttribute(depth_guided_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    depth_guided_kernel<EE><<<grid, kDgThreads, smem, st>>>(a, nc_pad, nc_pow2, g_dg_prefilter, prior_range);                           \
  } while (0)
  if (E == 1) PGRF_DG_LAUNCH(1);
  else if (E == 2) PGRF_DG_LAUNCH(2);
  else if (E == 4) PGRF_DG_LAUNCH(4);
  else PGRF_DG_LAUNCH(0);
#undef PGRF_DG_LAUNCH
  if (prior_range) PGRF_CUDA(cudaFreeAsync(prior_range, st));
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_project_gather_diner_fwd(const float* pts, long long pn, const float* w2c, int rfn, int dataset, int H, int W,
                                             const float* mvs_depth, const float* mvs_uncert, const float* mvs_normal, int map_h,
                                             int map_w, int img_h, int img_w, float* out_pix, float* out_depth, float* out_mu,
                                             float* out_uncert, float* out_normal, void* stream) {
  PGRF_REQUIRE(pts && w2c && mvs_depth && mvs_uncert && out_pix && out_depth && out_mu && out_uncert, "project_gather_diner: null pointer argument");
  PGRF_REQUIRE((out_normal == nullptr) || mvs_normal, "project_gather_diner: out_normal needs mvs_normal");
  PGRF_REQUIRE(pn >= 1 && rfn >= 1, "project_gather_diner: pn=%lld rfn=%d", pn, rfn);
  PGRF_REQUIRE(dataset >= 0 && dataset <= 3, "Unknown dataset id %d", dataset);
  const long long total = pn * rfn;
  const long long blocks = (total + 255) / 256;
  const int grid = (int)(blocks < 148LL * 8 ? blocks : 148LL * 8);
  project_gather_diner_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pts, pn, w2c, rfn, dataset, H, W, mvs_depth, mvs_uncert, mvs_normal,
                                                                    map_h, map_w, img_h, img_w, out_pix, out_depth, out_mu, out_uncert,
                                                                    out_normal);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}


extern "C" int pgrf_depth2normal_fwd(const float* mvs_depth, int N, int H, int W, int dataset, float* raw_ws, signed char* off_ws,
                                     float* out_normal, void* stream) {
  PGRF_REQUIRE(mvs_depth && raw_ws && off_ws && out_normal, "depth2normal: null pointer argument");
  PGRF_REQUIRE(N >= 1 && H >= 2 && W >= 2 && dataset >= 0 && dataset <= 3, "depth2normal: bad arguments N=%d H=%d W=%d", N, H, W);
  const long long total = (long long)N * H * W;
  const unsigned grid = (unsigned)((total + 255) / 256);
  depth2normal_raw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mvs_depth, N, H, W, dataset, raw_ws, off_ws);
  depth2normal_clean_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(mvs_depth, raw_ws, off_ws, N, H, W, out_normal);
  count_launch(2);
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

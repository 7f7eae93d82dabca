// Definitions shared by the forward (cost_volume.cu) and backward (cost_volume_bwd.cu) sweep kernels.
#pragma once
#include "common.cuh"

namespace pgrf {

constexpr int kCvWarps = 4;
constexpr int kCvThreads = kCvWarps * 32;
constexpr int kMaxSrc = 8;

struct CvParams {
  const float* images;
  const float* depths;
  const float* depth_volume;
  const float* rots;
  const float* trans;
  float* out;
  int* err;
  int B, S, H, W, D;
  int ref_idx, n_src;
  int src_views[kMaxSrc];
  float divisor;
  int dataset, cost_type, groups, OC;
  int d_chunk;
  int out_bf16;            // channels-last output stored as bf16 (PGRF_CV_BDHWC_BF16)
  long long sB, sD, sC;  // planar strides in elements
  float ang0, ang1, ang2, ang3;  // per-dataset pixel->angle constants (see host side)
};

struct __align__(16) TapRec {
  float tx, ty;   // fractional offsets inside the 2x2 footprint
  int off4;       // float4 index of the north-west texel inside one (H,W,C) view
  int pad;
};
// The footprint is canonicalised so that all four taps are inside the map: for uv in [-1,1] the only
// out-of-map tap of grid_sample(zeros padding) is x0+1 == W (or y0+1 == H) reached with weight exactly 0
// when ix == W-1; shifting the footprint one texel back (x0 = W-2, tx = 1) gives bit-identical weights
// (1-tx = 0 on the west taps) and needs no predication.  Out-of-range uv (flagged, the reference asserts)
// is clamped the same way so the gather stays memory-safe.

// ---- pixel -> unit ray, per dataset (spherical_cost_volume.py:272-301, my_torch_helpers.py:33-58) ----
__device__ __forceinline__ void pixel_ray(const CvParams& p, int x, int y, float& rx, float& ry, float& rz) {
  const float fx = (float)x, fy = (float)y;
  float theta, phi;
  switch (p.dataset) {
    case PGRF_DS_M3D:
      phi = (fy + 0.5f) * p.ang0;                       // (phi+0.5)*(pi/H)
      theta = (fx + 0.5f) * p.ang1 - PGRF_HALF_PI_F;    // (theta+0.5)*(2pi/W) - pi/2
      break;
    case PGRF_DS_REPLICA_TEST:
      theta = p.ang1 * (fx + 0.5f) - PGRF_PI_F;
      phi = (-(fy + 0.5f) * PGRF_PI_F) / p.ang0 + PGRF_HALF_PI_F;   // ang0 = H
      break;
    case PGRF_DS_RESIDENTIAL:
      theta = PGRF_PI_F * ((2.f * fx) / p.ang1 - 1.5f);             // ang1 = W-1
      phi = PGRF_PI_F * (0.5f - fy / p.ang0);                       // ang0 = H-1
      break;
    default:  // CoffeeArea
      theta = p.ang1 * fx + PGRF_TWO_PI_F;                          // ang1 = -2pi/(W-1)
      phi = p.ang0 * fy;                                            // ang0 = pi/(H-1)
      break;
  }
  float st, ct, sp, cp;
  sincosf(theta, &st, &ct);
  sincosf(phi, &sp, &cp);
  switch (p.dataset) {
    case PGRF_DS_M3D:          rx = sp * ct; ry = cp;  rz = sp * st; break;
    case PGRF_DS_REPLICA_TEST: rx = st * cp; ry = -sp; rz = ct * cp; break;
    case PGRF_DS_RESIDENTIAL:  rx = ct * cp; ry = sp;  rz = st * cp; break;
    default:                   rx = sp * ct; ry = sp * st; rz = cp;  break;
  }
}

// ---- camera-frame point -> normalised (u,v) (my_torch_helpers.py:62-120 + spherical_cost_volume.py:153-190) ----
__device__ __forceinline__ void point_uv(int dataset, float cx, float cy, float cz, float& u, float& v) {
  const float kLin = 0.17453292519943295f;       // deg2rad(10)
  const float kCosDeg = 0.984807753012208f;      // cos(10 deg)
  const float kOneMinusCos = 0.015192246987791981f;
  const float radius = sqrtf(cx * cx + cy * cy + cz * cz);
  float uu, vv;
  switch (dataset) {
    case PGRF_DS_M3D: {
      const float theta = atan2f(cz, cx);
      const float yr = cy / radius;
      float phi;
      if (fabsf(yr) < kCosDeg) phi = acosf(yr);
      else if (cy >= 0.f) phi = kLin * (1.f - yr) / kOneMinusCos;       // acos linearised near the poles
      else phi = PGRF_PI_F - kLin * (yr + 1.f) / kOneMinusCos;
      uu = fmod_two_pi(theta + PGRF_HALF_PI_F + PGRF_TWO_PI_F);
      vv = phi;
      break;
    }
    case PGRF_DS_REPLICA_TEST: {
      const float theta = atan2f(cx, cz);
      const float phi = -asinf(cz / radius);      // sic (reference :99 uses z)
      uu = fmod_two_pi(theta + PGRF_PI_F + PGRF_TWO_PI_F);
      vv = -phi + PGRF_HALF_PI_F;
      break;
    }
    case PGRF_DS_RESIDENTIAL: {
      float theta = -atan2f(-cz, cx);
      const float phi = asinf(cy / radius);
      if (theta > PGRF_HALF_PI_F && theta <= PGRF_TWO_PI_F) theta -= PGRF_TWO_PI_F;
      uu = fmod_two_pi(theta + 4.71238898038468985769f);   // 3/4 * 2pi
      vv = PGRF_HALF_PI_F - phi;
      break;
    }
    default: {
      float theta = atan2f(cy, cx);
      const float phi = acosf(cz / radius);
      if (theta < 0.f) theta += PGRF_TWO_PI_F;
      uu = PGRF_TWO_PI_F - theta;
      vv = phi;
      break;
    }
  }
  // tensor / python-scalar is a multiplication by the fp32 reciprocal in torch's CUDA kernels; do the same
  constexpr float kInvPi = 0.31830988618379067154f;
  u = uu * kInvPi - 1.f;
  v = 2.f * vv * kInvPi - 1.f;
}

// Bilinear blend of one 2x2 footprint (ATen's accumulation order nw, ne, sw, se) of 4 channels and the matching cost against the
// reference feature (`rf` holds -reference for abs_diff, so the difference is one packed add).  Packed fp32 pairs: mul/fma.rn.f32x2 perform the same IEEE operations as the scalar chain
// (bit-identical results) in half the issue slots — the sweep is issue bound (ncu: time proportional to executed instructions).
__device__ __forceinline__ float4 blend_cost(const float4& nw, const float4& ne, const float4& sw, const float4& se, float tx, float ty,
                                             const float4& rf, int cost_type) {
  const float tx1 = 1.f - tx, ty1 = 1.f - ty;   // exact: == (x0+1)-ix, (y0+1)-iy
  const float2 wn = fmul2(make_float2(tx1, tx), make_float2(ty1, ty1));     // (wnw, wne)
  const float2 ws = fmul2(make_float2(tx1, tx), make_float2(ty, ty));       // (wsw, wse)
  const float2 w0 = make_float2(wn.x, wn.x), w1 = make_float2(wn.y, wn.y), w2 = make_float2(ws.x, ws.x), w3 = make_float2(ws.y, ws.y);
  float2 lo = fmul2(make_float2(nw.x, nw.y), w0), hi = fmul2(make_float2(nw.z, nw.w), w0);
  lo = ffma2(make_float2(ne.x, ne.y), w1, lo); hi = ffma2(make_float2(ne.z, ne.w), w1, hi);
  lo = ffma2(make_float2(sw.x, sw.y), w2, lo); hi = ffma2(make_float2(sw.z, sw.w), w2, hi);
  lo = ffma2(make_float2(se.x, se.y), w3, lo); hi = ffma2(make_float2(se.z, se.w), w3, hi);
  if (cost_type == PGRF_COST_ABS_DIFF) {
    lo = fadd2(lo, make_float2(rf.x, rf.y)); hi = fadd2(hi, make_float2(rf.z, rf.w));
    return make_float4(fabsf(lo.x), fabsf(lo.y), fabsf(hi.x), fabsf(hi.y));
  }
  if (cost_type == PGRF_COST_DOT) {
    lo = fmul2(lo, make_float2(rf.x, rf.y)); hi = fmul2(hi, make_float2(rf.z, rf.w));
  }
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// Relative pose of swept view `slot` w.r.t. the reference view, in fp64, rounded once to fp32:
// dst[0..8] = A = R_s * R_ref^-1 (row-major), dst[9..11] = b = t_s - A t_ref   (spherical_cost_volume.py:139-150)
__device__ __forceinline__ void relative_pose(const CvParams& p, int b, int slot, float* dst) {
  const int s = p.src_views[slot];
  const float* Rr = p.rots + ((size_t)b * p.S + p.ref_idx) * 9;
  const float* tr = p.trans + ((size_t)b * p.S + p.ref_idx) * 3;
  const float* Rs = p.rots + ((size_t)b * p.S + s) * 9;
  const float* ts = p.trans + ((size_t)b * p.S + s) * 3;
  double r[9], inv[9];
  for (int i = 0; i < 9; ++i) r[i] = Rr[i];
  const double det = r[0] * (r[4] * r[8] - r[5] * r[7]) - r[1] * (r[3] * r[8] - r[5] * r[6]) +
                     r[2] * (r[3] * r[7] - r[4] * r[6]);
  const double id = 1.0 / det;
  inv[0] = (r[4] * r[8] - r[5] * r[7]) * id; inv[1] = (r[2] * r[7] - r[1] * r[8]) * id; inv[2] = (r[1] * r[5] - r[2] * r[4]) * id;
  inv[3] = (r[5] * r[6] - r[3] * r[8]) * id; inv[4] = (r[0] * r[8] - r[2] * r[6]) * id; inv[5] = (r[2] * r[3] - r[0] * r[5]) * id;
  inv[6] = (r[3] * r[7] - r[4] * r[6]) * id; inv[7] = (r[1] * r[6] - r[0] * r[7]) * id; inv[8] = (r[0] * r[4] - r[1] * r[3]) * id;
  double A[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      A[i * 3 + j] = (double)Rs[i * 3 + 0] * inv[0 * 3 + j] + (double)Rs[i * 3 + 1] * inv[1 * 3 + j] +
                     (double)Rs[i * 3 + 2] * inv[2 * 3 + j];
  for (int i = 0; i < 9; ++i) dst[i] = (float)A[i];
  for (int i = 0; i < 3; ++i)
    dst[9 + i] = (float)((double)ts[i] - (A[i * 3] * (double)tr[0] + A[i * 3 + 1] * (double)tr[1] + A[i * 3 + 2] * (double)tr[2]));
}

// host side: fills the per-dataset angle constants and validates what both directions share
int cv_fill_params(CvParams& p, const float* images, int B, int S, int H, int W, int C, const float* depths, const float* depth_volume,
                   int D, const float* rots, const float* trans, int ref_idx, const int* src_views, int n_src, float divisor,
                   int dataset, int cost_type);

}  // namespace pgrf

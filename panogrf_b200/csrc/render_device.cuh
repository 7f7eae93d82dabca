// Device building blocks of the render-path kernels: activations, the shared-memory SIMT GEMM used
// for every Linear layer of the fp32 parity path, TMA (cp.async.bulk) helpers, and the ERP geometry
// of the renderer (network/spt_utils.py, network/render_ops.py).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include "render_layout.cuh"

namespace pgrf {

enum : int { ACT_NONE = 0, ACT_ELU = 1, ACT_RELU = 2 };

// exp(x) as ONE MUFU.EX2 (ex2.approx.ftz, 2 ulp): __expf without -ftz expands to a branchy denormal-safe sequence,
// which dominated the epilogues of the render kernels (ncu: 15 % of all issued instructions).
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// branch-free ELU: max(x,0) + min(exp(x)-1, 0)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ELU(alpha=1) of two values with packed arithmetic: FMUL2 + 2 MUFU + FADD2 + 2 (FSETP+FSEL)
__device__ __forceinline__ float2 elu_pair(float2 v) {
  const float2 t = fmul2(v, make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 r = fadd2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(-1.f, -1.f));
  return make_float2(v.x > 0.f ? v.x : r.x, v.y > 0.f ? v.y : r.y);
}
// bf16x2 ELU of a biased fp32 pair: the select runs on the PACKED values (max(v, min(exp(v)-1, 0)) with HMNMX2); rounding is
// monotonic and sign preserving, so this equals rounding the fp32 ELU
__device__ __forceinline__ uint32_t elu_pack(float2 v) {
  const float2 t = fmul2(v, make_float2(1.4426950408889634f, 1.4426950408889634f));
  const float2 r = fadd2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(-1.f, -1.f));
  const __nv_bfloat162 V = __floats2bfloat162_rn(v.x, v.y), R = __floats2bfloat162_rn(r.x, r.y);
  const __nv_bfloat162 o = __hmax2(V, __hmin2(R, __floats2bfloat162_rn(0.f, 0.f)));
  return *reinterpret_cast<const uint32_t*>(&o);
}
__device__ __forceinline__ float elu1(float x) { return fmaxf(x, 0.f) + fminf(fast_exp(x) - 1.f, 0.f); }
__device__ __forceinline__ float sigmoidf(float x) { return fast_rcp(1.f + fast_exp(-x)); }
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(fast_exp(x)); }
template <int ACT>
__device__ __forceinline__ float activate(float x) {
  if (ACT == ACT_ELU) return elu1(x);
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  return x;
}

// -------------------------------------------------------------------------------------------------
// C[n][m] = ACT( rs[m] * sum_k A[k][m] * Wt[k][n] + bias[n] + add[n][m % T] )        (all in smem)
//   A  : [K][lda]  feature-major activations (m contiguous)
//   Wt : [K][ldw]  k-major weights (n contiguous)
//   C  : [N][ldc]
// Work is cut into warp tiles of 32 rows x (4*TN) columns; lane = (g = lane>>3 : column group,
// rg = lane&7 : group of 4 consecutive rows) owns a 4 x TN register tile.  Per k step a lane issues
// one 128-bit load of A (8 distinct addresses per warp, broadcast over g) and TN/4 128-bit loads of
// Wt (4 distinct addresses, broadcast over rg): 4*TN FMAs per 1+TN/4 shared-memory wavefronts.
// -------------------------------------------------------------------------------------------------
template <int TN, int ACT, bool ROW_SCALE, bool SAMPLE_ADD>
__device__ __forceinline__ void gemm_smem(const float* __restrict__ A, int lda, int K,
                                          const float* __restrict__ Wt, int ldw,
                                          const float* __restrict__ bias,
                                          float* __restrict__ C, int ldc, int M, int N,
                                          const float* __restrict__ row_scale,
                                          const float* __restrict__ sample_add, int T, int ld_add,
                                          int warp, int lane, int nwarps) {
  constexpr int CT = 4 * TN;  // columns per warp tile
  const int ntm = M >> 5, ntn = (N + CT - 1) / CT;
  const int rg = lane & 7, g = lane >> 3;
  for (int t = warp; t < ntm * ntn; t += nwarps) {
    const int m0 = (t % ntm) * 32 + 4 * rg;
    const int n0 = (t / ntm) * CT + TN * g;
    if (n0 >= N) continue;  // ragged last column tile (N multiple of TN but not of CT)
    float acc[4][TN];
    // accumulators as packed fp32 pairs over two neighbouring output columns: FFMA2 performs the same two fused
    // multiply-adds (bit-identical results) in one issue slot
    float2 acc2[4][TN / 2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) acc2[i][j] = make_float2(0.f, 0.f);
    const float* a_ptr = A + m0;
    const float* w_ptr = Wt + n0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(a_ptr + k * lda);
      float2 w2[TN / 2];
#pragma unroll
      for (int j4 = 0; j4 < TN / 4; ++j4) {
        const float4 wv = *reinterpret_cast<const float4*>(w_ptr + k * ldw + 4 * j4);
        w2[2 * j4] = make_float2(wv.x, wv.y); w2[2 * j4 + 1] = make_float2(wv.z, wv.w);
      }
      const float2 ax = make_float2(a.x, a.x), ay = make_float2(a.y, a.y), az = make_float2(a.z, a.z), aw = make_float2(a.w, a.w);
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) {
        acc2[0][j] = ffma2(ax, w2[j], acc2[0][j]);
        acc2[1][j] = ffma2(ay, w2[j], acc2[1][j]);
        acc2[2][j] = ffma2(az, w2[j], acc2[2][j]);
        acc2[3][j] = ffma2(aw, w2[j], acc2[3][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN / 2; ++j) { acc[i][2 * j] = acc2[i][j].x; acc[i][2 * j + 1] = acc2[i][j].y; }
    float rs[4] = {1.f, 1.f, 1.f, 1.f};
    if (ROW_SCALE) {
      const float4 r = *reinterpret_cast<const float4*>(row_scale + m0);
      rs[0] = r.x; rs[1] = r.y; rs[2] = r.z; rs[3] = r.w;
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + j;
      const float b = bias ? bias[n] : 0.f;
      float4 add = make_float4(0.f, 0.f, 0.f, 0.f);
      if (SAMPLE_ADD) add = *reinterpret_cast<const float4*>(sample_add + n * ld_add + (m0 % T));
      float4 o;
      o.x = activate<ACT>(fmaf(rs[0], acc[0][j], b) + add.x);
      o.y = activate<ACT>(fmaf(rs[1], acc[1][j], b) + add.y);
      o.z = activate<ACT>(fmaf(rs[2], acc[2][j], b) + add.z);
      o.w = activate<ACT>(fmaf(rs[3], acc[3][j], b) + add.w);
      *reinterpret_cast<float4*>(C + n * ldc + m0) = o;
    }
  }
}

// Small layers: one thread per row, outputs in registers.  Wt[k][NP] k-major, read as broadcast float4.
template <int K, int N, int NP>
__device__ __forceinline__ void row_layer(const float* __restrict__ A, int lda, int m,
                                          const float* __restrict__ Wt, const float* __restrict__ bias,
                                          float (&out)[NP]) {
#pragma unroll
  for (int n = 0; n < NP; ++n) out[n] = bias ? bias[n] : 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = A[k * lda + m];
#pragma unroll
    for (int n4 = 0; n4 < NP / 4; ++n4) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + k * NP + 4 * n4);
      out[4 * n4] = fmaf(a, w.x, out[4 * n4]);
      out[4 * n4 + 1] = fmaf(a, w.y, out[4 * n4 + 1]);
      out[4 * n4 + 2] = fmaf(a, w.z, out[4 * n4 + 2]);
      out[4 * n4 + 3] = fmaf(a, w.w, out[4 * n4 + 3]);
    }
  }
}
// same, input vector already in registers
template <int K, int N, int NP>
__device__ __forceinline__ void reg_layer(const float (&in)[K], const float* __restrict__ Wt,
                                          const float* __restrict__ bias, float (&out)[NP]) {
  // packed fp32 pairs over neighbouring output columns: the same IEEE fma per element (bit-identical), half the issue slots
  float2 acc[NP / 2];
#pragma unroll
  for (int n = 0; n < NP / 2; ++n) acc[n] = bias ? make_float2(bias[2 * n], bias[2 * n + 1]) : make_float2(0.f, 0.f);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float2 xk = make_float2(in[k], in[k]);
#pragma unroll
    for (int n4 = 0; n4 < NP / 4; ++n4) {
      const float4 w = *reinterpret_cast<const float4*>(Wt + k * NP + 4 * n4);
      acc[2 * n4] = ffma2(xk, make_float2(w.x, w.y), acc[2 * n4]);
      acc[2 * n4 + 1] = ffma2(xk, make_float2(w.z, w.w), acc[2 * n4 + 1]);
    }
  }
#pragma unroll
  for (int n = 0; n < NP / 2; ++n) { out[2 * n] = acc[n].x; out[2 * n + 1] = acc[n].y; }
}

// -------------------------------------------------------------------------------------------------
// TMA 1-D bulk copies + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
// global -> shared, completion signalled on the mbarrier (bytes multiple of 16, 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (bulk async group)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// -------------------------------------------------------------------------------------------------
// ERP geometry of the renderer — (W-1),(H-1) pixel convention (network/spt_utils.py)
// -------------------------------------------------------------------------------------------------

// pixel (x,y) -> unit direction in the camera frame (ray_utils.py:4-16, spt_utils.py:37-127)
template <bool FAST = false>   // FAST (bf16 path, m3d convention): reciprocal multiplies, __sincosf, rsqrt normalisation
__device__ __forceinline__ void equi_unit_dir(int dataset, float x, float y, int H, int W, float& dx, float& dy, float& dz) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  float theta, phi;
  switch (dataset) {
    case PGRF_DS_M3D:
      x = fminf(fmaxf(x, 0.f), wm1); y = fminf(fmaxf(y, 0.f), hm1);
      if (FAST) {
        theta = x * (2.f * PGRF_PI_F / wm1) - PGRF_HALF_PI_F;
        phi = y * (PGRF_PI_F / hm1);
      } else {
        theta = x / wm1 * 2.f * PGRF_PI_F - PGRF_HALF_PI_F;
        phi = y / hm1 * PGRF_PI_F;
      }
      break;
    case PGRF_DS_REPLICA_TEST:
      theta = x * 2.f * PGRF_PI_F / wm1 - PGRF_PI_F;
      phi = -y * PGRF_PI_F / hm1 + PGRF_HALF_PI_F;
      break;
    case PGRF_DS_RESIDENTIAL:
      x = fminf(fmaxf(x, 0.f), wm1); y = fminf(fmaxf(y, 0.f), hm1);
      theta = PGRF_PI_F * (2.f * x / wm1 - 1.5f);
      phi = PGRF_PI_F * (0.5f - y / hm1);
      break;
    default:
      x = fminf(fmaxf(x, 0.f), wm1); y = fminf(fmaxf(y, 0.f), hm1);
      theta = (float)(-2.0 * PGRF_PI_D / (double)(W - 1)) * x + PGRF_TWO_PI_F;
      phi = (float)(PGRF_PI_D / (double)(H - 1)) * y;
      break;
  }
  float st, ct, sp, cp;
  if (FAST && dataset == PGRF_DS_M3D) {   // |theta| <= 3pi/2, phi in [0, pi]: MUFU sin/cos, ~1e-6 absolute
    __sincosf(theta, &st, &ct);
    __sincosf(phi, &sp, &cp);
  } else {
    sincosf(theta, &st, &ct);
    sincosf(phi, &sp, &cp);
  }
  switch (dataset) {
    case PGRF_DS_M3D:          dx = sp * ct; dy = cp;  dz = sp * st; break;
    case PGRF_DS_REPLICA_TEST: dx = st * cp; dy = -sp; dz = ct * cp; break;
    case PGRF_DS_RESIDENTIAL:  dx = ct * cp; dy = sp;  dz = st * cp; break;
    default:                   dx = sp * ct; dy = sp * st; dz = cp;  break;
  }
  if (FAST) {
    const float in = rsqrtf(dx * dx + dy * dy + dz * dz);
    dx *= in; dy *= in; dz *= in;
  } else {
    const float n = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= n; dy /= n; dz /= n;
  }
}

// torch.remainder(a, b) for b > 0
__device__ __forceinline__ float py_mod(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && m < 0.f) m += b;
  return m;
}

// camera-frame point -> (radius, pixel x, pixel y) (ray_utils.py:18-22, spt_utils.py:129-199)
template <bool FAST = false>   // FAST (bf16 path, m3d convention): one MUFU.RCP, wrap by a compare, constant scales as multiplies
__device__ __forceinline__ void cam_to_equi(int dataset, float cx, float cy, float cz, int H, int W, float& radius, float& px, float& py) {
  const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
  radius = sqrtf(cx * cx + cy * cy + cz * cz);
  switch (dataset) {
    case PGRF_DS_M3D: {
      float theta = atan2f(cz, cx);
      if (FAST) {
        const float phi = acosf(cy * fast_rcp(radius + 1e-5f));
        theta += PGRF_HALF_PI_F;                       // in (-pi/2, 3pi/2]: python's % 2pi is one conditional add
        if (theta < 0.f) theta += PGRF_TWO_PI_F;
        px = theta * (wm1 / PGRF_TWO_PI_F);
        py = phi * (hm1 / PGRF_PI_F);
        break;
      }
      const float phi = acosf(cy / (radius + 1e-5f));
      theta = py_mod(theta + PGRF_HALF_PI_F, PGRF_TWO_PI_F);
      px = theta / PGRF_TWO_PI_F * wm1;
      py = phi / PGRF_PI_F * hm1;
      break;
    }
    case PGRF_DS_REPLICA_TEST: {
      const float theta = atan2f(cx, cz);
      const float phi = -asinf(cy / radius);
      px = (float)((double)(W - 1) / (2.0 * PGRF_PI_D)) * (theta + PGRF_PI_F);
      py = (float)((double)(H - 1) / PGRF_PI_D) * (-phi + PGRF_HALF_PI_F);
      break;
    }
    case PGRF_DS_RESIDENTIAL: {
      float theta = -atan2f(-cz, cx);
      const float phi = asinf(cy / radius);
      if (theta > PGRF_HALF_PI_F && theta <= PGRF_TWO_PI_F) theta -= PGRF_TWO_PI_F;
      px = ((float)(1.0 / (2.0 * PGRF_PI_D)) * theta + 0.75f) * wm1;
      py = (0.5f - phi / PGRF_PI_F) * hm1;
      break;
    }
    default: {
      float theta = atan2f(cy, cx);
      const float phi = acosf(cz / radius);
      if (theta < 0.f) theta += PGRF_TWO_PI_F;
      px = wm1 * (1.f - theta / PGRF_TWO_PI_F);
      py = phi * hm1 / PGRF_PI_F;
      break;
    }
  }
}

// Bilinear footprint of F.grid_sample(padding_mode='border') as called by interpolate_feats
// (network/ops.py:32-52): full-res pixel -> normalised -> map coordinates, clipped to the border.
struct Footprint {
  int off;        // texel index (y0*fw + x0) of the north-west tap
  int dx, dy;     // 1 when the east / south neighbour is inside the map, else 0 (weight is 0 there)
  float tx, ty;
};
__device__ __forceinline__ Footprint border_footprint(float px, float py, int h, int w, int fh, int fw) {
  const bool align = (fh == h && fw == w);
  const float xn = px / (float)(w - 1) * 2.f - 1.f;
  const float yn = py / (float)(h - 1) * 2.f - 1.f;
  float ix, iy;
  if (align) {
    ix = ((xn + 1.f) / 2.f) * (float)(fw - 1);
    iy = ((yn + 1.f) / 2.f) * (float)(fh - 1);
  } else {
    ix = ((xn + 1.f) * (float)fw - 1.f) / 2.f;
    iy = ((yn + 1.f) * (float)fh - 1.f) / 2.f;
  }
  ix = fminf(fmaxf(ix, 0.f), (float)(fw - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(fh - 1));
  const float x0 = floorf(ix), y0 = floorf(iy);
  Footprint f;
  const int xi = (int)x0, yi = (int)y0;
  f.tx = ix - x0;
  f.ty = iy - y0;
  f.dx = (xi + 1 < fw) ? 1 : 0;
  f.dy = (yi + 1 < fh) ? 1 : 0;
  f.off = yi * fw + xi;
  return f;
}

// same footprint with the two divisions by (w-1), (h-1) replaced by multiplications with reciprocals the caller hoisted
// (bf16 path only: positions differ from the reference's by ~1 ulp)
__device__ __forceinline__ Footprint border_footprint_r(float px, float py, float inv_wm1, float inv_hm1, bool align, int fh, int fw) {
  const float xn = px * inv_wm1 * 2.f - 1.f;
  const float yn = py * inv_hm1 * 2.f - 1.f;
  float ix, iy;
  if (align) {
    ix = ((xn + 1.f) * 0.5f) * (float)(fw - 1);
    iy = ((yn + 1.f) * 0.5f) * (float)(fh - 1);
  } else {
    ix = ((xn + 1.f) * (float)fw - 1.f) * 0.5f;
    iy = ((yn + 1.f) * (float)fh - 1.f) * 0.5f;
  }
  ix = fminf(fmaxf(ix, 0.f), (float)(fw - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(fh - 1));
  const float x0 = floorf(ix), y0 = floorf(iy);
  Footprint f;
  const int xi = (int)x0, yi = (int)y0;
  f.tx = ix - x0;
  f.ty = iy - y0;
  f.dx = (xi + 1 < fw) ? 1 : 0;
  f.dy = (yi + 1 < fh) ? 1 : 0;
  f.off = yi * fw + xi;
  return f;
}

// ---- shared by the fp32 and the bf16 render kernels ----
__device__ __forceinline__ float4 tap4(const float4* __restrict__ base, const Footprint& f, int stride_x, int stride_y) {
  // ATen order: nw, ne, sw, se
  const float4 nw = ldg4(base);
  const float4 ne = ldg4(base + f.dx * stride_x);
  const float4 sw = ldg4(base + f.dy * stride_y);
  const float4 se = ldg4(base + f.dy * stride_y + f.dx * stride_x);
  const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
  const float wnw = tx1 * ty1, wne = f.tx * ty1, wsw = tx1 * f.ty, wse = f.tx * f.ty;
  float4 o;
  o.x = nw.x * wnw; o.y = nw.y * wnw; o.z = nw.z * wnw; o.w = nw.w * wnw;
  o.x = fmaf(ne.x, wne, o.x); o.y = fmaf(ne.y, wne, o.y); o.z = fmaf(ne.z, wne, o.z); o.w = fmaf(ne.w, wne, o.w);
  o.x = fmaf(sw.x, wsw, o.x); o.y = fmaf(sw.y, wsw, o.y); o.z = fmaf(sw.z, wsw, o.z); o.w = fmaf(sw.w, wsw, o.w);
  o.x = fmaf(se.x, wse, o.x); o.y = fmaf(se.y, wse, o.y); o.z = fmaf(se.z, wse, o.z); o.w = fmaf(se.w, wse, o.w);
  return o;
}

// Geometry of one (view, sample) row. Outputs projected pixel/depth, projection direction and the
// (dir - que_dir, dot) feature of aggregate_net.get_dir_diff.
struct RowGeom {
  float px, py, pdepth;
  float dir[3];
  float dirdiff[4];
};
template <bool FAST = false>   // FAST (bf16 path): the two normalisations use rsqrt.approx instead of sqrt + 3 IEEE divisions each
__device__ __forceinline__ RowGeom row_geometry(const pgrf_render_args& a, int v, long long g) {
  int ray, s;
  if (FAST) {   // bf16 path: sample indices of one launch fit 32 bits (checked by the launcher): no 64-bit division
    ray = (int)((unsigned)g / (unsigned)a.dn); s = (int)((unsigned)g - (unsigned)ray * (unsigned)a.dn);
  } else {
    ray = (int)(g / a.dn); s = (int)(g % a.dn);
  }
  const float depth = __ldg(a.depth + (size_t)ray * a.depth_ray_stride + s);
  const float* c = a.que_c2w;  // (3,4) row-major
  float rdx, rdy, rdz;
  if (a.ray_dirs) {   // perspective / cube rays: the caller's world-space directions (render_ops.py:37-74)
    rdx = __ldg(a.ray_dirs + 3 * (size_t)ray); rdy = __ldg(a.ray_dirs + 3 * (size_t)ray + 1); rdz = __ldg(a.ray_dirs + 3 * (size_t)ray + 2);
  } else {
    const float cx = __ldg(a.coords + 2 * (size_t)ray), cy = __ldg(a.coords + 2 * (size_t)ray + 1);
    float dx, dy, dz;
    // `.long()` truncation of the pixel coordinate (render_ops.py:96-97)
    equi_unit_dir<FAST>(a.dataset, (float)(long long)cx, (float)(long long)cy, a.H, a.W, dx, dy, dz);
    rdx = c[0] * dx + c[1] * dy + c[2] * dz;
    rdy = c[4] * dx + c[5] * dy + c[6] * dz;
    rdz = c[8] * dx + c[9] * dy + c[10] * dz;
  }
  const float p0 = c[3] + rdx * depth, p1 = c[7] + rdy * depth, p2 = c[11] + rdz * depth;
  float q0, q1, q2;   // que_dir
  if (FAST) {
    const float ir = rsqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
    q0 = -rdx * ir; q1 = -rdy * ir; q2 = -rdz * ir;
  } else {
    const float rn = sqrtf(rdx * rdx + rdy * rdy + rdz * rdz);
    q0 = -rdx / rn; q1 = -rdy / rn; q2 = -rdz / rn;
  }
  const float* w = a.ref_w2c + 12 * v;
  const float pc0 = w[0] * p0 + w[1] * p1 + w[2] * p2 + w[3];
  const float pc1 = w[4] * p0 + w[5] * p1 + w[6] * p2 + w[7];
  const float pc2 = w[8] * p0 + w[9] * p1 + w[10] * p2 + w[11];
  RowGeom r;
  cam_to_equi<FAST>(a.dataset, pc0, pc1, pc2, a.H, a.W, r.pdepth, r.px, r.py);
  // camera centre -R^T t (render_ops.py:204), direction from the point to the source camera
  const float cam0 = -(w[0] * w[3] + w[4] * w[7] + w[8] * w[11]);
  const float cam1 = -(w[1] * w[3] + w[5] * w[7] + w[9] * w[11]);
  const float cam2 = -(w[2] * w[3] + w[6] * w[7] + w[10] * w[11]);
  const float e0 = p0 - cam0, e1 = p1 - cam1, e2 = p2 - cam2;
  if (FAST) {
    const float ie = fminf(rsqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e5f);     // == 1 / max(|e|, 1e-5)
    r.dir[0] = -e0 * ie; r.dir[1] = -e1 * ie; r.dir[2] = -e2 * ie;
  } else {
    const float en = fmaxf(sqrtf(e0 * e0 + e1 * e1 + e2 * e2), 1e-5f);
    r.dir[0] = -e0 / en; r.dir[1] = -e1 / en; r.dir[2] = -e2 / en;
  }
  r.dirdiff[0] = r.dir[0] - q0; r.dirdiff[1] = r.dir[1] - q1; r.dirdiff[2] = r.dir[2] - q2;
  r.dirdiff[3] = r.dir[0] * q0 + r.dir[1] * q1 + r.dir[2] * q2;
  return r;
}

// normalised inverse depth of dist_decoder.get_near_far_points / render_ops.depth2inv_dists
__device__ __forceinline__ float inv_norm(float depth, float near, float far) {
  const float nn = -1.f / near, ff = -1.f / far;
  return (-1.f / depth - nn) / (ff - nn);
}

}  // namespace pgrf

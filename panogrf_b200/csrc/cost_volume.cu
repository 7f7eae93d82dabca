// K1 — fused 360° spherical-sweep cost volume for sm_100a.
//
// Replaces the reference's per-depth python loop
//   models/spherical_cost_volume.py:231-341 (calculate_cost_volume_erp), :135-230 (get_cv_per_depth)
//   models/spherical_cost_volume_mv.py:219-347 (multi-view mean)
//   helpers/my_torch_helpers.py:12-120 (spherical <-> cartesian)
//   network/omni_mvsnet/pipeline3_model.py:849-853 (group-wise mean, fused as an epilogue)
// with ONE launch that never materialises the warped (B,D,C,H,W) source volume.
//
// Work decomposition (HBM-write bound: 4*C bytes/voxel out, features stay L2/L1 resident):
//   CTA  = 4 warps = one ERP row segment of 128 pixels x a chunk of depth hypotheses
//   warp = 32 consecutive pixels, loops over the depth chunk
//   phase A (lane <-> pixel): ray * depth -> rigid transform -> (theta,phi) -> uv -> tap record in smem
//   phase B (lane <-> (pixel, float4 channel group)): 4 x 128-bit gathers per source view (each
//            bilinear tap of a pixel is one contiguous C*4-byte line of the channels-last map),
//            cost vs. the reference feature held in registers, accumulation over source views
//   epilogue: channels-last -> direct 128-bit streaming stores (512 B contiguous per warp store);
//             planar (B,D,C,H,W)/(B,C,D,H,W)/group-mean -> conflict-free smem transpose, 128-B row stores
#include <cuda.h>

#include "cost_volume.cuh"

namespace pgrf {

template <int C, bool PLANAR, bool SINGLE, int JBT, int MINB>
__global__ void __launch_bounds__(kCvThreads, MINB) cost_volume_kernel(const CvParams p) {
  constexpr int CG = C / 4;      // lanes (float4 channel groups) per pixel
  constexpr int PPS = 32 / CG;   // pixels per sub-iteration of phase B
  constexpr int NSUB = 32 / PPS; // sub-iterations to cover the warp's 32 pixels

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_A[kMaxSrc][12];  // per source view: A = R_s * R_ref^-1 (row-major 3x3), b = t_s - A t_ref

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (p.W + kCvThreads - 1) / kCvThreads;
  const int y = blockIdx.x / tiles_x;
  const int x_warp = (blockIdx.x % tiles_x) * kCvThreads + warp * 32;
  const int b = blockIdx.z;
  const int d_begin = blockIdx.y * p.d_chunk;
  const int d_end = min(p.D, d_begin + p.d_chunk);

  TapRec* rec = reinterpret_cast<TapRec*>(smem_raw) + warp * (p.n_src * 32);
  float* tile = reinterpret_cast<float*>(smem_raw + (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec)) +
                warp * (C * 33);

  if (threadIdx.x < p.n_src) relative_pose(p, b, threadIdx.x, s_A[threadIdx.x]);
  __syncthreads();

  const int x = x_warp + lane;
  float rx, ry, rz;
  pixel_ray(p, min(x, p.W - 1), y, rx, ry, rz);

  // phase-B coordinates of this lane
  const int pp = lane / CG, cg = lane % CG;
  const size_t view_f4 = (size_t)p.H * p.W * CG;  // float4 per (H,W,C) view
  const float4* img4 = reinterpret_cast<const float4*>(p.images) + (size_t)b * p.S * view_f4;

  // reference features of the warp's 32 pixels stay in registers for the whole depth chunk
  float4 ref[NSUB];
#pragma unroll
  for (int j = 0; j < NSUB; ++j) {
    const int px = x_warp + j * PPS + pp;
    ref[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < p.W) ref[j] = ldg4(img4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + px) * CG + cg);
  }
  if (p.cost_type == PGRF_COST_ABS_DIFF) {     // |warped - ref| = |warped + (-ref)|: keep the negated feature
#pragma unroll
    for (int j = 0; j < NSUB; ++j) ref[j] = make_float4(-ref[j].x, -ref[j].y, -ref[j].z, -ref[j].w);
  }

  const float4* lane_base = img4 + cg;
  const int row_f4 = p.W * CG;
  const bool use_div = p.divisor != 0.f;
  // single swept view: the rotated ray is depth-invariant
  const float hax = s_A[0][0] * rx + s_A[0][1] * ry + s_A[0][2] * rz;
  const float hay = s_A[0][3] * rx + s_A[0][4] * ry + s_A[0][5] * rz;
  const float haz = s_A[0][6] * rx + s_A[0][7] * ry + s_A[0][8] * rz;
  const float hbx = s_A[0][9], hby = s_A[0][10], hbz = s_A[0][11];
  const size_t plane = (size_t)p.H * p.W;
  bool bad = false;

  for (int d = d_begin; d < d_end; ++d) {
    // ---------------- phase A: one lane per pixel ----------------
    float depth;
    if (p.depth_volume) depth = (x < p.W) ? __ldg(p.depth_volume + ((size_t)b * p.D + d) * plane + (size_t)y * p.W + x) : 1.f;
    else depth = __ldg(p.depths + d);
    for (int s = 0; s < (SINGLE ? 1 : p.n_src); ++s) {
      float cx, cy, cz;
      if (SINGLE) {
        cx = fmaf(depth, hax, hbx); cy = fmaf(depth, hay, hby); cz = fmaf(depth, haz, hbz);
      } else {
        const float* A = s_A[s];
        const float ax = A[0] * rx + A[1] * ry + A[2] * rz;
        const float ay = A[3] * rx + A[4] * ry + A[5] * rz;
        const float az = A[6] * rx + A[7] * ry + A[8] * rz;
        cx = fmaf(depth, ax, A[9]); cy = fmaf(depth, ay, A[10]); cz = fmaf(depth, az, A[11]);
      }
      float u, v;
      point_uv(p.dataset, cx, cy, cz, u, v);
      if (!(u >= -1.f && u <= 1.f && v >= -1.f && v <= 1.f) && x < p.W) bad = true;
      // grid_sample(align_corners=True) un-normalisation, ATen grid_sampler_unnormalize
      const float ix = ((u + 1.f) / 2.f) * (float)(p.W - 1);
      const float iy = ((v + 1.f) / 2.f) * (float)(p.H - 1);
      float x0f = floorf(ix), y0f = floorf(iy);
      x0f = fminf(fmaxf(x0f, 0.f), (float)(p.W - 2));   // NaN -> 0
      y0f = fminf(fmaxf(y0f, 0.f), (float)(p.H - 2));
      TapRec r;
      r.tx = ix - x0f;
      r.ty = iy - y0f;
      r.off4 = ((int)y0f * p.W + (int)x0f) * CG;
      r.pad = 0;
      rec[s * 32 + lane] = r;
    }
    __syncwarp();

    // ---------------- phase B: one lane per (pixel, float4 of channels) ----------------
    // Sub-iterations are processed JB at a time with ALL their gathers issued before the first use, so a
    // lane keeps 4*JB independent 128-bit loads in flight (the kernel is latency bound on these L1/L2 hits).
    constexpr int JB = (NSUB >= JBT) ? JBT : NSUB;
    const size_t cl_idx = ((((size_t)b * p.D + d) * p.H + y) * p.W + x_warp) * CG + lane;      // 4-channel group index of this lane
    float4* out_cl = reinterpret_cast<float4*>(p.out) + cl_idx;
    uint2* out_cl16 = reinterpret_cast<uint2*>(p.out) + cl_idx;                                // bf16 channels-last variant
#pragma unroll
    for (int j0 = 0; j0 < NSUB; j0 += JB) {
      float4 acc[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < (SINGLE ? 1 : p.n_src); ++s) {
        const float4* vbase = lane_base + (size_t)p.src_views[s] * view_f4;
        float4 t[JB][4];
        float tx[JB], ty[JB];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const TapRec r = rec[s * 32 + (j0 + jj) * PPS + pp];
          const float4* row0 = vbase + r.off4;
          const float4* row1 = row0 + row_f4;
          t[jj][0] = ldg4(row0); t[jj][1] = ldg4(row0 + CG); t[jj][2] = ldg4(row1); t[jj][3] = ldg4(row1 + CG);
          tx[jj] = r.tx; ty[jj] = r.ty;
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          float4 val = blend_cost(t[jj][0], t[jj][1], t[jj][2], t[jj][3], tx[jj], ty[jj], ref[j0 + jj], p.cost_type);
          if (SINGLE) { acc[jj] = val; continue; }          // one swept view, no divisor: the sum is the term itself
          if (use_div) { val.x = val.x / p.divisor; val.y = val.y / p.divisor; val.z = val.z / p.divisor; val.w = val.w / p.divisor; }
          acc[jj].x += val.x; acc[jj].y += val.y; acc[jj].z += val.z; acc[jj].w += val.w;
        }
      }
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        const int pi = (j0 + jj) * PPS + pp;
        if (!PLANAR) {
          // lane (pp,cg) of sub-iteration j writes float4 index j*32 + lane of the warp's 32*CG contiguous float4
          if (x_warp + pi < p.W) {
            if (p.out_bf16) {   // what the tensor-core regulariser consumes: half the bytes, no conversion pass
              uint2 q;
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.x) : "f"(acc[jj].y), "f"(acc[jj].x));
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q.y) : "f"(acc[jj].w), "f"(acc[jj].z));
              __stcs(out_cl16 + (j0 + jj) * 32, q);
            } else {
              stcs4(out_cl + (j0 + jj) * 32, acc[jj]);
            }
          }
        } else {
          // tile[c][pixel] with row pitch 33: bank = (4cg + k + 4j + pp) mod 32 -> conflict-free
          tile[(cg * 4 + 0) * 33 + pi] = acc[jj].x;
          tile[(cg * 4 + 1) * 33 + pi] = acc[jj].y;
          tile[(cg * 4 + 2) * 33 + pi] = acc[jj].z;
          tile[(cg * 4 + 3) * 33 + pi] = acc[jj].w;
        }
      }
    }
    if (PLANAR) {
      __syncwarp();
      if (x < p.W) {
        float* dst = p.out + (size_t)b * p.sB + (size_t)d * p.sD + (size_t)y * p.W + x;
        if (p.groups > 0) {
          const int cpg = C / p.groups;
          const float inv_n = (float)cpg;
          for (int g = 0; g < p.groups; ++g) {
            float sum = 0.f;
            for (int k = 0; k < cpg; ++k) sum += tile[(g * cpg + k) * 33 + lane];
            __stcs(dst + (size_t)g * p.sC, sum / inv_n);
          }
        } else {
          float* q = dst;
#pragma unroll 8
          for (int c = 0; c < C; ++c) { __stcs(q, tile[c * 33 + lane]); q += p.sC; }
        }
      }
    }
    __syncwarp();
  }
  if (bad) atomicOr(p.err, 1);
}


// ---------------------------------------------------------------------------------------------------------------------
// Planar layouts, C = 32, W % 128 == 0: the output leaves through the TMA engine (tensor-map store) instead of the LSU.
// The L1TEX data stage is this kernel's limiter (profiles/r1_final_ncu_summary.md: 89 % busy, DRAM 39 %): per voxel column it
// serves 4 x 128 B of bilinear taps, and in the version above another 32 STS + 32 LDS + 32 STG wavefronts for the
// (pixel, channel) -> (channel, pixel) transposition.  Here
//   * the gather keeps its lane = (pixel pp, channel group cg) mapping (a quarter-warp of a 128-bit load = one 128-byte line);
//     a 4x4 register transpose over the lanes pp = 0..3 of a channel group (4 SHFL) turns "4 channels of 1 pixel" into
//     "1 channel of 4 pixels";
//   * one STS.128 per sub-iteration writes it into the warp's own [32 rows][32 pixels] tile, row = pp * 8 + cg (channel
//     4 cg + pp), in the 128-byte swizzle of the tensor map (16-byte chunk index XOR row % 8 = cg): conflict-free;
//   * one lane hands the whole 4 KB tile to the TMA engine (cp.async.bulk.tensor.5d): the output is described as
//     (pixel, cg, pp, depth, batch) with channel = 4 cg + pp split over two dimensions, box 32 x 8 x 4 x 1 x 1, so the row
//     permutation costs nothing; double-buffered over the depth loop — warp-local, no CTA barrier.
// Measured dead ends: 512-byte cp.async.bulk row copies (2.1 M requests per volume: 0.62 ms, request bound); pixel-fastest
// lane mapping for an un-permuted tile (every quarter-warp of a tap load then touches 4 lines instead of 1: 0.55 ms).
// L1TEX wavefronts per voxel column: 128 (taps) + 32 (STS.128) instead of 128 + 96.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kCvWarpTileBytes = 32 * 32 * 4;           // one warp's [32 channels][32 pixels] fp32 tile

template <bool SINGLE, int JBT, int MINB>
__global__ void __launch_bounds__(kCvThreads, MINB) cost_volume_planar_tma_kernel(const CvParams p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int CG = 8, PPS = 4, NSUB = 8;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ float s_A[kMaxSrc][12];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = p.W / kCvThreads;
  const int y = blockIdx.x / tiles_x;
  const int x_warp = (blockIdx.x % tiles_x) * kCvThreads + warp * 32;
  const int b = blockIdx.z;
  const int d_begin = blockIdx.y * p.d_chunk;
  const int d_end = min(p.D, d_begin + p.d_chunk);

  unsigned char* tiles = smem_raw + (size_t)warp * 2 * kCvWarpTileBytes;                 // this warp's two tiles (1024 B aligned)
  TapRec* rec = reinterpret_cast<TapRec*>(smem_raw + (size_t)kCvWarps * 2 * kCvWarpTileBytes) + warp * (p.n_src * 32);

  if (threadIdx.x < p.n_src) relative_pose(p, b, threadIdx.x, s_A[threadIdx.x]);
  __syncthreads();

  const int x = x_warp + lane;
  float rx, ry, rz;
  pixel_ray(p, x, y, rx, ry, rz);

  const int pp = lane >> 3, cg = lane & 7;
  const size_t view_f4 = (size_t)p.H * p.W * CG;
  const float4* img4 = reinterpret_cast<const float4*>(p.images) + (size_t)b * p.S * view_f4;
  float4 ref[NSUB];
#pragma unroll
  for (int j = 0; j < NSUB; ++j)
    ref[j] = ldg4(img4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + x_warp + j * PPS + pp) * CG + cg);
  if (p.cost_type == PGRF_COST_ABS_DIFF) {     // |warped - ref| = |warped + (-ref)|: keep the negated feature
#pragma unroll
    for (int j = 0; j < NSUB; ++j) ref[j] = make_float4(-ref[j].x, -ref[j].y, -ref[j].z, -ref[j].w);
  }

  const float4* lane_base = img4 + cg;
  const int row_f4 = p.W * CG;
  const bool use_div = p.divisor != 0.f;
  const float hax = s_A[0][0] * rx + s_A[0][1] * ry + s_A[0][2] * rz;
  const float hay = s_A[0][3] * rx + s_A[0][4] * ry + s_A[0][5] * rz;
  const float haz = s_A[0][6] * rx + s_A[0][7] * ry + s_A[0][8] * rz;
  const float hbx = s_A[0][9], hby = s_A[0][10], hbz = s_A[0][11];
  const size_t plane = (size_t)p.H * p.W;
  bool bad = false;
  // after the 4x4 transpose this lane holds channel 4cg + pp of pixels 4j .. 4j+3: tile row pp*8 + cg, 16-byte chunk j ^ cg
  const int crow = pp * 8 + cg;
  uint64_t l2_evict_first;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_evict_first));
  const int pix0 = y * p.W + x_warp;                     // coordinate 0 of the warp's box

  for (int d = d_begin; d < d_end; ++d) {
    unsigned char* tile = tiles + ((d - d_begin) & 1) * kCvWarpTileBytes;
    // the store issued two depths ago read this buffer: at most the most recent one may still be reading
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    // ---------------- phase A: one lane per pixel ----------------
    float depth;
    if (p.depth_volume) depth = __ldg(p.depth_volume + ((size_t)b * p.D + d) * plane + (size_t)y * p.W + x);
    else depth = __ldg(p.depths + d);
    for (int s = 0; s < (SINGLE ? 1 : p.n_src); ++s) {
      float cx, cy, cz;
      if (SINGLE) {
        cx = fmaf(depth, hax, hbx); cy = fmaf(depth, hay, hby); cz = fmaf(depth, haz, hbz);
      } else {
        const float* A = s_A[s];
        const float ax = A[0] * rx + A[1] * ry + A[2] * rz;
        const float ay = A[3] * rx + A[4] * ry + A[5] * rz;
        const float az = A[6] * rx + A[7] * ry + A[8] * rz;
        cx = fmaf(depth, ax, A[9]); cy = fmaf(depth, ay, A[10]); cz = fmaf(depth, az, A[11]);
      }
      float u, v;
      point_uv(p.dataset, cx, cy, cz, u, v);
      if (!(u >= -1.f && u <= 1.f && v >= -1.f && v <= 1.f)) bad = true;
      const float ix = ((u + 1.f) / 2.f) * (float)(p.W - 1);
      const float iy = ((v + 1.f) / 2.f) * (float)(p.H - 1);
      float x0f = floorf(ix), y0f = floorf(iy);
      x0f = fminf(fmaxf(x0f, 0.f), (float)(p.W - 2));   // NaN -> 0
      y0f = fminf(fmaxf(y0f, 0.f), (float)(p.H - 2));
      TapRec r;
      r.tx = ix - x0f;
      r.ty = iy - y0f;
      r.off4 = ((int)y0f * p.W + (int)x0f) * CG;
      r.pad = 0;
      rec[s * 32 + lane] = r;
    }
    __syncwarp();

    // ---------------- phase B: one lane per (pixel, float4 of channels) ----------------
    constexpr int JB = JBT;
#pragma unroll
    for (int j0 = 0; j0 < NSUB; j0 += JB) {
      float4 acc[JB];
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) acc[jj] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < (SINGLE ? 1 : p.n_src); ++s) {
        const float4* vbase = lane_base + (size_t)p.src_views[s] * view_f4;
        float4 t[JB][4];
        float tx[JB], ty[JB];
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          const TapRec r = rec[s * 32 + (j0 + jj) * PPS + pp];
          const float4* row0 = vbase + r.off4;
          const float4* row1 = row0 + row_f4;
          t[jj][0] = ldg4(row0); t[jj][1] = ldg4(row0 + CG); t[jj][2] = ldg4(row1); t[jj][3] = ldg4(row1 + CG);
          tx[jj] = r.tx; ty[jj] = r.ty;
        }
#pragma unroll
        for (int jj = 0; jj < JB; ++jj) {
          float4 val = blend_cost(t[jj][0], t[jj][1], t[jj][2], t[jj][3], tx[jj], ty[jj], ref[j0 + jj], p.cost_type);
          if (SINGLE) { acc[jj] = val; continue; }          // one swept view, no divisor: the sum is the term itself
          if (use_div) { val.x = val.x / p.divisor; val.y = val.y / p.divisor; val.z = val.z / p.divisor; val.w = val.w / p.divisor; }
          acc[jj].x += val.x; acc[jj].y += val.y; acc[jj].z += val.z; acc[jj].w += val.w;
        }
      }
#pragma unroll
      for (int jj = 0; jj < JB; ++jj) {
        // 4x4 transpose over the lanes pp = 0..3 of one channel group: (4 channels of pixel pp) -> (channel pp of 4 pixels)
        float4 v = acc[jj];
        {
          const bool odd = pp & 1;
          const float s0 = odd ? v.x : v.y, s1 = odd ? v.z : v.w;
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 8), r1 = __shfl_xor_sync(0xffffffffu, s1, 8);
          v = odd ? make_float4(r0, v.y, r1, v.w) : make_float4(v.x, r0, v.z, r1);
        }
        {
          const bool up = pp & 2;
          const float s0 = up ? v.x : v.z, s1 = up ? v.y : v.w;
          const float r0 = __shfl_xor_sync(0xffffffffu, s0, 16), r1 = __shfl_xor_sync(0xffffffffu, s1, 16);
          v = up ? make_float4(r0, r1, v.z, v.w) : make_float4(v.x, v.y, r0, r1);
        }
        *reinterpret_cast<float4*>(tile + crow * 128 + (((j0 + jj) ^ (crow & 7)) << 4)) = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      // evict-first: the volume is written once and never re-read here; without the hint the stores push the feature maps out
      // of L2 (ncu: DRAM reads 67 MB -> 161 MB, long_scoreboard on the tap loads)
      asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4, %5}], [%6], %7;" ::"l"(&tmap),
                   "r"(pix0), "r"(0), "r"(0), "r"(d), "r"(b), "r"((uint32_t)__cvta_generic_to_shared(tile)), "l"(l2_evict_first)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (bad) atomicOr(p.err, 1);
}

// tensor map of the planar output as (pixel = y*W + x, cg, pp, depth, batch) with channel = 4 cg + pp; box 32 x 8 x 4 x 1 x 1,
// 128-byte swizzle
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_out_tensor_map(const CvParams& p, CUtensorMap* tm) {
  static PFN_encodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    PGRF_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    PGRF_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cost_volume: cuTensorMapEncodeTiled unavailable");
    encode = (PFN_encodeTiled)fn;
  }
  const cuuint64_t dims[5] = {(cuuint64_t)p.H * p.W, 8, 4, (cuuint64_t)p.D, (cuuint64_t)p.B};
  const cuuint64_t strides[4] = {(cuuint64_t)p.sC * 16, (cuuint64_t)p.sC * 4, (cuuint64_t)p.sD * 4, (cuuint64_t)p.sB * 4};   // bytes, dims 1..4
  const cuuint32_t box[5] = {32, 8, 4, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)p.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PGRF_REQUIRE(r == CUDA_SUCCESS, "cost_volume: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return PGRF_OK;
}

int g_cv_jb = 0;      // gathers batched per lane (4*JB 128-bit loads in flight); 0 = measured default per layout
int g_cv_dchunk = 0;  // 0 = heuristic
int g_cv_pad_smem = 0;
int g_cv_tma = 0;     // 1 = planar layouts (C = 32, W % 128 == 0) leave through the TMA engine; measured 2 % slower than the LSU path
                      // on B200 although it takes L1TEX from 89 % to 69 %: opt-in (debug knob cv_tma)
int g_cv_minb = 0;    // __launch_bounds__ min blocks per SM: 0 = default (5 blocks = 20 warps/SM), 1 = uncapped registers

template <int C>
static int launch_c(const CvParams& p, int layout, cudaStream_t st) {
  const int tiles_x = (p.W + kCvThreads - 1) / kCvThreads;
  const int n_chunks = (p.D + p.d_chunk - 1) / p.d_chunk;
  dim3 grid((unsigned)(tiles_x * p.H), (unsigned)n_chunks, (unsigned)p.B);
  const bool planar = layout != PGRF_CV_BDHWC;
  size_t smem = (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec);
  if (planar) smem += (size_t)kCvWarps * C * 33 * sizeof(float);
  smem += (size_t)g_cv_pad_smem;       // experiment knob: shrink the L1 carve-out without touching the kernel
  const bool single = p.n_src == 1 && p.divisor == 0.f;
#define PGRF_CV_LAUNCH1(PL, SG, J, MB)                                                                                \
  do {                                                                                                                \
    PGRF_CUDA(cudaFuncSetAttribute(cost_volume_kernel<C, PL, SG, J, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cost_volume_kernel<C, PL, SG, J, MB><<<grid, kCvThreads, smem, st>>>(p);                                          \
  } while (0)
#define PGRF_CV_LAUNCH(PL, SG, J)                                                                                     \
  do {                                                                                                                \
    if (g_cv_minb == 1) PGRF_CV_LAUNCH1(PL, SG, J, 1);                                                                \
    else PGRF_CV_LAUNCH1(PL, SG, J, 5);                                                                               \
  } while (0)
  const int jb = g_cv_jb ? g_cv_jb : 2;   // B200 sweep (tools/time_cost_volume.py): 8 gathers in flight per lane, <= 102 registers
  if (planar && C == 32 && p.groups == 0 && p.W % kCvThreads == 0 && g_cv_tma) {
    // planar store through the TMA engine (see cost_volume_planar_tma_kernel)
    const size_t smem_t = (size_t)kCvWarps * 2 * kCvWarpTileBytes + (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec);
    CUtensorMap tmap;
    const int trc = make_out_tensor_map(p, &tmap);
    if (trc != PGRF_OK) return trc;
#define PGRF_CV_TMA(SG, J, MB)                                                                                        \
  do {                                                                                                                \
    PGRF_CUDA(cudaFuncSetAttribute(cost_volume_planar_tma_kernel<SG, J, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
    cost_volume_planar_tma_kernel<SG, J, MB><<<grid, kCvThreads, smem_t, st>>>(p, tmap);                              \
  } while (0)
    if (!single) PGRF_CV_TMA(false, 2, 5);
    else if (jb == 4) PGRF_CV_TMA(true, 4, 4);
    else if (g_cv_minb == 6) PGRF_CV_TMA(true, 2, 6);
    else if (g_cv_minb == 4) PGRF_CV_TMA(true, 2, 4);
    else PGRF_CV_TMA(true, 2, 5);
#undef PGRF_CV_TMA
    count_launch();
    PGRF_CUDA(cudaGetLastError());
    return PGRF_OK;
  }
  if (planar) {
    if (!single) PGRF_CV_LAUNCH(true, false, 2);
    else if (jb == 4) PGRF_CV_LAUNCH(true, true, 4);
    else PGRF_CV_LAUNCH(true, true, 2);
  } else {
    if (!single) PGRF_CV_LAUNCH(false, false, 2);
    else if (jb == 4) PGRF_CV_LAUNCH(false, true, 4);
    else PGRF_CV_LAUNCH(false, true, 2);
  }
#undef PGRF_CV_LAUNCH
#undef PGRF_CV_LAUNCH1
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

}  // namespace pgrf

using namespace pgrf;

namespace pgrf { extern int g_mlp_wait_mode; extern int g_rays_tc; extern int g_pg_variant; extern int g_pg_grid; extern int g_conv_stages; extern int g_conv_row; extern int g_conv_splits; extern int g_conv_kc; extern int g_conv_smem_kb; extern int g_conv_persist; extern int g_conv_persist_kb; extern int g_cv_bwd_variant; extern int g_cv_bwd_run; extern int g_cv_bwd_chunks; extern int g_cv_bwd_nored; extern int g_cv_bwd_minb; extern int g_dg_prefilter; extern int g_dg_regsort; }
extern "C" int pgrf_debug_set(const char* key, int value) {
  if (!strcmp(key, "mlp_wait_mode")) { pgrf::g_mlp_wait_mode = value; return PGRF_OK; }
  if (!strcmp(key, "rays_tc")) { pgrf::g_rays_tc = value; return PGRF_OK; }
  if (!strcmp(key, "pg_variant")) { pgrf::g_pg_variant = value; return PGRF_OK; }
  if (!strcmp(key, "pg_grid")) { pgrf::g_pg_grid = value; return PGRF_OK; }
  if (!strcmp(key, "conv_stages")) { pgrf::g_conv_stages = value <= 0 ? 0 : (value < 2 ? 2 : (value > 6 ? 6 : value)); return PGRF_OK; }
  if (!strcmp(key, "conv_kc")) { pgrf::g_conv_kc = value >= 64 ? 64 : (value >= 32 ? 32 : 16); return PGRF_OK; }
  if (!strcmp(key, "conv_smem_kb")) { pgrf::g_conv_smem_kb = value; return PGRF_OK; }
  if (!strcmp(key, "conv_persist")) { pgrf::g_conv_persist = value; return PGRF_OK; }
  if (!strcmp(key, "conv_persist_kb")) { pgrf::g_conv_persist_kb = value; return PGRF_OK; }
  if (!strcmp(key, "conv_splits")) { pgrf::g_conv_splits = value; return PGRF_OK; }
  if (!strcmp(key, "conv_row")) { pgrf::g_conv_row = value; return PGRF_OK; }
  if (!strcmp(key, "cv_bwd_variant")) { pgrf::g_cv_bwd_variant = value; return PGRF_OK; }
  if (!strcmp(key, "cv_bwd_minb")) { pgrf::g_cv_bwd_minb = value; return PGRF_OK; }
  if (!strcmp(key, "cv_bwd_nored")) { pgrf::g_cv_bwd_nored = value; return PGRF_OK; }
  if (!strcmp(key, "cv_bwd_chunks")) { pgrf::g_cv_bwd_chunks = value; return PGRF_OK; }
  if (!strcmp(key, "cv_bwd_run")) { pgrf::g_cv_bwd_run = value; return PGRF_OK; }
  if (!strcmp(key, "dg_regsort")) { pgrf::g_dg_regsort = value; return PGRF_OK; }
  if (!strcmp(key, "dg_prefilter")) { pgrf::g_dg_prefilter = value; return PGRF_OK; }
  if (!strcmp(key, "cv_jb")) { g_cv_jb = value; return PGRF_OK; }
  if (!strcmp(key, "cv_dchunk")) { g_cv_dchunk = value; return PGRF_OK; }
  if (!strcmp(key, "cv_minb")) { g_cv_minb = value; return PGRF_OK; }
  if (!strcmp(key, "cv_tma")) { g_cv_tma = value; return PGRF_OK; }
  if (!strcmp(key, "cv_pad_smem")) { g_cv_pad_smem = value; return PGRF_OK; }
  set_error("pgrf_debug_set: unknown key %s", key);
  return PGRF_EINVAL;
}

namespace pgrf {
int cv_fill_params(CvParams& p, const float* images, int B, int S, int H, int W, int C, const float* depths, const float* depth_volume,
                   int D, const float* rots, const float* trans, int ref_idx, const int* src_views, int n_src, float divisor,
                   int dataset, int cost_type) {
  p.out_bf16 = 0;
  PGRF_REQUIRE(images && rots && trans && src_views, "cost_volume: null pointer argument");
  PGRF_REQUIRE(depths || depth_volume, "cost_volume: need depths or depth_volume");
  PGRF_REQUIRE(B > 0 && S > 0 && H > 1 && W > 1 && D > 0, "cost_volume: bad shape B=%d S=%d H=%d W=%d D=%d", B, S, H, W, D);
  PGRF_REQUIRE(B <= 65535, "cost_volume: B=%d exceeds grid.z", B);
  PGRF_REQUIRE(C == 4 || C == 8 || C == 16 || C == 32 || C == 64, "cost_volume: C=%d unsupported (need 4,8,16,32,64)", C);
  PGRF_REQUIRE(n_src >= 1 && n_src <= kMaxSrc, "cost_volume: n_src=%d out of [1,%d]", n_src, kMaxSrc);
  PGRF_REQUIRE(ref_idx >= 0 && ref_idx < S, "cost_volume: ref_idx=%d out of range", ref_idx);
  PGRF_REQUIRE(dataset >= 0 && dataset <= 3, "cost_volume: unknown dataset id %d", dataset);
  PGRF_REQUIRE(cost_type >= 0 && cost_type <= 2, "Unknown cost type");
  PGRF_REQUIRE((size_t)H * W * C < (1ull << 31), "cost_volume: one view exceeds 2^31 elements");
  PGRF_REQUIRE(((uintptr_t)images & 15) == 0, "cost_volume: images/out must be 16-byte aligned");
  memset(&p, 0, sizeof(p));
  p.images = images; p.depths = depths; p.depth_volume = depth_volume; p.rots = rots; p.trans = trans;
  p.B = B; p.S = S; p.H = H; p.W = W; p.D = D;
  p.ref_idx = ref_idx; p.n_src = n_src;
  for (int i = 0; i < n_src; ++i) {
    PGRF_REQUIRE(src_views[i] >= 0 && src_views[i] < S, "cost_volume: src view %d out of range", src_views[i]);
    p.src_views[i] = src_views[i];
  }
  p.divisor = divisor; p.dataset = dataset; p.cost_type = cost_type;
  switch (dataset) {
    case PGRF_DS_M3D: p.ang0 = (float)(PGRF_PI_D / H); p.ang1 = (float)(2 * PGRF_PI_D / W); break;
    case PGRF_DS_REPLICA_TEST: p.ang0 = (float)H; p.ang1 = (float)(2 * PGRF_PI_D / W); break;
    case PGRF_DS_RESIDENTIAL: p.ang0 = (float)(H - 1); p.ang1 = (float)(W - 1); break;
    default: p.ang0 = (float)(PGRF_PI_D / (H - 1)); p.ang1 = (float)(-2 * PGRF_PI_D / (W - 1)); break;
  }
  return PGRF_OK;
}
}  // namespace pgrf

extern "C" int pgrf_cost_volume_fwd(const float* images, int B, int S, int H, int W, int C,
                                    const float* depths, const float* depth_volume, int D,
                                    const float* rots, const float* trans,
                                    int ref_idx, const int* src_views, int n_src, float divisor,
                                    int dataset, int cost_type, int layout, int groups,
                                    float* out, int* err_flag, void* stream) {
  PGRF_REQUIRE(out && err_flag, "cost_volume: null pointer argument");
  PGRF_REQUIRE(layout >= 0 && layout <= 3, "cost_volume: unknown layout %d", layout);
  PGRF_REQUIRE(layout != PGRF_CV_BDHWC_BF16 || C % 16 == 0, "cost_volume: the bf16 channels-last layout needs C %% 16 == 0 (C=%d)", C);
  PGRF_REQUIRE(groups == 0 || (layout == PGRF_CV_BCDHW && groups > 0 && C % groups == 0),
               "cost_volume: groups=%d needs layout BCDHW and C %% groups == 0", groups);
  PGRF_REQUIRE(((uintptr_t)out & 15) == 0, "cost_volume: images/out must be 16-byte aligned");
  CvParams p;
  const int frc = cv_fill_params(p, images, B, S, H, W, C, depths, depth_volume, D, rots, trans, ref_idx, src_views, n_src, divisor,
                                 dataset, cost_type);
  if (frc != PGRF_OK) return frc;
  p.out = out; p.err = err_flag;
  p.out_bf16 = layout == PGRF_CV_BDHWC_BF16;
  if (p.out_bf16) layout = PGRF_CV_BDHWC;
  p.groups = groups;
  p.OC = groups > 0 ? groups : C;
  const long long HW = (long long)H * W;
  if (layout == PGRF_CV_BDCHW) { p.sC = HW; p.sD = HW * p.OC; p.sB = p.sD * D; }
  else { p.sD = HW; p.sC = HW * D; p.sB = p.sC * p.OC; }
  // depth chunking: enough CTAs for >= ~4 waves of 148 SMs x 8 resident CTAs, chunks of >= 4 depths
  const long long ctas_per_chunk = (long long)((W + kCvThreads - 1) / kCvThreads) * H * B;
  int n_chunks = (int)((148LL * 8 * 4 + ctas_per_chunk - 1) / ctas_per_chunk);
  if (n_chunks < 1) n_chunks = 1;
  int d_chunk = (D + n_chunks - 1) / n_chunks;
  if (d_chunk < 4) d_chunk = D < 4 ? D : 4;
  if (g_cv_dchunk > 0) d_chunk = g_cv_dchunk < D ? g_cv_dchunk : D;
  p.d_chunk = d_chunk;
  PGRF_REQUIRE((D + d_chunk - 1) / d_chunk <= 65535, "cost_volume: too many depth chunks");

  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 4: return launch_c<4>(p, layout, st);
    case 8: return launch_c<8>(p, layout, st);
    case 16: return launch_c<16>(p, layout, st);
    case 32: return launch_c<32>(p, layout, st);
    default: return launch_c<64>(p, layout, st);
  }
}

extern "C" int pgrf_cost_volume_host(const float* images, int B, int S, int H, int W, int C,
                                     const float* depths, const float* depth_volume, int D,
                                     const float* rots, const float* trans,
                                     int ref_idx, const int* src_views, int n_src, float divisor,
                                     int dataset, int cost_type, int layout, int groups, float* out) {
  PGRF_REQUIRE(images && rots && trans && out, "cost_volume_host: null pointer argument");
  PGRF_REQUIRE(groups >= 0 && (groups == 0 || C % groups == 0), "cost_volume_host: bad groups");
  const size_t n_img = (size_t)B * S * H * W * C, n_dv = (size_t)B * D * H * W;
  const size_t n_out = (size_t)B * D * H * W * (groups > 0 ? groups : C);
  float *d_img = nullptr, *d_depth = nullptr, *d_rots = nullptr, *d_trans = nullptr, *d_out = nullptr;
  int* d_err = nullptr;
  cudaStream_t st;
  PGRF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int rc = PGRF_OK;
  int h_err = 0;
  auto body = [&]() -> int {
    PGRF_CUDA(cudaMallocAsync(&d_img, n_img * 4, st));
    PGRF_CUDA(cudaMallocAsync(&d_depth, (depth_volume ? n_dv : (size_t)D) * 4, st));
    PGRF_CUDA(cudaMallocAsync(&d_rots, (size_t)B * S * 9 * 4, st));
    PGRF_CUDA(cudaMallocAsync(&d_trans, (size_t)B * S * 3 * 4, st));
    PGRF_CUDA(cudaMallocAsync(&d_out, n_out * 4, st));
    PGRF_CUDA(cudaMallocAsync(&d_err, 4, st));
    PGRF_CUDA(cudaMemsetAsync(d_err, 0, 4, st));
    PGRF_CUDA(cudaMemcpyAsync(d_img, images, n_img * 4, cudaMemcpyHostToDevice, st));
    PGRF_CUDA(cudaMemcpyAsync(d_depth, depth_volume ? depth_volume : depths, (depth_volume ? n_dv : (size_t)D) * 4,
                              cudaMemcpyHostToDevice, st));
    PGRF_CUDA(cudaMemcpyAsync(d_rots, rots, (size_t)B * S * 9 * 4, cudaMemcpyHostToDevice, st));
    PGRF_CUDA(cudaMemcpyAsync(d_trans, trans, (size_t)B * S * 3 * 4, cudaMemcpyHostToDevice, st));
    int r = pgrf_cost_volume_fwd(d_img, B, S, H, W, C, depth_volume ? nullptr : d_depth, depth_volume ? d_depth : nullptr, D,
                                 d_rots, d_trans, ref_idx, src_views, n_src, divisor, dataset, cost_type, layout, groups,
                                 d_out, d_err, st);
    if (r != PGRF_OK) return r;
    PGRF_CUDA(cudaMemcpyAsync(out, d_out, n_out * 4, cudaMemcpyDeviceToHost, st));
    PGRF_CUDA(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
    PGRF_CUDA(cudaStreamSynchronize(st));
    return PGRF_OK;
  };
  rc = body();
  cudaFreeAsync(d_img, st); cudaFreeAsync(d_depth, st); cudaFreeAsync(d_rots, st);
  cudaFreeAsync(d_trans, st); cudaFreeAsync(d_out, st); cudaFreeAsync(d_err, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  if (rc == PGRF_OK && h_err) {
    set_error("Wrong UV mapping, UV must be in [-1, 1]!");
    return PGRF_ERANGE;
  }
  return rc;
}

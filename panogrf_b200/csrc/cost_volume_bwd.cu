// K1 backward — gradient of the fused spherical-sweep cost volume w.r.t. the feature maps.
//
// The reference obtains it from autograd through grid_sample + abs/mul
// (models/spherical_cost_volume.py:135-230); depth hypotheses and poses carry no gradient
// (built under no_grad / detached, pipeline3_model.py:647,671).  Same decomposition as the forward kernel:
//   CTA = 4 warps = 128 pixels of one ERP row x a chunk of depth hypotheses
//   phase A (lane <-> pixel): recompute the bilinear footprint of every swept view (bit-identical to the forward)
//   phase B (lane <-> (pixel, float4 channel group)): read the upstream gradient (coalesced 128-bit), recompute the
//            warped feature where the cost needs it, scatter 4 x 128-bit vector atomics into the source view, keep the
//            reference-view gradient in shared memory (lane-private slots) over the depth chunk and add it once at the end.
// grad_images must be zeroed by the caller; float atomics make the summation order run-dependent (as in ATen's
// grid_sampler_2d_backward on CUDA).
// Measured at configs[0] (tools/time_cv_bwd.py, tools/cv_bwd_nored_probe.py): 0.98-1.01 ms = 0.63 ms of streaming + gathers + 0.35 ms of
// atomics (a build without the atomics); round 1 kept the reference-view state in registers (122 registers, 16 warps/SM): 1.19 ms.
// The run-merged variant below halves the atomics and measures the same 1.03 ms (its walk is a dependent chain with divergent
// footprint transitions); it wins only at C = 64 (0.92 vs 1.11 ms) and stays opt-in (debug knob cv_bwd_variant = 1).
#include "cost_volume.cuh"

namespace pgrf {

int g_cv_bwd_nored = 0;

struct CvBwdParams {
  CvParams f;
  const float* grad_out;   // (B,D,H,W,C) contiguous
  float* grad_images;      // (B,S,H,W,C)
  int no_red;              // timing experiments only (debug knob cv_bwd_nored): skip the atomics
};

__device__ __forceinline__ void red_add4_(float4* dst, const float4& v, int skip) { if (!skip) atomicAdd(dst, v); }
#define red_add4(dst, v) red_add4_(dst, v, q.no_red)

template <int C, int MINB>
__global__ void __launch_bounds__(kCvThreads, MINB) cost_volume_bwd_lean_kernel(const CvBwdParams q) {
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
  // reference-view gradient of the warp's 32 pixels: shared memory, every slot private to one lane (the register copy cost 64 registers
  // and with them half of the resident warps: the kernel waits on L2 latency, occupancy is what it needs)
  float4* gs = reinterpret_cast<float4*>(smem_raw + (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec)) + warp * (32 * CG) + lane;

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

#pragma unroll
  for (int j = 0; j < NSUB; ++j) gs[j * 32] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* ref4 = img4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + x_warp) * CG + lane;   // pixel j*PPS+pp, group cg = float4 j*32+lane

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
#pragma unroll 4
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
          const float4 rf = ldg4(ref4 + j * 32);
          float4 gr = gs[j * 32];
          if (p.cost_type == PGRF_COST_ABS_DIFF) {
            // torch.abs backward: grad * sign(x), sign(0) = 0
            const float sx = (float)((val.x > rf.x) - (val.x < rf.x)), sy = (float)((val.y > rf.y) - (val.y < rf.y));
            const float sz = (float)((val.z > rf.z) - (val.z < rf.z)), sw_ = (float)((val.w > rf.w) - (val.w < rf.w));
            dw.x = g.x * sx; dw.y = g.y * sy; dw.z = g.z * sz; dw.w = g.w * sw_;
            gr.x -= dw.x; gr.y -= dw.y; gr.z -= dw.z; gr.w -= dw.w;
          } else {  // dot
            dw.x = g.x * rf.x; dw.y = g.y * rf.y; dw.z = g.z * rf.z; dw.w = g.w * rf.w;
            gr.x = fmaf(g.x, val.x, gr.x); gr.y = fmaf(g.y, val.y, gr.y);
            gr.z = fmaf(g.z, val.z, gr.z); gr.w = fmaf(g.w, val.w, gr.w);
          }
          gs[j * 32] = gr;
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
      if (px < p.W) red_add4(gimg4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + px) * CG + cg, gs[j * 32]);
    }
  }
}


// ---- run-merged variant (the default) ---------------------------------------------------------------------------
// The one-thread-per-(pixel, channel group) kernel above issues 4 vector atomics per (voxel, channel group, view) and is bound
// by the L2's atomic throughput (ncu: long_scoreboard 72 % behind `red.v4.f32`).  Neighbouring pixels of an ERP row land on
// neighbouring texels of the swept view, so consecutive footprints share a column.  Here a lane owns a RUN of L consecutive
// pixels of the row (lane <-> (run, float4 channel group)) and keeps the gradient of the two texel columns of the current footprint
// in registers: when the footprint moves one texel along the row only the column that falls out of it is flushed (2 atomics
// instead of 4), when it stays it costs none; the texels needed to recompute the warped feature are reused the same way.
// The reference-view gradient of the CTA's pixels accumulates in shared memory over the depth chunk (each slot is private to a lane).
template <int C, int L, int MINB>
__global__ void __launch_bounds__(kCvThreads, MINB) cost_volume_bwd_run_kernel(const CvBwdParams q) {
  const CvParams& p = q.f;
  constexpr int CG = C / 4;
  constexpr int NRUN = 32 / CG;       // runs per warp
  constexpr int PXW = NRUN * L;       // pixels per warp
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_A[kMaxSrc][12];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_x = (p.W + kCvWarps * PXW - 1) / (kCvWarps * PXW);
  const int y = blockIdx.x / tiles_x;
  const int x_warp = (blockIdx.x % tiles_x) * (kCvWarps * PXW) + warp * PXW;
  const int b = blockIdx.z;
  const int d_begin = blockIdx.y * p.d_chunk;
  const int d_end = min(p.D, d_begin + p.d_chunk);
  const bool need_val = p.cost_type != PGRF_COST_NONE;

  // per-warp shared memory: upstream gradient of one depth, double buffered [2][PXW][CG] float4 | reference-view gradient [PXW][CG] float4 |
  // tap records [n_src][PXW] | ray [3][PXW]
  const size_t warp_bytes = (size_t)(need_val ? 3 : 2) * PXW * C * sizeof(float) + (size_t)p.n_src * PXW * sizeof(TapRec) + 3 * PXW * sizeof(float);
  unsigned char* wbase = smem_raw + warp * warp_bytes;
  float4* gbuf = reinterpret_cast<float4*>(wbase);
  float4* gs = gbuf + 2 * PXW * CG;
  TapRec* rec = reinterpret_cast<TapRec*>(gs + (need_val ? PXW * CG : 0));
  float* ray = reinterpret_cast<float*>(rec + p.n_src * PXW);

  const int run = lane / CG, cg = lane % CG;
  const int x_run = x_warp + run * L;
  const int n_valid = max(0, min(L, p.W - x_run));
  const size_t view_f4 = (size_t)p.H * p.W * CG;
  const float4* img4 = reinterpret_cast<const float4*>(p.images) + (size_t)b * p.S * view_f4;
  float4* gimg4 = reinterpret_cast<float4*>(q.grad_images) + (size_t)b * p.S * view_f4;
  const int row_f4 = p.W * CG;
  const size_t plane = (size_t)p.H * p.W;
  const float inv_div = p.divisor != 0.f ? 1.f / p.divisor : 1.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4* my_gs = gs + run * L * CG + cg;          // slots (pixel i of the run, this lane's channel group): private to the lane
  const float4* my_ref = img4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + x_run) * CG + cg;
  // the warp's PXW pixels are one contiguous stretch of grad_out per depth: 16-byte cp.async, zero-filled past the row end
  const float4* g_row = reinterpret_cast<const float4*>(q.grad_out) + (((size_t)b * p.D * p.H + y) * p.W + x_warp) * CG;
  const size_t g_dstride = (size_t)p.H * p.W * CG;
  const int g_valid = max(0, min(PXW, p.W - x_warp)) * CG;      // float4s of the stretch inside the row
  auto fetch_g = [&](int d, int buf) {
    const float4* src = g_row + (size_t)d * g_dstride;
    float4* dst = gbuf + buf * PXW * CG;
    for (int j = lane; j < PXW * CG; j += 32) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + j);
      const int n = j < g_valid ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(j < g_valid ? src + j : reinterpret_cast<const float4*>(q.grad_out)), "r"(n) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (threadIdx.x < p.n_src) relative_pose(p, b, threadIdx.x, s_A[threadIdx.x]);
  for (int i = lane; i < PXW; i += 32) {
    float rx, ry, rz;
    pixel_ray(p, min(x_warp + i, p.W - 1), y, rx, ry, rz);
    ray[i] = rx; ray[PXW + i] = ry; ray[2 * PXW + i] = rz;
  }
  if (need_val)
    for (int i = 0; i < L; ++i) gs[(run * L + i) * CG + cg] = zero4;
  if (d_begin < d_end) fetch_g(d_begin, 0);
  __syncthreads();

  for (int d = d_begin; d < d_end; ++d) {
    const int buf = (d - d_begin) & 1;
    if (d + 1 < d_end) fetch_g(d + 1, buf ^ 1);
    // phase A: footprints of the warp's PXW pixels in every swept view (lane <-> pixel)
    for (int i = lane; i < PXW; i += 32) {
      const int x = x_warp + i;
      float depth;
      if (p.depth_volume) depth = (x < p.W) ? __ldg(p.depth_volume + ((size_t)b * p.D + d) * plane + (size_t)y * p.W + x) : 1.f;
      else depth = __ldg(p.depths + d);
      const float rx = ray[i], ry = ray[PXW + i], rz = ray[2 * PXW + i];
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
        rec[s * PXW + i] = r;
      }
    }
    if (d + 1 < d_end) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    // phase B: walk the run; the texels of pixel i+1 are requested while pixel i is processed
    const float4* g_run = gbuf + buf * PXW * CG + run * L * CG + cg;
    for (int s = 0; s < p.n_src; ++s) {
      if (n_valid == 0) break;
      const size_t vbase = (size_t)p.src_views[s] * view_f4 + cg;
      const float4* timg = img4 + vbase;
      float4* gview = gimg4 + vbase;
      const TapRec* rrec = rec + s * PXW + run * L;
      float4 aW0 = zero4, aW1 = zero4, aE0 = zero4, aE1 = zero4;   // gradient of (row0,x0) (row1,x0) (row0,x0+1) (row1,x0+1)
      float4 tW0 = zero4, tW1 = zero4, tE0 = zero4, tE1 = zero4;   // the texels themselves
      float4 nW0 = zero4, nW1 = zero4, nE0 = zero4, nE1 = zero4;   // texels requested for the next footprint
      int cur = 0;
      bool have = false;
      // transition into pixel 0: everything is new
      TapRec rn = rrec[0];
      int trans = 3;                                               // 0 same footprint, 1 one texel east, 2 one texel west, 3 elsewhere
      if (need_val) { nW0 = ldg4(timg + rn.off4); nE0 = ldg4(timg + rn.off4 + CG); nW1 = ldg4(timg + rn.off4 + row_f4); nE1 = ldg4(timg + rn.off4 + row_f4 + CG); }
#pragma unroll 1
      for (int i = 0; i < n_valid; ++i) {
        const TapRec r = rn;
        if (trans != 0) {
          if (have) {
            float4* gw = gview + cur;
            if (trans != 2) { red_add4(gw, aW0); red_add4(gw + row_f4, aW1); }
            if (trans != 1) { red_add4(gw + CG, aE0); red_add4(gw + row_f4 + CG, aE1); }
          }
          if (trans == 1) { aW0 = aE0; aW1 = aE1; aE0 = zero4; aE1 = zero4; tW0 = tE0; tW1 = tE1; tE0 = nE0; tE1 = nE1; }
          else if (trans == 2) { aE0 = aW0; aE1 = aW1; aW0 = zero4; aW1 = zero4; tE0 = tW0; tE1 = tW1; tW0 = nW0; tW1 = nW1; }
          else { aW0 = zero4; aW1 = zero4; aE0 = zero4; aE1 = zero4; tW0 = nW0; tW1 = nW1; tE0 = nE0; tE1 = nE1; }
          cur = r.off4;
          have = true;
        }
        float4 g = g_run[i * CG];
        if (i + 1 < n_valid) {
          rn = rrec[i + 1];
          const int off = rn.off4;
          trans = off == cur ? 0 : (off == cur + CG ? 1 : (off == cur - CG ? 2 : 3));
          if (need_val && trans != 0) {
            const float4* tw = timg + off;
            if (trans != 1) { nW0 = ldg4(tw); nW1 = ldg4(tw + row_f4); }
            if (trans != 2) { nE0 = ldg4(tw + CG); nE1 = ldg4(tw + row_f4 + CG); }
          }
        }
        g.x *= inv_div; g.y *= inv_div; g.z *= inv_div; g.w *= inv_div;
        const float tx1 = 1.f - r.tx, ty1 = 1.f - r.ty;
        const float wnw = tx1 * ty1, wne = r.tx * ty1, wsw = tx1 * r.ty, wse = r.tx * r.ty;
        float4 dw = g;
        if (need_val) {
          float4 val;   // ATen's accumulation order nw, ne, sw, se: bit-identical to the forward kernel
          val.x = tW0.x * wnw; val.y = tW0.y * wnw; val.z = tW0.z * wnw; val.w = tW0.w * wnw;
          val.x = fmaf(tE0.x, wne, val.x); val.y = fmaf(tE0.y, wne, val.y); val.z = fmaf(tE0.z, wne, val.z); val.w = fmaf(tE0.w, wne, val.w);
          val.x = fmaf(tW1.x, wsw, val.x); val.y = fmaf(tW1.y, wsw, val.y); val.z = fmaf(tW1.z, wsw, val.z); val.w = fmaf(tW1.w, wsw, val.w);
          val.x = fmaf(tE1.x, wse, val.x); val.y = fmaf(tE1.y, wse, val.y); val.z = fmaf(tE1.z, wse, val.z); val.w = fmaf(tE1.w, wse, val.w);
          const float4 rf = ldg4(my_ref + (size_t)i * CG);
          float4 gr = my_gs[i * CG];
          if (p.cost_type == PGRF_COST_ABS_DIFF) {
            // torch.abs backward: grad * sign(x), sign(0) = 0
            dw.x = val.x > rf.x ? g.x : (val.x < rf.x ? -g.x : 0.f);
            dw.y = val.y > rf.y ? g.y : (val.y < rf.y ? -g.y : 0.f);
            dw.z = val.z > rf.z ? g.z : (val.z < rf.z ? -g.z : 0.f);
            dw.w = val.w > rf.w ? g.w : (val.w < rf.w ? -g.w : 0.f);
            gr.x -= dw.x; gr.y -= dw.y; gr.z -= dw.z; gr.w -= dw.w;
          } else {
            dw.x = g.x * rf.x; dw.y = g.y * rf.y; dw.z = g.z * rf.z; dw.w = g.w * rf.w;
            gr.x = fmaf(g.x, val.x, gr.x); gr.y = fmaf(g.y, val.y, gr.y); gr.z = fmaf(g.z, val.z, gr.z); gr.w = fmaf(g.w, val.w, gr.w);
          }
          my_gs[i * CG] = gr;
        }
        aW0.x = fmaf(dw.x, wnw, aW0.x); aW0.y = fmaf(dw.y, wnw, aW0.y); aW0.z = fmaf(dw.z, wnw, aW0.z); aW0.w = fmaf(dw.w, wnw, aW0.w);
        aE0.x = fmaf(dw.x, wne, aE0.x); aE0.y = fmaf(dw.y, wne, aE0.y); aE0.z = fmaf(dw.z, wne, aE0.z); aE0.w = fmaf(dw.w, wne, aE0.w);
        aW1.x = fmaf(dw.x, wsw, aW1.x); aW1.y = fmaf(dw.y, wsw, aW1.y); aW1.z = fmaf(dw.z, wsw, aW1.z); aW1.w = fmaf(dw.w, wsw, aW1.w);
        aE1.x = fmaf(dw.x, wse, aE1.x); aE1.y = fmaf(dw.y, wse, aE1.y); aE1.z = fmaf(dw.z, wse, aE1.z); aE1.w = fmaf(dw.w, wse, aE1.w);
      }
      {
        float4* gw = gview + cur;
        red_add4(gw, aW0); red_add4(gw + row_f4, aW1); red_add4(gw + CG, aE0); red_add4(gw + row_f4 + CG, aE1);
      }
    }
    __syncwarp();
  }
  if (need_val) {
    float4* gref4 = gimg4 + (size_t)p.ref_idx * view_f4 + ((size_t)y * p.W + x_run) * CG + cg;
    for (int i = 0; i < n_valid; ++i) red_add4(gref4 + (size_t)i * CG, my_gs[i * CG]);
  }
}

int g_cv_bwd_variant = 0;   // 0 = one lane per (pixel, channel group) (default), 1 = run-merged (measured equal at C = 32, faster at C = 64)
int g_cv_bwd_minb = 8;      // resident CTAs per SM the default kernel is compiled for (8 = 64 registers, 6 = 77)
int g_cv_bwd_chunks = 0;    // depth chunks override
int g_cv_bwd_run = 0;       // run length override (8 / 16 / 32); 0 = by map width

template <int C, int L, int MINB>
static int launch_bwd_run(const CvBwdParams& q0, cudaStream_t st) {
  CvBwdParams q = q0;
  const CvParams& p = q.f;
  constexpr int PXW = (32 / (C / 4)) * L;
  const size_t warp_bytes = (size_t)(p.cost_type != PGRF_COST_NONE ? 3 : 2) * PXW * C * sizeof(float) + (size_t)p.n_src * PXW * sizeof(TapRec) +
                            3 * PXW * sizeof(float);
  const size_t smem = kCvWarps * warp_bytes;
  PGRF_REQUIRE(smem <= 200 * 1024, "cost_volume_bwd: %zu bytes of shared memory", smem);
  static bool done[64] = {};
  int dev = 0;
  PGRF_CUDA(cudaGetDevice(&dev));
  if (!done[dev & 63]) {
    PGRF_CUDA(cudaFuncSetAttribute(cost_volume_bwd_run_kernel<C, L, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    done[dev & 63] = true;
  }
  const int tiles_x = (p.W + kCvWarps * PXW - 1) / (kCvWarps * PXW);
  // the reference-view gradient is added once per (CTA, pixel): long depth chunks, but >= ~2 waves of CTAs
  const long long ctas_per_chunk = (long long)tiles_x * p.H * p.B;
  int n_chunks = (int)((148LL * 5 * 2 + ctas_per_chunk - 1) / ctas_per_chunk);
  if (n_chunks < 1) n_chunks = 1;
  if (g_cv_bwd_chunks > 0) n_chunks = g_cv_bwd_chunks;
  int d_chunk = (p.D + n_chunks - 1) / n_chunks;
  if (d_chunk < 4) d_chunk = p.D < 4 ? p.D : 4;
  q.f.d_chunk = d_chunk;
  PGRF_REQUIRE((p.D + d_chunk - 1) / d_chunk <= 65535, "cost_volume_bwd: too many depth chunks");
  dim3 grid((unsigned)(tiles_x * p.H), (unsigned)((p.D + d_chunk - 1) / d_chunk), (unsigned)p.B);
  cost_volume_bwd_run_kernel<C, L, MINB><<<grid, kCvThreads, smem, st>>>(q);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

template <int C>
static int launch_bwd_run_c(const CvBwdParams& q, cudaStream_t st) {
  constexpr int NRUN = 32 / (C / 4);
  int L = g_cv_bwd_run;
  if (L != 8 && L != 16) {
    // the longest run whose CTA tile (4 warps x NRUN runs) still fits the row
    L = q.f.W >= kCvWarps * NRUN * 32 ? 32 : (q.f.W >= kCvWarps * NRUN * 16 ? 16 : 8);
    if (L > 8) L = 8;      // measured on B200 (tools/time_cv_bwd.py): short runs, many resident warps
  }
  return L == 16 ? launch_bwd_run<C, 16, 4>(q, st) : launch_bwd_run<C, 8, 4>(q, st);
}

template <int C, int MINB>
static int launch_bwd_lean(const CvBwdParams& q, cudaStream_t st) {
  const CvParams& p = q.f;
  const size_t smem = (size_t)kCvWarps * p.n_src * 32 * sizeof(TapRec) + (size_t)kCvWarps * 32 * C * sizeof(float);
  static bool done[64] = {};
  int dev = 0;
  PGRF_CUDA(cudaGetDevice(&dev));
  if (!done[dev & 63]) {
    PGRF_CUDA(cudaFuncSetAttribute(cost_volume_bwd_lean_kernel<C, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    done[dev & 63] = true;
  }
  dim3 grid((unsigned)(((p.W + kCvThreads - 1) / kCvThreads) * p.H), (unsigned)((p.D + p.d_chunk - 1) / p.d_chunk), (unsigned)p.B);
  cost_volume_bwd_lean_kernel<C, MINB><<<grid, kCvThreads, smem, st>>>(q);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}
template <int C>
static int launch_bwd_lean_c(const CvBwdParams& q, cudaStream_t st) {
  if (g_cv_bwd_minb >= 8) return launch_bwd_lean<C, 8>(q, st);
  return launch_bwd_lean<C, 6>(q, st);
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
  q.no_red = g_cv_bwd_nored;
  // reference-view gradients are added once per (CTA, pixel): prefer long depth chunks, but keep >= ~2 waves of CTAs
  const long long ctas_per_chunk = (long long)((W + kCvThreads - 1) / kCvThreads) * H * B;
  int n_chunks = (int)((148LL * 8 * 2 + ctas_per_chunk - 1) / ctas_per_chunk);
  if (n_chunks < 1) n_chunks = 1;
  int d_chunk = (D + n_chunks - 1) / n_chunks;
  if (d_chunk < 4) d_chunk = D < 4 ? D : 4;
  q.f.d_chunk = d_chunk;
  PGRF_REQUIRE((D + d_chunk - 1) / d_chunk <= 65535, "cost_volume_bwd: too many depth chunks");
  cudaStream_t st = (cudaStream_t)stream;
  if (g_cv_bwd_variant == 1) {
    switch (C) {
      case 4: return launch_bwd_run_c<4>(q, st);
      case 8: return launch_bwd_run_c<8>(q, st);
      case 16: return launch_bwd_run_c<16>(q, st);
      case 32: return launch_bwd_run_c<32>(q, st);
      default: return launch_bwd_run_c<64>(q, st);
    }
  }
  switch (C) {
    case 4: return launch_bwd_lean_c<4>(q, st);
    case 8: return launch_bwd_lean_c<8>(q, st);
    case 16: return launch_bwd_lean_c<16>(q, st);
    case 32: return launch_bwd_lean_c<32>(q, st);
    default: return launch_bwd_lean_c<64>(q, st);
  }
}

// K2/K3/K4 — the per-ray render path of PanoGRF as three persistent sm_100a kernels (fp32 parity path).
//
//   rows kernel    (R1): per (view,sample) row — ray point, projection into every source panorama,
//                        three bilinear gathers, dist-decoder MLPs, logistic-mixture probabilities,
//                        prob_embed, ray_dir_fc, neuray_fc.
//                        reference: render_ops.py:76-106,158-257 ; ops.py:32-52 ; dist_decoder.py:99-140 ;
//                                   renderer.py:120-136,180-188 ; aggregate_net.py:41-70 ; ibrnet.py:326-336
//   samples kernel (R2): per sample across views — weighted mean/var pooling, base_fc, vis_fc, vis_fc2,
//                        rgb_fc + view softmax.            reference: ibrnet.py:336-352,366-372
//   rays kernel    (R3): per ray — geometry_fc, positional code, 4-head ray transformer, out_geometry_fc,
//                        alpha compositing, inverse-CDF fine resampling + sort.
//                        reference: ibrnet.py:352-364 ; render_ops.py:145-153,413-473 ; renderer.py:210-219,302-304,472
//
// Activations live in shared memory feature-major ([feature][row], 128 rows per tile); every Linear
// layer is the register-tiled SIMT GEMM of render_device.cuh with the section's weights resident in
// shared memory for the lifetime of the (persistent) CTA.  Tiles travel between kernels as contiguous
// blocks moved with TMA bulk copies (cp.async.bulk), sized so a ray batch stays L2 resident.
#include <type_traits>

#include "render_device.cuh"
#include "render_layout16.cuh"

// Layout offsets must be constant-evaluated: kLayers is a host constexpr table and may not be indexed at run time
// in device code.
#define WOFF(L) (std::integral_constant<int, pgrf::sec_off(L)>::value)
#define BOFF(L) (std::integral_constant<int, pgrf::bias_off(L)>::value)

namespace pgrf {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int LD = kTileRows;  // row pitch of every [feature][row] activation buffer

struct RenderParams {
  pgrf_render_args a;
  int V, T, M;          // source views, samples per tile, rows per tile (V*T)
  long long total;      // rn * dn samples
  int n_tiles;          // tiles of T samples
  int rays_per_tile3, n_tiles3;  // rays kernel tiling
};

// cooperative copy of one weight section into shared memory
__device__ __forceinline__ void load_section(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x * 4; i < n; i += kThreads * 4) {
    if (i + 3 < n) *reinterpret_cast<float4*>(dst + i) = __ldg(reinterpret_cast<const float4*>(src + i));
    else for (int j = i; j < n; ++j) dst[j] = __ldg(src + j);
  }
}

// =================================================================================================
// R1: rows kernel
// =================================================================================================
constexpr int R1_W = section_floats(0);
constexpr int R1_IN = R1_W;                    // [34][LD]
constexpr int R1_H1 = R1_IN + 34 * LD;         // [32][LD]
constexpr int R1_H2 = R1_H1 + 32 * LD;         // [32][LD]
constexpr int R1_OUT = R1_H2 + 32 * LD;        // [kF1][LD]
constexpr int R1_SC = R1_OUT + kF1 * LD;       // per-row scalars, 10 x [LD]
constexpr int R1_FLOATS = R1_SC + 10 * LD;
enum { SC_PX = 0, SC_PY, SC_PDEPTH, SC_MEAN0, SC_MEAN1, SC_VAR0, SC_VAR1, SC_AW, SC_VIS, SC_SPARE };

// One decoder of MixtureLogisticsDistDecoder (dist_decoder.py:64-97): D = 0 mean, 1 var, 2 aw, 3 vis.
template <int D>
__device__ __forceinline__ void decoder_stage(const float* W, const float* IN, float* H1, float* H2, float* SC, int Mp,
                                              float bias_val, int tid, int warp, int lane) {
  constexpr int l0 = L_MEAN0 + 3 * D;
  gemm_smem<4, ACT_ELU, false, false>(IN, LD, 32, W + WOFF(l0), 32, W + BOFF(l0), H1, LD, Mp, 32,
                                      nullptr, nullptr, 1, 0, warp, lane, kWarps);
  __syncthreads();
  gemm_smem<4, ACT_ELU, false, false>(H1, LD, 32, W + WOFF(l0 + 1), 32, W + BOFF(l0 + 1), H2, LD, Mp, 32,
                                      nullptr, nullptr, 1, 0, warp, lane, kWarps);
  __syncthreads();
  if (tid < LD) {
    float o[4];
    row_layer<32, 4, 4>(H2, LD, tid, W + WOFF(l0 + 2), W + BOFF(l0 + 2), o);
    if (D == 0) { SC[SC_MEAN0 * LD + tid] = softplusf(o[0]); SC[SC_MEAN1 * LD + tid] = softplusf(o[1]); }
    else if (D == 1) { SC[SC_VAR0 * LD + tid] = softplusf(o[0]) + bias_val; SC[SC_VAR1 * LD + tid] = softplusf(o[1]) + bias_val; }
    else if (D == 2) SC[SC_AW * LD + tid] = sigmoidf(o[0]);
    else SC[SC_VIS * LD + tid] = sigmoidf(o[0]);
  }
  // the next stage's first GEMM writes H1 only (H2 readers are separated from the next H2 writer by its barrier)
}

__global__ void __launch_bounds__(kThreads, 1) render_rows_kernel(const RenderParams p) {
  extern __shared__ __align__(128) float sm[];
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* W = sm;
  float* IN = sm + R1_IN;
  float* H1 = sm + R1_H1;
  float* H2 = sm + R1_H2;
  float* OUT = sm + R1_OUT;
  float* SC = sm + R1_SC;
  { constexpr int sec0 = section_begin(0); load_section(W, a.weights + sec0, R1_W); }
  const int T = p.T, V = p.V, M = p.M;
  const int Mp = (M + 31) & ~31;
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    // ---------------- geometry: one thread per row ----------------
    if (tid < LD) {
      const int m = tid;
      if (m < M) {
        const int v = m / T, t = m % T;
        long long g = (long long)tile * T + t;
        if (g >= p.total) g = p.total - 1;   // padded rows replicate the last sample (never read back)
        RowGeom r;
        if (a.prj_in) {   // the reference's prj_dict rows: pixel, depth and direction are given
          const float* d = a.prj_in + ((size_t)v * p.total + g) * 6;
          const float* q = a.que_dir_in + (size_t)g * 3;
          r.px = __ldg(d); r.py = __ldg(d + 1); r.pdepth = __ldg(d + 2);
          r.dir[0] = __ldg(d + 3); r.dir[1] = __ldg(d + 4); r.dir[2] = __ldg(d + 5);
          const float q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
          r.dirdiff[0] = r.dir[0] - q0; r.dirdiff[1] = r.dir[1] - q1; r.dirdiff[2] = r.dir[2] - q2;
          r.dirdiff[3] = r.dir[0] * q0 + r.dir[1] * q1 + r.dir[2] * q2;
        } else {
          r = row_geometry<false>(a, v, g);
        }
        SC[SC_PX * LD + m] = r.px; SC[SC_PY * LD + m] = r.py; SC[SC_PDEPTH * LD + m] = r.pdepth;
#pragma unroll
        for (int i = 0; i < 4; ++i) OUT[(F1_DIRDIFF + i) * LD + m] = r.dirdiff[i];
        if (a.prj_dbg) {  // optional prj_dict dump: pts(2) depth(1) dir(3)
          const long long gg = (long long)tile * T + t;
          if (gg < p.total) {
            float* d = a.prj_dbg + ((size_t)v * p.total + gg) * 6;
            d[0] = r.px; d[1] = r.py; d[2] = r.pdepth; d[3] = r.dir[0]; d[4] = r.dir[1]; d[5] = r.dir[2];
          }
        }
      } else {
        SC[SC_PX * LD + m] = 0.f; SC[SC_PY * LD + m] = 0.f; SC[SC_PDEPTH * LD + m] = 1.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) OUT[(F1_DIRDIFF + i) * LD + m] = 0.f;
      }
    }
    __syncthreads();

    // ---------------- gathers: lane <-> (row, float4 channel group) ----------------
    if (a.feat_in) {   // gathered features are given (prj_dict['ray_feats'], ['rgb'], ['img_feats'])
      for (int it = tid; it < LD * 67; it += kThreads) {
        const int m = it / 67, c = it % 67;
        float val = 0.f;
        if (m < M) {
          const int v = m / T, t = m % T;
          long long g = (long long)tile * T + t;
          if (g >= p.total) g = p.total - 1;
          val = __ldg(a.feat_in + ((size_t)v * p.total + g) * 67 + c);
        }
        if (c < 32) IN[c * LD + m] = val;
        else if (c < 35) { OUT[(c - 32) * LD + m] = val; OUT[(F1_RGBRAW + c - 32) * LD + m] = val; }
        else OUT[(3 + c - 35) * LD + m] = val;
      }
      if (tid < LD) OUT[(F1_RGBRAW + 3) * LD + tid] = 0.f;
    } else {
    for (int it = tid; it < LD * 8; it += kThreads) {
      const int m = it >> 3, cg = it & 7;
      float4 rf = make_float4(0.f, 0.f, 0.f, 0.f), imf = rf;
      if (m < M) {
        const int v = m / T;
        const float px = SC[SC_PX * LD + m], py = SC[SC_PY * LD + m];
        const Footprint f1 = border_footprint(px, py, a.img_h, a.img_w, a.rf_h, a.rf_w);
        rf = tap4(reinterpret_cast<const float4*>(a.ray_feats_cl) + ((size_t)v * a.rf_h * a.rf_w + f1.off) * 8 + cg,
                  f1, 8, a.rf_w * 8);
        const Footprint f2 = border_footprint(px, py, a.img_h, a.img_w, a.if_h, a.if_w);
        imf = tap4(reinterpret_cast<const float4*>(a.img_feats_cl) + ((size_t)v * a.if_h * a.if_w + f2.off) * 8 + cg,
                   f2, 8, a.if_w * 8);
      }
      IN[(4 * cg + 0) * LD + m] = rf.x; IN[(4 * cg + 1) * LD + m] = rf.y;
      IN[(4 * cg + 2) * LD + m] = rf.z; IN[(4 * cg + 3) * LD + m] = rf.w;
      OUT[(3 + 4 * cg + 0) * LD + m] = imf.x; OUT[(3 + 4 * cg + 1) * LD + m] = imf.y;
      OUT[(3 + 4 * cg + 2) * LD + m] = imf.z; OUT[(3 + 4 * cg + 3) * LD + m] = imf.w;
    }
    if (tid < LD) {
      const int m = tid;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const int v = m / T;
        const Footprint f = border_footprint(SC[SC_PX * LD + m], SC[SC_PY * LD + m], a.img_h, a.img_w, a.img_h, a.img_w);
        c = tap4(reinterpret_cast<const float4*>(a.imgs_cl) + (size_t)v * a.img_h * a.img_w + f.off, f, 1, a.img_w);
      }
      OUT[0 * LD + m] = c.x; OUT[1 * LD + m] = c.y; OUT[2 * LD + m] = c.z;
      OUT[(F1_RGBRAW + 0) * LD + m] = c.x; OUT[(F1_RGBRAW + 1) * LD + m] = c.y; OUT[(F1_RGBRAW + 2) * LD + m] = c.z;
      OUT[(F1_RGBRAW + 3) * LD + m] = 0.f;
    }
    }
    __syncthreads();
    if (a.feat_dbg) {  // optional prj_dict dump of the gathered features: ray_feats(32) rgb(3) img_feats(32)
      for (int it = tid; it < M * 67; it += kThreads) {
        const int m = it / 67, c = it % 67;
        const int v = m / T, t = m % T;
        const long long gg = (long long)tile * T + t;
        if (gg < p.total) {
          const float val = c < 32 ? IN[c * LD + m] : (c < 35 ? OUT[(c - 32) * LD + m] : OUT[(3 + c - 35) * LD + m]);
          a.feat_dbg[((size_t)v * p.total + gg) * 67 + c] = val;
        }
      }
    }

    // ---------------- dist decoder: 3 (4) MLPs 32 -> 32 -> 32 -> {2,2,1,1} ----------------
    if (!a.prob_in) {
      decoder_stage<0>(W, IN, H1, H2, SC, Mp, a.bias_val, tid, warp, lane);
      decoder_stage<1>(W, IN, H1, H2, SC, Mp, a.bias_val, tid, warp, lane);
      decoder_stage<2>(W, IN, H1, H2, SC, Mp, a.bias_val, tid, warp, lane);
      if (a.use_vis) decoder_stage<3>(W, IN, H1, H2, SC, Mp, a.bias_val, tid, warp, lane);
    }
    __syncthreads();

    // ---------------- logistic-mixture probabilities (dist_decoder.compute_prob, is_ref=True) ----------------
    if (tid < LD) {
      const int m = tid;
      float hit = 0.5f, vis = 0.5f, alpha = 0.f;
      if (m < M) {
        const int v = m / T, t = m % T;
        long long g = (long long)tile * T + t;
        if (g >= p.total) g = p.total - 1;
        const int ray = (int)(g / a.dn), s = (int)(g % a.dn);
        if (a.prob_in) {   // prj_dict['alpha' | 'vis' | 'hit_prob'] are given: the dist decoder did not run
          const float* d = a.prob_in + ((size_t)v * p.total + g) * 3;
          alpha = __ldg(d); vis = __ldg(d + 1); hit = __ldg(d + 2);
        } else {
          float d_s, d_prev;
          if (a.interval_in) {   // que_dists given (get_near_far_points, dist_decoder.py:6-51, is_ref=True)
            d_s = __ldg(a.interval_in + g);
            d_prev = s > 0 ? __ldg(a.interval_in + g - 1) : d_s;
          } else {
            const float* dp = a.depth + (size_t)ray * a.depth_ray_stride;
            // que_dists = depth2inv_dists(que_depth, que depth_range): interval s = inv[s+1]-inv[s], last 1e6
            const float i_s = inv_norm(__ldg(dp + s), a.que_near, a.que_far);
            d_s = (s + 1 < a.dn) ? inv_norm(__ldg(dp + s + 1), a.que_near, a.que_far) - i_s : 1e6f;
            d_prev = d_s;  // interval_ext[0] = interval_half[0]
            if (s > 0) d_prev = i_s - inv_norm(__ldg(dp + s - 1), a.que_near, a.que_far);
          }
          const float rnear = __ldg(a.ref_depth_range + 2 * v), rfar = __ldg(a.ref_depth_range + 2 * v + 1);
          const float dv = inv_norm(fmaxf(SC[SC_PDEPTH * LD + m], 1e-5f), rnear, rfar);
          const float nearp = dv - d_prev / 2.f, farp = dv + d_s / 2.f;
          const float aw = SC[SC_AW * LD + m];
          const float mix[2] = {aw, 1.f - aw};
          const float mean[2] = {SC[SC_MEAN0 * LD + m], SC[SC_MEAN1 * LD + m]};
          const float var[2] = {SC[SC_VAR0 * LD + m], SC[SC_VAR1 * LD + m]};
          float visibility = 0.f, hp = 0.f;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float cdf0 = 0.5f + 0.5f * tanhf((nearp - mean[j]) * var[j]);
            float cdf1 = 0.5f + 0.5f * tanhf((farp - mean[j]) * var[j]);
            if (a.use_vis) { cdf0 *= SC[SC_VIS * LD + m]; cdf1 *= SC[SC_VIS * LD + m]; }
            visibility += (1.f - cdf0) * mix[j];
            hp += (cdf1 - cdf0) * mix[j];
          }
          hit = hp; vis = visibility;
          alpha = logf(hp / (visibility - hp + 1e-5f) + 1e-5f);
          if (a.dec_dbg) {
            const long long gg = (long long)tile * T + t;
            if (gg < p.total) {
              float* d = a.dec_dbg + ((size_t)v * p.total + gg) * 6;
              d[0] = mean[0]; d[1] = mean[1]; d[2] = var[0]; d[3] = var[1];
              d[4] = a.use_vis ? SC[SC_VIS * LD + m] : 1.f; d[5] = aw;
            }
          }
        }
        if (a.prob_dbg) {
          const long long gg = (long long)tile * T + t;
          if (gg < p.total) {
            float* d = a.prob_dbg + ((size_t)v * p.total + gg) * 3;
            d[0] = alpha; d[1] = vis; d[2] = hit;
          }
        }
      }
      IN[32 * LD + m] = (hit - 0.5f) * 2.f;
      IN[33 * LD + m] = (vis - 0.5f) * 2.f;
    }
    __syncthreads();
    if (a.wo_appearance) {   // aggregate_net.py:79-81: [rgb, img_feats] = 0 (raw colours included: ibrnet takes rgb_in from them)
      for (int it = tid; it < 35 * LD; it += kThreads) OUT[it] = 0.f;
      for (int it = tid; it < 3 * LD; it += kThreads) OUT[F1_RGBRAW * LD + it] = 0.f;
      __syncthreads();
    }

    // ---------------- prob_embed 34 -> 32 (ReLU) -> 32, straight into the output block ----------------
    if (a.wo_geometry) {     // aggregate_net.py:60-62: prob_embedding = 0
      for (int it = tid; it < 32 * LD; it += kThreads) OUT[F1_NEURAY * LD + it] = 0.f;
    } else {
    gemm_smem<4, ACT_RELU, false, false>(IN, LD, 34, W + WOFF(L_PE0), 32, W + BOFF(L_PE0), H1, LD, Mp, 32,
                                         nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    gemm_smem<4, ACT_NONE, false, false>(H1, LD, 32, W + WOFF(L_PE1), 32, W + BOFF(L_PE1), OUT + F1_NEURAY * LD, LD,
                                         Mp, 32, nullptr, nullptr, 1, 0, warp, lane, kWarps);
    }
    __syncthreads();

    // ---------------- neuray_fc (threads 0..127) and ray_dir_fc (threads 128..255) ----------------
    if (tid < LD) {
      float h[8], o[4];
      row_layer<32, 8, 8>(OUT + F1_NEURAY * LD, LD, tid, W + WOFF(L_NF0), W + BOFF(L_NF0), h);
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = elu1(h[i]);
      reg_layer<8, 1, 4>(h, W + WOFF(L_NF1), W + BOFF(L_NF1), o);
      OUT[F1_W0 * LD + tid] = sigmoidf(o[0]);
    } else {
      const int m = tid - LD;
      float in[4], h[16], o[36];
#pragma unroll
      for (int i = 0; i < 4; ++i) in[i] = OUT[(F1_DIRDIFF + i) * LD + m];
      reg_layer<4, 16, 16>(in, W + WOFF(L_RD0), W + BOFF(L_RD0), h);
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = elu1(h[i]);
      reg_layer<16, 35, 36>(h, W + WOFF(L_RD1), W + BOFF(L_RD1), o);
#pragma unroll
      for (int i = 0; i < 35; ++i) OUT[i * LD + m] += elu1(o[i]);
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.f1 + (size_t)tile * kF1 * LD, OUT, kF1 * LD * sizeof(float));
      bulk_wait_read();
    }
    __syncthreads();
  }
}

// =================================================================================================
// R2: samples kernel
// =================================================================================================
constexpr int R2_W = section_floats(1);
constexpr int R2_IN = (R2_W + 31) & ~31;       // [kF1][LD]   (TMA destination, 128-byte aligned)
constexpr int R2_RA = R2_IN + kF1 * LD;        // [140][T] globalfeat | [36][LD] vis_fc out | [kF2][T] pooled
constexpr int R2_RA_FLOATS = 140 * 64 + 64;
constexpr int R2_G = R2_RA + R2_RA_FLOATS;     // [64][64]
constexpr int R2_HID = R2_G + 64 * 64;         // [64][LD]
constexpr int R2_X = R2_HID + 64 * LD;         // [32][LD]
constexpr int R2_SV = R2_X + 32 * LD;          // small per-row vectors, 4 x [LD]
constexpr int R2_BAR = R2_SV + 4 * LD;         // mbarrier (8 bytes)
constexpr int R2_FLOATS = R2_BAR + 4;
enum { SV_SCALE = 0, SV_VIS, SV_WT, SV_LOGIT };

__global__ void __launch_bounds__(kThreads, 1) render_samples_kernel(const RenderParams p) {
  extern __shared__ __align__(128) float sm[];
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* W = sm;
  float* IN = sm + R2_IN;
  float* RA = sm + R2_RA;
  float* G = sm + R2_G;
  float* HID = sm + R2_HID;
  float* X = sm + R2_X;
  float* SV = sm + R2_SV;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + R2_BAR);
  const int T = p.T, V = p.V, M = p.M;
  const int Mp = (M + 31) & ~31;
  const float wgt = 1.f / ((float)V + 1e-8f);   // mask / (sum(mask) + 1e-8), mask == 1 (ibrnet.py:336)
  if (tid == 0) mbar_init(bar, 1);
  { constexpr int sec1 = section_begin(1); load_section(W, a.weights + sec1, R2_W); }
  for (int i = tid; i < LD; i += kThreads) SV[SV_SCALE * LD + i] = wgt;
  __syncthreads();
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    if (tid == 0) {
      mbar_expect_tx(bar, kF1 * LD * sizeof(float));
      bulk_g2s(IN, a.f1 + (size_t)tile * kF1 * LD, kF1 * LD * sizeof(float), bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;

    // ---- weighted mean / variance pooling over views (fused_mean_variance x2, ibrnet.py:338-341) ----
    for (int it = tid; it < 35 * T; it += kThreads) {
      const int f = it / T, t = it % T;
      float m0 = 0.f, m1 = 0.f;
      for (int v = 0; v < V; ++v) {
        const float x = IN[f * LD + v * T + t];
        m0 += x * (IN[F1_W0 * LD + v * T + t] * wgt);
        m1 += x * wgt;
      }
      float v0 = 0.f, v1 = 0.f;
      for (int v = 0; v < V; ++v) {
        const float x = IN[f * LD + v * T + t];
        const float w0 = IN[F1_W0 * LD + v * T + t] * wgt;
        v0 += w0 * ((x - m0) * (x - m0));
        v1 += wgt * ((x - m1) * (x - m1));
      }
      RA[f * T + t] = m0; RA[(35 + f) * T + t] = v0; RA[(70 + f) * T + t] = m1; RA[(105 + f) * T + t] = v1;
    }
    __syncthreads();
    // ---- base_fc.0 split: per-sample part (140 pooled features) + per-row part (67 features) ----
    gemm_smem<4, ACT_NONE, false, false>(RA, T, 140, W + WOFF(L_BASE0G), 64, nullptr, G, 64, T, 64,
                                         nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    gemm_smem<8, ACT_ELU, false, true>(IN, LD, 67, W + WOFF(L_BASE0R), 64, W + BOFF(L_BASE0R), HID, LD, Mp, 64,
                                       nullptr, G, T, 64, warp, lane, kWarps);
    __syncthreads();
    gemm_smem<4, ACT_ELU, false, false>(HID, LD, 64, W + WOFF(L_BASE1), 32, W + BOFF(L_BASE1), X, LD, Mp, 32,
                                        nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    // ---- vis_fc(x * weight): 32 -> 32 -> 33 ----
    gemm_smem<4, ACT_ELU, true, false>(X, LD, 32, W + WOFF(L_VFC0), 32, W + BOFF(L_VFC0), HID, LD, Mp, 32,
                                       SV + SV_SCALE * LD, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    float* XV = RA;  // [36][LD]
    gemm_smem<4, ACT_ELU, false, false>(HID, LD, 32, W + WOFF(L_VFC1), 36, W + BOFF(L_VFC1), XV, LD, Mp, 36,
                                        nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    for (int it = tid; it < 32 * LD; it += kThreads) X[it] += XV[it];         // x = x + x_res
    if (tid < LD) SV[SV_VIS * LD + tid] = sigmoidf(XV[32 * LD + tid]);        // vis = sigmoid(vis) * mask
    __syncthreads();
    // ---- vis_fc2(x * vis): 32 -> 32 -> 1 ----
    gemm_smem<4, ACT_ELU, true, false>(X, LD, 32, W + WOFF(L_VFC2_0), 32, W + BOFF(L_VFC2_0), HID, LD, Mp, 32,
                                       SV + SV_VIS * LD, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    if (tid < LD) {
      float o[4];
      row_layer<32, 4, 4>(HID, LD, tid, W + WOFF(L_VFC2_1), W + BOFF(L_VFC2_1), o);
      SV[SV_VIS * LD + tid] = sigmoidf(o[0]);                                  // vis (second estimate) * mask
    }
    __syncthreads();
    // ---- weight = vis / (sum_v vis + 1e-8); pooled mean/var of x; mean of weights ----
    float* MV = RA;  // [kF2][T]  (XV is dead)
    if (tid < T) {
      float s = 0.f;
      for (int v = 0; v < V; ++v) s += SV[SV_VIS * LD + v * T + tid];
      float ws = 0.f;
      for (int v = 0; v < V; ++v) {
        const float w = SV[SV_VIS * LD + v * T + tid] / (s + 1e-8f);
        SV[SV_WT * LD + v * T + tid] = w;
        ws += w;
      }
      SV[SV_LOGIT * LD + tid] = ws / (float)V;   // stash weight.mean(dim=2) until XV readers are done
    }
    __syncthreads();
    for (int it = tid; it < 32 * T; it += kThreads) {
      const int c = it / T, t = it % T;
      float mean = 0.f;
      for (int v = 0; v < V; ++v) mean += X[c * LD + v * T + t] * SV[SV_WT * LD + v * T + t];
      float var = 0.f;
      for (int v = 0; v < V; ++v) {
        const float d = X[c * LD + v * T + t] - mean;
        var += SV[SV_WT * LD + v * T + t] * (d * d);
      }
      MV[c * T + t] = mean; MV[(32 + c) * T + t] = var;
    }
    if (tid < T) MV[64 * T + tid] = SV[SV_LOGIT * LD + tid];
    __syncthreads();
    // ---- rgb_fc([x, vis, ray_diff]) 37 -> 16 -> 8 -> 1, softmax over views, blend raw colours ----
    if (tid < LD) {
      const int m = tid;
      const float* w0 = W + WOFF(L_RGB0);
      float h[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) h[n] = W[BOFF(L_RGB0) + n];
      for (int k = 0; k < 37; ++k) {
        const float x = k < 32 ? X[k * LD + m] : (k == 32 ? SV[SV_VIS * LD + m] : IN[(F1_DIRDIFF + k - 33) * LD + m]);
#pragma unroll
        for (int n4 = 0; n4 < 4; ++n4) {
          const float4 w = *reinterpret_cast<const float4*>(w0 + k * 16 + 4 * n4);
          h[4 * n4] = fmaf(x, w.x, h[4 * n4]); h[4 * n4 + 1] = fmaf(x, w.y, h[4 * n4 + 1]);
          h[4 * n4 + 2] = fmaf(x, w.z, h[4 * n4 + 2]); h[4 * n4 + 3] = fmaf(x, w.w, h[4 * n4 + 3]);
        }
      }
#pragma unroll
      for (int n = 0; n < 16; ++n) h[n] = elu1(h[n]);
      float h2[8], o[4];
      reg_layer<16, 8, 8>(h, W + WOFF(L_RGB1), W + BOFF(L_RGB1), h2);
#pragma unroll
      for (int n = 0; n < 8; ++n) h2[n] = elu1(h2[n]);
      reg_layer<8, 1, 4>(h2, W + WOFF(L_RGB2), W + BOFF(L_RGB2), o);
      SV[SV_LOGIT * LD + m] = o[0];
    }
    __syncthreads();
    if (tid < T) {
      float mx = -INFINITY;
      for (int v = 0; v < V; ++v) mx = fmaxf(mx, SV[SV_LOGIT * LD + v * T + tid]);
      float den = 0.f, r = 0.f, g = 0.f, b = 0.f;
      for (int v = 0; v < V; ++v) {
        const float e = expf(SV[SV_LOGIT * LD + v * T + tid] - mx);
        den += e;
        r += IN[(F1_RGBRAW + 0) * LD + v * T + tid] * e;
        g += IN[(F1_RGBRAW + 1) * LD + v * T + tid] * e;
        b += IN[(F1_RGBRAW + 2) * LD + v * T + tid] * e;
      }
      MV[(F2_RGB + 0) * T + tid] = r / den; MV[(F2_RGB + 1) * T + tid] = g / den; MV[(F2_RGB + 2) * T + tid] = b / den;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g(a.f2 + (size_t)tile * kF2 * T, MV, kF2 * T * sizeof(float));
      bulk_wait_read();
    }
    __syncthreads();
  }
}

// =================================================================================================
// R3: rays kernel
// =================================================================================================
constexpr int R3_WF = section_floats(2) + 32;        // + layer-norm weight/bias
constexpr int R3_PE = (R3_WF + 3) & ~3;              // positional table [kMaxSamplesPerRay][16]
constexpr int R3_A = R3_PE + kMaxSamplesPerRay * 16; // [kF2][LD]
constexpr int R3_H = R3_A + kF2 * LD;                // [64][LD]
constexpr int R3_G16 = R3_H + 64 * LD;               // [16][LD]
constexpr int R3_QKV = R3_G16 + 16 * LD;             // [48][LD]
constexpr int R3_AO = R3_QKV + 48 * LD;              // [16][LD]
constexpr int R3_RV = R3_AO + 16 * LD;               // per-row vectors, 6 x [2*LD]
constexpr int R3_FLOATS = R3_RV + 6 * 2 * LD;
enum { RV_SIGMA = 0, RV_ALPHA, RV_HIT, RV_CDF, RV_CENTER, RV_FINE };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) render_rays_kernel(const RenderParams p) {
  extern __shared__ __align__(128) float sm[];
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* W = sm;
  float* PE = sm + R3_PE;
  float* A = sm + R3_A;
  float* H = sm + R3_H;
  float* G16 = sm + R3_G16;
  float* QKV = sm + R3_QKV;
  float* AO = sm + R3_AO;
  float* RV = sm + R3_RV;
  const int dn = a.dn, T = p.T, V = p.V;
  const int rpt = p.rays_per_tile3;
  const int Mv = rpt * dn;                 // valid rows per tile
  const int Mp = (Mv + 31) & ~31;
  constexpr int sec2 = section_begin(2);
  load_section(W, a.weights + sec2, R3_WF);
  for (int i = tid; i < dn * 16; i += kThreads) PE[i] = __ldg(a.weights + kPosencOffset + i);
  __syncthreads();
  constexpr int ln_rel = section_floats(2);
  const float* LNW = W + ln_rel;
  const float* LNB = LNW + 16;

  for (int tile = blockIdx.x; tile < p.n_tiles3; tile += gridDim.x) {
    const long long g0 = (long long)tile * Mv;
    // ---- load pooled features of the tile's samples (generic sample -> F2 block mapping) ----
    for (int it = tid; it < kF2 * LD; it += kThreads) {
      const int f = it / LD, m = it % LD;
      const long long g = g0 + m;
      float val = 0.f;
      if (m < Mv && g < p.total) val = __ldg(a.f2 + (size_t)(g / T) * kF2 * T + (size_t)f * T + (g % T));
      A[f * LD + m] = val;
    }
    __syncthreads();
    // ---- geometry_fc 65 -> 64 -> 16 (+ positional code) ----
    gemm_smem<8, ACT_ELU, false, false>(A, LD, 65, W + WOFF(L_GEO0), 64, W + BOFF(L_GEO0), H, LD, Mp, 64,
                                        nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    gemm_smem<4, ACT_ELU, false, false>(H, LD, 64, W + WOFF(L_GEO1), 16, W + BOFF(L_GEO1), G16, LD, Mp, 16,
                                        nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    for (int it = tid; it < 16 * LD; it += kThreads) {
      const int c = it / LD, m = it % LD;
      if (m < Mv) G16[it] += PE[(m % dn) * 16 + c];
    }
    __syncthreads();
    // ---- q,k,v projections (no bias) ----
    gemm_smem<4, ACT_NONE, false, false>(G16, LD, 16, W + WOFF(L_QKV), 48, nullptr, QKV, LD, Mp, 48,
                                         nullptr, nullptr, 1, 0, warp, lane, kWarps);
    __syncthreads();
    // ---- 4-head attention over the samples of each ray (ibrnet.py:15-27,72-102) ----
    for (int it = tid; it < 4 * LD; it += kThreads) {
      const int h = it / LD, m = it % LD;
      if (m >= Mv) continue;
      const int r0 = (m / dn) * dn;
      const float q0 = QKV[(4 * h + 0) * LD + m] / 2.f, q1 = QKV[(4 * h + 1) * LD + m] / 2.f;
      const float q2 = QKV[(4 * h + 2) * LD + m] / 2.f, q3 = QKV[(4 * h + 3) * LD + m] / 2.f;
      const float* Kp = QKV + (16 + 4 * h) * LD + r0;
      const float* Vp = QKV + (32 + 4 * h) * LD + r0;
      const bool masked = !(V > 1);   // mask = (num_valid_obs > 1), broadcast over keys
      float mx = -INFINITY;
      for (int j = 0; j < dn; ++j) {
        float sc = q0 * Kp[j] + q1 * Kp[LD + j] + q2 * Kp[2 * LD + j] + q3 * Kp[3 * LD + j];
        if (masked) sc = -1e9f;
        mx = fmaxf(mx, sc);
      }
      float den = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      for (int j = 0; j < dn; ++j) {
        float sc = q0 * Kp[j] + q1 * Kp[LD + j] + q2 * Kp[2 * LD + j] + q3 * Kp[3 * LD + j];
        if (masked) sc = -1e9f;
        const float e = expf(sc - mx);
        den += e;
        o0 = fmaf(e, Vp[j], o0); o1 = fmaf(e, Vp[LD + j], o1); o2 = fmaf(e, Vp[2 * LD + j], o2); o3 = fmaf(e, Vp[3 * LD + j], o3);
      }
      AO[(4 * h + 0) * LD + m] = o0 / den; AO[(4 * h + 1) * LD + m] = o1 / den;
      AO[(4 * h + 2) * LD + m] = o2 / den; AO[(4 * h + 3) * LD + m] = o3 / den;
    }
    __syncthreads();
    // ---- fc + residual + LayerNorm(eps 1e-6) + out_geometry_fc 16 -> 16 -> 1 (ReLU) ----
    if (tid < LD && tid < Mv) {
      const int m = tid;
      float o[16], x[16];
      row_layer<16, 16, 16>(AO, LD, m, W + WOFF(L_AFC), nullptr, o);
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] = o[i] + G16[i * LD + m]; mean += x[i]; }
      mean /= 16.f;
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) var += (x[i] - mean) * (x[i] - mean);
      var /= 16.f;
      const float rstd = 1.f / sqrtf(var + 1e-6f);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (x[i] - mean) * rstd * LNW[i] + LNB[i];
      float h1[16], s4[4];
      reg_layer<16, 16, 16>(x, W + WOFF(L_OG0), W + BOFF(L_OG0), h1);
#pragma unroll
      for (int i = 0; i < 16; ++i) h1[i] = elu1(h1[i]);
      reg_layer<16, 1, 4>(h1, W + WOFF(L_OG1), W + BOFF(L_OG1), s4);
      RV[RV_SIGMA * 2 * LD + m] = fmaxf(s4[0], 0.f);
    }
    __syncthreads();
    // ---- per ray: alpha compositing (+ fine resampling), one warp per ray ----
    for (int r = warp; r < rpt; r += kWarps) {
      const long long ray = (long long)tile * rpt + r;
      if (ray >= a.rn) continue;
      const int m0 = r * dn;
      float* alpha = RV + RV_ALPHA * 2 * LD + m0;
      float* hit = RV + RV_HIT * 2 * LD + m0;
      const float* dp = a.depth ? a.depth + (size_t)ray * a.depth_ray_stride : nullptr;   // NULL only in the module-level agg mode
      for (int s = lane; s < dn; s += 32) alpha[s] = 1.f - expf(-RV[RV_SIGMA * 2 * LD + m0 + s]);
      __syncwarp();
      if (lane == 0) {  // sequential fp32 cumprod, the order torch uses on the CPU (render_ops.py:150-152)
        float trans = 1.f;
        for (int s = 0; s < dn; ++s) {
          hit[s] = alpha[s] * trans;
          trans = trans * (1.f - alpha[s] + 1e-10f);
        }
      }
      __syncwarp();
      float cr = 0.f, cg = 0.f, cb = 0.f, cd = 0.f;
      for (int s = lane; s < dn; s += 32) {
        const float hs = hit[s];
        const float r_ = A[(F2_RGB + 0) * LD + m0 + s], g_ = A[(F2_RGB + 1) * LD + m0 + s], b_ = A[(F2_RGB + 2) * LD + m0 + s];
        cr = fmaf(hs, r_, cr); cg = fmaf(hs, g_, cg); cb = fmaf(hs, b_, cb);
        if (dp) cd = fmaf(hs, __ldg(dp + s), cd);
        if (a.hit_prob) a.hit_prob[(size_t)ray * dn + s] = hs;
        if (a.density) a.density[(size_t)ray * dn + s] = RV[RV_SIGMA * 2 * LD + m0 + s];
        if (a.colors) {
          float* c = a.colors + ((size_t)ray * dn + s) * 3;
          c[0] = r_; c[1] = g_; c[2] = b_;
        }
      }
      cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); cd = warp_sum(cd);
      if (lane == 0) {
        a.pixel_colors[(size_t)ray * 3 + 0] = cr; a.pixel_colors[(size_t)ray * 3 + 1] = cg; a.pixel_colors[(size_t)ray * 3 + 2] = cb;
        if (a.render_depth) a.render_depth[ray] = cd;
      }
      if (a.fine_depth) {
        // ---- sample_fine_depth (render_ops.py:413-473), deterministic u-table ----
        float* cdf = RV + RV_CDF * 2 * LD + 2 * m0;       // dn+1 entries (rows are 2*LD wide)
        float* center = RV + RV_CENTER * 2 * LD + 2 * m0; // dn+1 entries
        float* fine = RV + RV_FINE * 2 * LD + 2 * m0;     // up to fine_dn + dn entries
        const bool inv = a.use_disp != 0;
        const float nn = -1.f / a.que_near, ff = -1.f / a.que_far;
        for (int s = lane; s <= dn; s += 32) {
          float d1 = __ldg(dp + min(s, dn - 1));
          float d0 = __ldg(dp + max(s - 1, 0));
          if (inv) { d1 = (-1.f / d1 - nn) / (ff - nn); d0 = (-1.f / d0 - nn) / (ff - nn); }
          center[s] = (s == 0 || s == dn) ? d1 : (d1 + d0) / 2.f;
        }
        if (lane == 0) {  // sequential sum and cumsum (stated accumulation order)
          float tot = 0.f;
          for (int s = 0; s < dn; ++s) tot += hit[s] + 1e-5f;
          float c = 0.f;
          cdf[0] = 0.f;
          for (int s = 0; s < dn; ++s) { c += (hit[s] + 1e-5f) / tot; cdf[s + 1] = c; }
        }
        __syncwarp();
        const int fdn = a.fine_dn;
        for (int k = lane; k < fdn; k += 32) {
          const float u = __ldg(a.fine_u + k);
          int lo = 0, hi = dn + 1;                 // searchsorted(cdf, u, right=True): first index with cdf > u
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
        // rank sort (value-only result == torch.sort)
        for (int k = lane; k < total_out; k += 32) {
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
    __syncthreads();
  }
}

// SM count of the CURRENT device (cached per device: a process may drive several GPUs)
static int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cached[dev & 63];
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace pgrf

namespace pgrf {
int launch_render_mlp_bf16(const pgrf_render_args& a, int V, int T, long long total, int n_tiles, int Mv, int sms, cudaStream_t st);
int launch_render_rays_bf16(const pgrf_render_args& a, int V, int T, long long total, int sms, cudaStream_t st);
int launch_render_rays_tc(const pgrf_render_args& a, int V, long long total, int sms, cudaStream_t st);
int g_rays_tc = 1;     // debug knob "rays_tc": 0 forces the SIMT-attention rays kernel for every dn
}
using namespace pgrf;

extern "C" int pgrf_weight_blob_floats(void) { return kBlobFloats; }
extern "C" int pgrf_weight_num_layers(void) { return kNumLayers; }
extern "C" int pgrf_weight_layer_info(int i, char* name, int name_cap, int* K, int* N, int* Npad, int* has_bias,
                                      int* k_begin, int* w_offset, int* b_offset) {
  PGRF_REQUIRE(i >= 0 && i < kNumLayers, "layer index %d out of range", i);
  snprintf(name, name_cap, "%s", kLayers[i].name);
  *K = kLayers[i].K; *N = kLayers[i].N; *Npad = kLayers[i].Npad; *has_bias = kLayers[i].has_bias;
  *k_begin = kLayers[i].k_begin;
  *w_offset = layer_offset(i);
  *b_offset = layer_offset(i) + kLayers[i].K * kLayers[i].Npad;
  return PGRF_OK;
}
extern "C" int pgrf_weight_aux_offsets(int* layer_norm_offset, int* posenc_offset, int* max_samples) {
  *layer_norm_offset = kLnOffset; *posenc_offset = kPosencOffset; *max_samples = kMaxSamplesPerRay;
  return PGRF_OK;
}

extern "C" int pgrf_render_workspace(int rfn, long long n_samples, long long* f1_floats, long long* f2_floats) {
  const int T = tile_samples(rfn);
  PGRF_REQUIRE(T >= 32, "render: rfn=%d source views unsupported (1..4)", rfn);
  const long long tiles = (n_samples + T - 1) / T;
  *f1_floats = tiles * kF1 * kTileRows;
  *f2_floats = tiles * kF2 * T;
  // bf16 path: operand tiles of the rays kernel (<= 128 samples each, > 64 used) + one float4 per sample
  const long long f2_bf16 = ((n_samples / 65 + 2) * kF2TileBytes + n_samples * 16 + 3) / 4;
  if (f2_bf16 > *f2_floats) *f2_floats = f2_bf16;
  return PGRF_OK;
}

extern "C" int pgrf_render_pass_fwd(const pgrf_render_args* args, void* stream) {
  PGRF_REQUIRE(args != nullptr, "render: null args");
  const pgrf_render_args& a = *args;
  PGRF_REQUIRE(a.dataset >= 0 && a.dataset <= 3, "render: unknown dataset id %d", a.dataset);
  PGRF_REQUIRE(a.rfn >= 1 && a.rfn <= 4, "render: rfn=%d source views unsupported (1..4)", a.rfn);
  PGRF_REQUIRE(a.dn >= 3 && a.dn <= kMaxSamplesPerRay, "render: dn=%d samples per ray out of [3,%d]", a.dn, kMaxSamplesPerRay);
  PGRF_REQUIRE(a.rn >= 1, "render: rn=%d", a.rn);
  PGRF_REQUIRE(a.H > 1 && a.W > 1 && a.img_h > 1 && a.img_w > 1 && a.if_h > 0 && a.if_w > 0 && a.rf_h > 0 && a.rf_w > 0,
               "render: bad map sizes");
  const bool dict_in = a.prj_in || a.feat_in || a.prob_in;
  PGRF_REQUIRE(!(dict_in || a.dec_dbg) || !a.mlp_bf16, "render: prj_in / feat_in / prob_in / dec_dbg exist on the fp32 path only");
  PGRF_REQUIRE(!a.prj_in || (a.que_dir_in && (a.interval_in || a.prob_in || a.depth)),
               "render: prj_in needs que_dir_in and interval_in (or depth, or prob_in)");
  PGRF_REQUIRE(a.prj_in || ((a.coords || a.ray_dirs) && a.que_c2w && a.ref_w2c), "render: null pointer argument (coords / que_c2w / ref_w2c)");
  PGRF_REQUIRE(a.feat_in || (a.imgs_cl && a.img_feats_cl && a.ray_feats_cl), "render: null pointer argument (source maps)");
  PGRF_REQUIRE(a.depth || (a.prj_in && (a.interval_in || a.prob_in) && !a.render_depth && !a.fine_depth),
               "render: depth may only be omitted with prj_in + interval_in/prob_in and without render_depth / fine sampling");
  PGRF_REQUIRE(a.prob_in || a.ref_depth_range, "render: null pointer argument (ref_depth_range)");
  PGRF_REQUIRE(a.weights && (a.f1 || a.mlp_bf16) && a.f2 && a.pixel_colors, "render: null pointer argument");
  PGRF_REQUIRE(a.depth_ray_stride == 0 || a.depth_ray_stride == a.dn, "render: depth_ray_stride must be 0 or dn");
  PGRF_REQUIRE(!a.fine_depth || (a.fine_u && a.fine_dn >= 1 && a.fine_dn + (a.fine_use_all ? a.dn : 0) <= 2 * kMaxSamplesPerRay - 2),
               "render: bad fine sampling arguments");
  PGRF_REQUIRE((((uintptr_t)a.f1 | (uintptr_t)a.f2 | (uintptr_t)a.weights | (uintptr_t)a.imgs_cl |
                 (uintptr_t)a.img_feats_cl | (uintptr_t)a.ray_feats_cl) & 15) == 0,
               "render: maps, weights and workspaces must be 16-byte aligned");
  RenderParams p;
  p.a = a;
  p.V = a.rfn;
  p.T = tile_samples(a.rfn);
  p.M = p.V * p.T;
  p.total = (long long)a.rn * a.dn;
  p.n_tiles = (int)((p.total + p.T - 1) / p.T);
  p.rays_per_tile3 = a.dn >= kTileRows ? 1 : kTileRows / a.dn;
  p.n_tiles3 = (a.rn + p.rays_per_tile3 - 1) / p.rays_per_tile3;
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = num_sms();
  const size_t s1 = R1_FLOATS * sizeof(float), s2 = R2_FLOATS * sizeof(float), s3 = R3_FLOATS * sizeof(float);
  {   // the > 48 KB dynamic shared memory opt-in is a per-device function attribute
    static bool attr_done[64] = {};
    int dev = 0;
    PGRF_CUDA(cudaGetDevice(&dev));
    if (!attr_done[dev & 63]) {
      PGRF_CUDA(cudaFuncSetAttribute(render_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
      PGRF_CUDA(cudaFuncSetAttribute(render_samples_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
      PGRF_CUDA(cudaFuncSetAttribute(render_rays_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
      attr_done[dev & 63] = true;
    }
  }
  const int mask = a.stage_mask ? a.stage_mask : 7;
  if (a.mlp_bf16) {
    if (a.sched) PGRF_CUDA(cudaMemsetAsync(a.sched, 0, 2 * sizeof(int), st));
    const int T16 = tile_samples16(a.rfn);
    const int Mv = (a.dn >= kTileRows ? 1 : kTileRows / a.dn) * a.dn;     // samples per tile of the rays kernel (whole rays)
    if (mask & 3) {
      const int rc = launch_render_mlp_bf16(a, p.V, T16, p.total, (int)((p.total + T16 - 1) / T16), Mv, sms, st);
      if (rc != PGRF_OK) return rc;
    }
    if (mask & 4) {
      // attention on the tensor cores for the production sample counts (whole rays of 64 or 128 samples per 128-row tile)
      const int rc = (g_rays_tc && (a.dn == 64 || a.dn == 128)) ? launch_render_rays_tc(a, p.V, p.total, sms, st)
                                                                : launch_render_rays_bf16(a, p.V, T16, p.total, sms, st);
      if (rc != PGRF_OK) return rc;
    }
    return PGRF_OK;
  }
  if (mask & 1) { render_rows_kernel<<<min(p.n_tiles, sms), kThreads, s1, st>>>(p); count_launch(); }
  if (mask & 2) { render_samples_kernel<<<min(p.n_tiles, sms), kThreads, s2, st>>>(p); count_launch(); }
  if (mask & 4) { render_rays_kernel<<<min(p.n_tiles3, sms), kThreads, s3, st>>>(p); count_launch(); }
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_agg_mlp_fwd(const pgrf_render_args* args, void* stream) {
  PGRF_REQUIRE(args != nullptr, "agg_mlp: null args");
  PGRF_REQUIRE(args->prj_in && args->feat_in && args->prob_in && args->que_dir_in,
               "agg_mlp: prj_in, feat_in, prob_in and que_dir_in are required");
  return pgrf_render_pass_fwd(args, stream);
}

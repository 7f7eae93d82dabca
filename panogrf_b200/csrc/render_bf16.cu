// bf16 tensor-core variant of the per-(view,sample) and per-sample stages of the render path:
// rows kernel + samples kernel of render_kernels.cu fused into ONE persistent tcgen05 kernel.
//
//   * CTA = 2 independent warpgroups (128 threads each); a warpgroup owns one tile of 128 (view,sample) rows at a
//     time: thread <-> row for geometry, every epilogue and all element-wise math (state stays in registers);
//   * every Linear layer is a tcgen05.mma (kind::f16, bf16 operands in shared memory, fp32 accumulators in TMEM,
//     M = 128 rows, N = 16..64, K = 16..240) issued by one thread of the warpgroup and completed through
//     tcgen05.commit -> mbarrier; the epilogue reads its row with tcgen05.ld (32x32b) and writes the next layer's
//     A operand with one conflict-free 128-bit store per 8 features (k-chunk-major layout, umma.cuh);
//   * all 27 layers' bf16 weights (76 KB) stay resident in shared memory; the F1 inter-kernel tiles of the fp32
//     path never exist; only the pooled per-sample features (F2: 68 floats / sample) go to HBM for the rays kernel.
// Reference semantics: same as render_rows_kernel + render_samples_kernel (see render_kernels.cu header).
// Numerics: operands rounded to bf16, accumulation and all non-GEMM math in fp32 (north star: rtol 1e-2).
#include <type_traits>

#include "render_device.cuh"
#include "render_layout16.cuh"
#include "umma.cuh"

#define W16(L) (std::integral_constant<int, pgrf::w16_offset(L)>::value)
#define B16(L) (std::integral_constant<int, pgrf::b16_offset(L)>::value)
#define SMW(L) (std::integral_constant<int, pgrf::small16_offset(L)>::value)
#define SMB(L) (std::integral_constant<int, pgrf::small16_bias_offset(L)>::value)

namespace pgrf {

constexpr int kWG = 3;                      // warpgroups per CTA (12 warps; shared memory: 72 KB weights + 3 x 44 KB)
constexpr int kThreads16 = 128 * kWG;
constexpr int ROWS = 128;                   // rows per tile == operand row pitch
constexpr int CH = ROWS * 16;               // bytes per k-chunk of an A operand

// ---- per-warpgroup shared memory map (bytes): two 10-chunk operand regions + 8 float vectors ----
//  E early : RF = chunks 0..5 (ray_feats 4 | [hit',vis',0..] | zero; chunks 4,5 hold ray_dir_fc.0's output until compute_prob)
//            HDA = chunks 6..9 (decoder / prob_embed hidden)
//  E mid   : pooled view statistics, 10 chunks at a time (mean0|var0, then mean1|var1) -> base_fc.0 K-slices
//  E late  : H64 = 0..7 ; HV = 0..3 ; HV2 = 4..7 ; RG = 4..9
//  P early : img_feats 0..3 | rgb 4 | HDB 5..8 (second decoder hidden buffer) | zero 9
//  P mid   : rgb_feat' 0..4 | prob_embedding 5..8 | zero 9   == last 80 K-columns of base_fc.0
//  P late  : x in fp32 [32][128] (chunks 0..7)
constexpr int E_BYTES = 10 * CH;
constexpr int P_BYTES = 10 * CH;
constexpr int S_OFF = E_BYTES + P_BYTES;
constexpr int S_BYTES = 8 * ROWS * 4;       // SF vectors; all 8 double as the 2 x 128 footprint records during geometry/gather
constexpr int WG_BYTES = S_OFF + S_BYTES;
constexpr int E_RF = 0, E_RDH = 4 * CH, E_HDA = 6 * CH;
constexpr int E_H64 = 0, E_HV = 0, E_HV2 = 4 * CH, E_RG = 4 * CH;
constexpr int P_IMG = 0, P_RGB = 4 * CH, P_NEU = 5 * CH, P_HDB = 5 * CH, P_ZERO = 9 * CH;
constexpr int S_F = 0;
enum { SF_PX = 0, SF_PY, SF_W0, SF_VIS2, SF_LOGIT, SF_R, SF_G, SF_B };
constexpr int kTmemPerWG = 128;

constexpr int SM16_W = 0;
constexpr int SM16_WG = (kW16Sec0Bytes + 127) & ~127;
constexpr int SM16_BAR = SM16_WG + kWG * WG_BYTES;
constexpr int SM16_BYTES = SM16_BAR + 64;

// softplus for the bf16 path: log(1 + e^x) through lg2.approx (absolute error ~1e-7; the fp32 path keeps log1pf)
__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + fast_exp(x)); }
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

// ---- epilogues (thread = row m).  Kept out of line and rolled: the kernel is a long straight-line sequence of
// stages executed once per tile by only 8 warps, so instruction-cache footprint matters more than unrolling. ----
// dst chunks [0,nchunks) = bf16( act(acc + bias) )
// acc[16] (+bias) -> activation, as 8 packed fp32 pairs (FADD2 / FMUL2: one issue slot per two elements)
__device__ __forceinline__ void act_pairs(float (&v)[16], const float4 (&bb)[4], int act) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b = bb[i];
    const float2 lo = fadd2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
    const float2 hi = fadd2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
    v[4 * i] = lo.x; v[4 * i + 1] = lo.y; v[4 * i + 2] = hi.x; v[4 * i + 3] = hi.y;
  }
  if (act == ACT_ELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 r = elu_pair(make_float2(v[2 * i], v[2 * i + 1]));
      v[2 * i] = r.x; v[2 * i + 1] = r.y;
    }
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}
static __device__ __noinline__ void epi_act_store(uint32_t taddr, const float* __restrict__ bias, unsigned char* dst, int m,
                                                  int nchunks, int act) {
  // two chunks (16 accumulator columns) per TMEM round trip; nchunks is even for every caller
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2) {
    float4 bb[4];   // issued before the TMEM load: the tcgen05.ld / wait::ld asm is a compiler barrier for shared-memory loads
#pragma unroll
    for (int i = 0; i < 4; ++i) bb[i] = reinterpret_cast<const float4*>(bias + 8 * c)[i];
    float v[16];
    umma::ld16(taddr + 8 * c, v);
    if (act == ACT_ELU) {
      uint32_t q[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 b = bb[i];
        q[2 * i] = elu_pack(fadd2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y)));
        q[2 * i + 1] = elu_pack(fadd2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w)));
      }
      *reinterpret_cast<uint4*>(dst + ((size_t)c * ROWS + m) * 16) = make_uint4(q[0], q[1], q[2], q[3]);
      *reinterpret_cast<uint4*>(dst + ((size_t)(c + 1) * ROWS + m) * 16) = make_uint4(q[4], q[5], q[6], q[7]);
      continue;
    }
    act_pairs(v, bb, act);
    umma::store_chunk(dst, ROWS, c, m, v);
    umma::store_chunk(dst, ROWS, c + 1, m, v + 8);
  }
}
// dst chunks (optional) = bf16(h), h = act(acc + bias); out[j] = sum_k h[k] * Wsm[j][k]  (tiny fp32 output layer fused
// into the epilogue of the layer that feeds it: saves one MMA stage and keeps its input in fp32)
template <int NOUT>
static __device__ __noinline__ void epi_act_gemv(uint32_t taddr, const float* __restrict__ bias, unsigned char* dst, int m,
                                                 int nchunks, int act, const float* __restrict__ Wsm, int K, float (&out)[NOUT]) {
  float2 acc[NOUT];                      // even / odd k partial sums (FFMA2)
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2) {
    float4 bb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bb[i] = reinterpret_cast<const float4*>(bias + 8 * c)[i];
    float v[16];
    umma::ld16(taddr + 8 * c, v);
    act_pairs(v, bb, act);
    if (dst) {
      umma::store_chunk(dst, ROWS, c, m, v);
      umma::store_chunk(dst, ROWS, c + 1, m, v + 8);
    }
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float4* w4 = reinterpret_cast<const float4*>(Wsm + j * K + 8 * c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = w4[i];
        acc[j] = ffma2(make_float2(v[4 * i], v[4 * i + 1]), make_float2(w.x, w.y), acc[j]);
        acc[j] = ffma2(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(w.z, w.w), acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) out[j] = acc[j].x + acc[j].y;
}
// first accumulator column (+bias) of a padded N=16 output layer
__device__ __forceinline__ float epi_scalar(uint32_t taddr, const float* __restrict__ bias) {
  float a, b;
  umma::ld2(taddr, a, b);
  return a + bias[0];
}
__device__ __forceinline__ void zero_chunk(unsigned char* dst, int chunk, int m) {
  *reinterpret_cast<uint4*>(dst + ((size_t)chunk * ROWS + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
}

struct __align__(16) FootRec {
  int off;      // texel index of the north-west tap inside the stacked (rfn*h*w) map
  int dxy;      // bit0: east neighbour inside the map, bit1: south neighbour inside the map
  float tx, ty;
};
__device__ __forceinline__ float4 tap4_rec(const float4* __restrict__ base, const FootRec& f, int stride_x, int stride_y) {
  const int sx = (f.dxy & 1) ? stride_x : 0, sy = (f.dxy & 2) ? stride_y : 0;
  const float4 nw = ldg4(base), ne = ldg4(base + sx), sw = ldg4(base + sy), se = ldg4(base + sy + sx);
  const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
  const float wnw = tx1 * ty1, wne = f.tx * ty1, wsw = tx1 * f.ty, wse = f.tx * f.ty;
  // same products and accumulation order (nw, ne, sw, se) as the scalar form, two channels per instruction
  const float2 w0 = make_float2(wnw, wnw), w1 = make_float2(wne, wne), w2 = make_float2(wsw, wsw), w3 = make_float2(wse, wse);
  float2 lo = fmul2(make_float2(nw.x, nw.y), w0), hi = fmul2(make_float2(nw.z, nw.w), w0);
  lo = ffma2(make_float2(ne.x, ne.y), w1, lo); hi = ffma2(make_float2(ne.z, ne.w), w1, hi);
  lo = ffma2(make_float2(sw.x, sw.y), w2, lo); hi = ffma2(make_float2(sw.z, sw.w), w2, hi);
  lo = ffma2(make_float2(se.x, se.y), w3, lo); hi = ffma2(make_float2(se.z, se.w), w3, hi);
  const float4 o = make_float4(lo.x, lo.y, hi.x, hi.y);
  return o;
}

// one MMA stage: publish the operand writes, sync the warpgroup, let one thread issue, wait for completion
#define STAGE_BEGIN()                                                       \
  umma::fence_smem_to_async(); umma::fence_before_sync(); wg_sync(wg);      \
  if (m == 0) { umma::fence_after_sync();
#define STAGE_END()                                                         \
    umma::commit(bar); }                                                    \
  mbar_wait(bar, phase); phase ^= 1; umma::fence_after_sync();

// weighted mean / variance over the V views of sample t (fused_mean_variance, ibrnet.py:112-116) of the 40-wide
// rgb_feat' block in P (chunks 0..4) -> E chunks 0..4 (mean) and 5..9 (variance), written to row m
template <int V>
__device__ __forceinline__ void pool_views(const unsigned char* P, unsigned char* E, int t, int T, int v, const float (&w)[V]) {
  // the V threads of a sample split the 40 channels in units of 4 (half a k-chunk) and each writes its results to all
  // V rows (every (view, sample) row of base_fc.0 sees the same pooled statistics); packed fp32 pairs (FFMA2)
  if (v >= V) return;
#pragma unroll 1
  for (int u = v; u < 10; u += V) {
    const int c = u >> 1, hb = (u & 1) * 8;                      // chunk, byte offset of the half inside the 16-byte row
    float2 x[V][2];
#pragma unroll
    for (int vv = 0; vv < V; ++vv) {
      const uint2 q = *reinterpret_cast<const uint2*>(P + ((size_t)c * ROWS + vv * T + t) * 16 + hb);
      x[vv][0] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
      x[vv][1] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
    }
    uint2 mu, var;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float2 a0 = fmul2(x[0][i], make_float2(w[0], w[0]));
#pragma unroll
      for (int vv = 1; vv < V; ++vv) a0 = ffma2(x[vv][i], make_float2(w[vv], w[vv]), a0);
      const float2 na = make_float2(-a0.x, -a0.y);
      float2 b0 = make_float2(0.f, 0.f);
#pragma unroll
      for (int vv = 0; vv < V; ++vv) {
        const float2 d = fadd2(x[vv][i], na);
        b0 = ffma2(make_float2(w[vv], w[vv]), fmul2(d, d), b0);
      }
      (i == 0 ? mu.x : mu.y) = umma::pack2(a0.x, a0.y);
      (i == 0 ? var.x : var.y) = umma::pack2(b0.x, b0.y);
    }
#pragma unroll
    for (int vv = 0; vv < V; ++vv) {
      *reinterpret_cast<uint2*>(E + ((size_t)c * ROWS + vv * T + t) * 16 + hb) = mu;
      *reinterpret_cast<uint2*>(E + ((size_t)(5 + c) * ROWS + vv * T + t) * 16 + hb) = var;
    }
  }
}

struct Render16Params {
  pgrf_render_args a;
  int V, T, M;
  long long total;
  int n_tiles;
};

template <int V>
__global__ void __launch_bounds__(kThreads16, 1) render_mlp_bf16_kernel(const Render16Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x;
  const int wg = tid >> 7;            // warpgroup
  const int m = tid & 127;            // row inside the warpgroup's tile
  const int wq = (tid >> 5) & 3;      // warp inside the warpgroup -> TMEM lane quadrant
  unsigned char* Wb = smem + SM16_W;
  const float* Bias = reinterpret_cast<const float*>(Wb + kW16WeightBytes);
  const float* Wsm = reinterpret_cast<const float*>(Wb + kW16SmallBegin0);     // fp32 tiny output layers
  unsigned char* E = smem + SM16_WG + wg * WG_BYTES;
  unsigned char* P = E + E_BYTES;
  unsigned char* S = E + S_OFF;
  float* SF = reinterpret_cast<float*>(S + S_F);
  FootRec* FP = reinterpret_cast<FootRec*>(SF);   // [2][128] records = SF vectors 0..7 (all dead during geometry/gather)
  float* XF = reinterpret_cast<float*>(P);        // x in fp32 [32][128]; P is dead once base_fc.0 has consumed it
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM16_BAR) + wg;

  // one-time setup: weights -> smem, TMEM allocation (512 columns: 256 per warpgroup), mbarriers
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights16);
    uint4* dst = reinterpret_cast<uint4*>(Wb);
    for (int i = tid; i < kW16Sec0Bytes / 16; i += kThreads16) dst[i] = __ldg(src + i);
  }
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 512);
  if (m == 0) mbar_init(bar, 1);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s + wg * kTmemPerWG;          // this warpgroup's TMEM columns
  const uint32_t tq = tb + ((uint32_t)(wq * 32) << 16);       // + this warp's lane quadrant
  uint32_t phase = 0;

  const int T = p.T, M = p.M;
  const int v = min(m / T, V - 1), t = m % T;
  const int vrow = m / T;             // unclamped: rows beyond the V*T valid ones take no share of the pooling work
  const float wgt = 1.f / ((float)V + 1e-8f);

  __shared__ int s_tile[kWG];
  int static_tile = blockIdx.x * kWG + wg;
  // per-thread constants of the whole launch (the view of a row never changes: v = m / T)
  const float inv_wm1 = 1.f / (float)(a.img_w - 1), inv_hm1 = 1.f / (float)(a.img_h - 1);
  const float q_nn = -1.f / a.que_near, q_inv = 1.f / (-1.f / a.que_far - q_nn);
  const float r_nn = -1.f / __ldg(a.ref_depth_range + 2 * v), r_inv = 1.f / (-1.f / __ldg(a.ref_depth_range + 2 * v + 1) - r_nn);
#pragma unroll 1
  while (true) {
    // dynamic tile scheduler (one atomic per 128-row tile) when the caller provides a counter, else static striding
    int tile;
    if (a.sched) {
      if (m == 0) s_tile[wg] = atomicAdd(a.sched + 0, 1);
      wg_sync(wg);
      tile = s_tile[wg];
    } else {
      tile = static_tile;
      static_tile += gridDim.x * kWG;
    }
    if (tile >= p.n_tiles) break;
    long long g = (long long)tile * T + t;
    const bool row_valid = (m < M) && (g < p.total);
    if (g >= p.total) g = p.total - 1;
    const int ray = (int)((unsigned)g / (unsigned)a.dn), s = (int)((unsigned)g - (unsigned)ray * (unsigned)a.dn);   // total < 2^31

    // ------------------------------------------------------------ geometry (thread = row)
    const RowGeom rg = row_geometry<true>(a, v, g);
    {
      Footprint f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, a.rf_h == a.img_h && a.rf_w == a.img_w, a.rf_h, a.rf_w);
      FootRec r1; r1.off = v * a.rf_h * a.rf_w + f.off; r1.dxy = f.dx | (f.dy << 1); r1.tx = f.tx; r1.ty = f.ty;
      FP[m] = r1;
      f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, a.if_h == a.img_h && a.if_w == a.img_w, a.if_h, a.if_w);
      r1.off = v * a.if_h * a.if_w + f.off; r1.dxy = f.dx | (f.dy << 1); r1.tx = f.tx; r1.ty = f.ty;
      FP[ROWS + m] = r1;
    }
    {   // ray_dir_fc.0 (4 -> 16, ELU) as an fp32 register GEMV; its output is the K = 16 operand of ray_dir_fc.2
      float h[16];
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        const float4 w = *reinterpret_cast<const float4*>(Wsm + SMW(M_RD0) + 4 * n);
        h[n] = elu1(Wsm[SMB(M_RD0) + n] + rg.dirdiff[0] * w.x + rg.dirdiff[1] * w.y + rg.dirdiff[2] * w.z + rg.dirdiff[3] * w.w);
      }
      umma::store_chunk(E + E_RDH, ROWS, 0, m, h);
      umma::store_chunk(E + E_RDH, ROWS, 1, m, h + 8);
    }
    // own-row colour taps (fp32, kept in registers for the final blend)
    float rgb_in[3];
    {
      const Footprint f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, true, a.img_h, a.img_w);
      const float4 c = tap4(reinterpret_cast<const float4*>(a.imgs_cl) + (size_t)v * a.img_h * a.img_w + f.off, f, 1, a.img_w);
      rgb_in[0] = c.x; rgb_in[1] = c.y; rgb_in[2] = c.z;
    }
    // sampling interval of this sample along its ray (depth2inv_dists) and the view's normalised depth
    float d_prev, d_s, dv;
    {
      // normalised inverse depth (-1/d - nn) / (ff - nn) with the range constants hoisted and one MUFU.RCP per depth
      const float* dp = a.depth + (size_t)ray * a.depth_ray_stride;
      const float i_s = (-fast_rcp(__ldg(dp + s)) - q_nn) * q_inv;
      d_s = (s + 1 < a.dn) ? (-fast_rcp(__ldg(dp + s + 1)) - q_nn) * q_inv - i_s : 1e6f;
      d_prev = d_s;
      if (s > 0) d_prev = i_s - (-fast_rcp(__ldg(dp + s - 1)) - q_nn) * q_inv;
      dv = (-fast_rcp(fmaxf(rg.pdepth, 1e-5f)) - r_nn) * r_inv;
    }
    wg_sync(wg);

    // ------------------------------------------------------------ cooperative gathers (lane = (row, float4 group))
    // each row's two bilinear footprints were computed once by its own thread (FP records in smem); here 8 lanes
    // per row fetch the 4 x 128-byte taps of both feature maps and blend
#pragma unroll 2
    for (int it = m; it < ROWS * 8; it += 128) {
      const int r = it >> 3, cg = it & 7;
      const FootRec f1 = FP[r], f2 = FP[ROWS + r];
      const float4* b1 = reinterpret_cast<const float4*>(a.ray_feats_cl) + (size_t)f1.off * 8 + cg;
      const float4* b2 = reinterpret_cast<const float4*>(a.img_feats_cl) + (size_t)f2.off * 8 + cg;
      const float4 rf = tap4_rec(b1, f1, 8, a.rf_w * 8);
      const float4 imf = tap4_rec(b2, f2, 8, a.if_w * 8);
      uint2 q;   // 4 channels = half a chunk: chunk cg/2, 8-byte half cg%2
      q.x = umma::pack2(rf.x, rf.y); q.y = umma::pack2(rf.z, rf.w);
      *reinterpret_cast<uint2*>(E + E_RF + ((size_t)(cg >> 1) * ROWS + r) * 16 + (cg & 1) * 8) = q;
      q.x = umma::pack2(imf.x, imf.y); q.y = umma::pack2(imf.z, imf.w);
      *reinterpret_cast<uint2*>(P + P_IMG + ((size_t)(cg >> 1) * ROWS + r) * 16 + (cg & 1) * 8) = q;
    }
    zero_chunk(P, 9, m);

    // ------------------------------------------------------------ stage A: mean_decoder.0 + ray_dir_fc.2
    STAGE_BEGIN()
      umma::gemm_issue(tb + 0, E + E_RF, ROWS, Wb + W16(M_MEAN0), 32, 32, 32);
      umma::gemm_issue(tb + 64, E + E_RDH, ROWS, Wb + W16(M_RD1), 48, 48, 16);
    STAGE_END()
    epi_act_store(tq + 0, Bias + B16(M_MEAN0), E + E_HDA, m, 4, ACT_ELU);
    // direction feature (f' order: img_feats 0..31, rgb 32..34): rgb_feat = [img_feats, rgb] + ELU(ray_dir_fc)
#pragma unroll 1
    for (int c = 0; c < 5; ++c) {
      float df[8], x[8];
      umma::ld8(tq + 64 + 8 * c, df);
      if (c < 4) {
        umma::load_chunk(P + P_IMG, ROWS, c, m, x);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = i < 3 ? rgb_in[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 bb = *reinterpret_cast<const float2*>(Bias + B16(M_RD1) + 8 * c + i);
        const float2 d = elu_pair(fadd2(make_float2(df[i], df[i + 1]), bb));
        const float2 xs = fadd2(make_float2(x[i], x[i + 1]), d);
        x[i] = (c < 4 || i < 3) ? xs.x : 0.f;
        x[i + 1] = (c < 4 || i + 1 < 3) ? xs.y : 0.f;
      }
      umma::store_chunk(P, ROWS, c, m, x);     // P_IMG chunks 0..3, P_RGB = chunk 4
    }
    // ------------------------------------------------------------ stages B, C, D: the decoders, software-pipelined over two
    // hidden buffers (HDA in E, HDB in P); each output layer (32 -> 2, 2, 1) is an fp32 register GEMV fused into the epilogue
    float mean[2], var[2], aw, visd = 1.f;
    STAGE_BEGIN()
      umma::gemm_issue(tb + 0, E + E_HDA, ROWS, Wb + W16(M_MEAN1), 32, 32, 32);
      umma::gemm_issue(tb + 32, E + E_RF, ROWS, Wb + W16(M_VAR0), 32, 32, 32);
    STAGE_END()
    {
      float o2[2];
      epi_act_gemv<2>(tq + 0, Bias + B16(M_MEAN1), nullptr, m, 4, ACT_ELU, Wsm + SMW(M_MEAN2), 32, o2);
      mean[0] = softplus_fast(o2[0] + Wsm[SMB(M_MEAN2)]); mean[1] = softplus_fast(o2[1] + Wsm[SMB(M_MEAN2) + 1]);
      epi_act_store(tq + 32, Bias + B16(M_VAR0), P + P_HDB, m, 4, ACT_ELU);
    }
    STAGE_BEGIN()
      umma::gemm_issue(tb + 0, P + P_HDB, ROWS, Wb + W16(M_VAR1), 32, 32, 32);
      umma::gemm_issue(tb + 32, E + E_RF, ROWS, Wb + W16(M_AW0), 32, 32, 32);
    STAGE_END()
    {
      float o2[2];
      epi_act_gemv<2>(tq + 0, Bias + B16(M_VAR1), nullptr, m, 4, ACT_ELU, Wsm + SMW(M_VAR2), 32, o2);
      var[0] = softplus_fast(o2[0] + Wsm[SMB(M_VAR2)]) + a.bias_val; var[1] = softplus_fast(o2[1] + Wsm[SMB(M_VAR2) + 1]) + a.bias_val;
      epi_act_store(tq + 32, Bias + B16(M_AW0), E + E_HDA, m, 4, ACT_ELU);
    }
    STAGE_BEGIN()
      umma::gemm_issue(tb + 0, E + E_HDA, ROWS, Wb + W16(M_AW1), 32, 32, 32);
      if (a.use_vis) umma::gemm_issue(tb + 32, E + E_RF, ROWS, Wb + W16(M_VIS0), 32, 32, 32);
    STAGE_END()
    {
      float o1[1];
      epi_act_gemv<1>(tq + 0, Bias + B16(M_AW1), nullptr, m, 4, ACT_ELU, Wsm + SMW(M_AW2), 32, o1);
      aw = sigmoidf(o1[0] + Wsm[SMB(M_AW2)]);
    }
    if (a.use_vis) {   // 4th decoder
      epi_act_store(tq + 32, Bias + B16(M_VIS0), P + P_HDB, m, 4, ACT_ELU);
      STAGE_BEGIN() umma::gemm_issue(tb + 0, P + P_HDB, ROWS, Wb + W16(M_VIS1), 32, 32, 32); STAGE_END()
      float o1[1];
      epi_act_gemv<1>(tq + 0, Bias + B16(M_VIS1), nullptr, m, 4, ACT_ELU, Wsm + SMW(M_VIS2), 32, o1);
      visd = sigmoidf(o1[0] + Wsm[SMB(M_VIS2)]);
    }

    // ------------------------------------------------------------ logistic-mixture probabilities (dist_decoder.compute_prob)
    {
      const float nearp = dv - d_prev / 2.f, farp = dv + d_s / 2.f;
      const float mix[2] = {aw, 1.f - aw};
      float visibility = 0.f, hp = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        // 0.5 + 0.5 tanh(x) == sigmoid(2x): one ex2 + one rcp (~3e-7 relative) instead of two libm tanhf
        float cdf0 = sigmoidf(2.f * ((nearp - mean[j]) * var[j]));
        float cdf1 = sigmoidf(2.f * ((farp - mean[j]) * var[j]));
        if (a.use_vis) { cdf0 *= visd; cdf1 *= visd; }
        visibility += (1.f - cdf0) * mix[j];
        hp += (cdf1 - cdf0) * mix[j];
      }
      if (a.prob_dbg && row_valid) {
        float* d = a.prob_dbg + ((size_t)v * p.total + g) * 3;
        d[0] = logf(hp / (visibility - hp + 1e-5f) + 1e-5f); d[1] = visibility; d[2] = hp;
      }
      float hv8[8] = {(hp - 0.5f) * 2.f, (visibility - 0.5f) * 2.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      umma::store_chunk(E + E_RF, ROWS, 4, m, hv8);     // overwrites ray_dir_fc.0's output (its MMA completed in stage A)
      zero_chunk(E + E_RF, 5, m);
    }

    // ------------------------------------------------------------ prob_embed 34 -> 32 (ReLU) -> 32, neuray_fc fused
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E + E_RF, ROWS, Wb + W16(M_PE0), 32, 32, 48); STAGE_END()
    epi_act_store(tq + 0, Bias + B16(M_PE0), E + E_HDA, m, 4, ACT_RELU);
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E + E_HDA, ROWS, Wb + W16(M_PE1), 32, 32, 32); STAGE_END()
    {   // prob_embedding -> P (bf16 operand of base_fc.0) and, fused, neuray_fc 32 -> 8 (ELU) -> 1 (sigmoid) in fp32
      float h8[8];
      epi_act_gemv<8>(tq + 0, Bias + B16(M_PE1), P + P_NEU, m, 4, ACT_NONE, Wsm + SMW(M_NF0), 32, h8);
      float o = Wsm[SMB(M_NF1)];
#pragma unroll
      for (int i = 0; i < 8; ++i) o = fmaf(elu1(h8[i] + Wsm[SMB(M_NF0) + i]), Wsm[SMW(M_NF1) + i], o);
      SF[SF_W0 * ROWS + m] = sigmoidf(o);
    }
    // ------------------------------------------------------------ base_fc.0 in three accumulating K-slices:
    //   (a) per-row slice [rgb_feat' | prob_embedding | 0] = P (K = 80), issued now;
    //   (b1) pooled [mean0 | var0], (b2) pooled [mean1 | var1] (K = 80 each) through the 10-chunk E region
    umma::fence_smem_to_async(); umma::fence_before_sync(); wg_sync(wg);
    if (m == 0) {
      umma::fence_after_sync();
      umma::gemm_issue(tb + 0, P, ROWS, Wb + W16(M_BASE0) + 20 * 64 * 16, 64, 64, 80);
    }
    float w0n[V];
#pragma unroll
    for (int vv = 0; vv < V; ++vv) w0n[vv] = SF[SF_W0 * ROWS + vv * T + t] * wgt;
    pool_views<V>(P, E, t, T, vrow, w0n);
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E, ROWS, Wb + W16(M_BASE0), 64, 64, 80, true); STAGE_END()
#pragma unroll
    for (int vv = 0; vv < V; ++vv) w0n[vv] = wgt;
    pool_views<V>(P, E, t, T, vrow, w0n);
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E, ROWS, Wb + W16(M_BASE0) + 10 * 64 * 16, 64, 64, 80, true); STAGE_END()
    epi_act_store(tq + 0, Bias + B16(M_BASE0), E + E_H64, m, 8, ACT_ELU);

    // ------------------------------------------------------------ base_fc.2 -> x (fp32, column m of XF = P region)
    STAGE_BEGIN() umma::gemm_issue(tb + 64, E + E_H64, ROWS, Wb + W16(M_BASE1), 32, 32, 64); STAGE_END()
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      float r[8], sx[8];
      umma::ld8(tq + 64 + 8 * c, r);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 bb = *reinterpret_cast<const float2*>(Bias + B16(M_BASE1) + 8 * c + i);
        const float2 xv = elu_pair(fadd2(make_float2(r[i], r[i + 1]), bb));
        XF[(8 * c + i) * ROWS + m] = xv.x;
        XF[(8 * c + i + 1) * ROWS + m] = xv.y;
        const float2 sv = fmul2(xv, make_float2(wgt, wgt));
        sx[i] = sv.x; sx[i + 1] = sv.y;
      }
      umma::store_chunk(E + E_HV, ROWS, c, m, sx);      // H64 is dead: its MMA completed
    }
    // ------------------------------------------------------------ vis_fc(x * weight) 32 -> 32 -> 33
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E + E_HV, ROWS, Wb + W16(M_VFC0), 32, 32, 32); STAGE_END()
    epi_act_store(tq + 0, Bias + B16(M_VFC0), E + E_HV2, m, 4, ACT_ELU);
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E + E_HV2, ROWS, Wb + W16(M_VFC1), 48, 48, 32); STAGE_END()
    {
      const float vis1 = sigmoidf(elu1(epi_scalar(tq + 32, Bias + B16(M_VFC1) + 32)));   // vis = sigmoid(vis) * mask
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float r[8], sx[8];
        umma::ld8(tq + 8 * c, r);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const float2 bb = *reinterpret_cast<const float2*>(Bias + B16(M_VFC1) + 8 * c + i);
          const float2 xo = make_float2(XF[(8 * c + i) * ROWS + m], XF[(8 * c + i + 1) * ROWS + m]);
          const float2 xv = fadd2(xo, elu_pair(fadd2(make_float2(r[i], r[i + 1]), bb)));   // x = x + x_res
          XF[(8 * c + i) * ROWS + m] = xv.x;
          XF[(8 * c + i + 1) * ROWS + m] = xv.y;
          const float2 sv = fmul2(xv, make_float2(vis1, vis1));
          sx[i] = sv.x; sx[i + 1] = sv.y;
        }
        umma::store_chunk(E + E_HV, ROWS, c, m, sx);
      }
    }
    // ------------------------------------------------------------ vis_fc2(x * vis) 32 -> 32 -> 1 (output layer fused)
    STAGE_BEGIN() umma::gemm_issue(tb + 0, E + E_HV, ROWS, Wb + W16(M_V2_0), 32, 32, 32); STAGE_END()
    {
      float o1[1];
      epi_act_gemv<1>(tq + 0, Bias + B16(M_V2_0), nullptr, m, 4, ACT_ELU, Wsm + SMW(M_V2_1), 32, o1);
      const float vis2 = sigmoidf(o1[0] + Wsm[SMB(M_V2_1)]);
      SF[SF_VIS2 * ROWS + m] = vis2;
      // rgb_fc input [x(32), vis, ray_diff(4)] -> K = 48 in E chunks 4..9 (HV2 is dead)
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float xx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xx[i] = XF[(8 * c + i) * ROWS + m];
        umma::store_chunk(E + E_RG, ROWS, c, m, xx);
      }
      float tail[8] = {vis2, rg.dirdiff[0], rg.dirdiff[1], rg.dirdiff[2], rg.dirdiff[3], 0.f, 0.f, 0.f};
      umma::store_chunk(E + E_RG, ROWS, 4, m, tail);
      zero_chunk(E + E_RG, 5, m);
    }
    // ------------------------------------------------------------ rgb_fc.0, overlapped with view pooling #2
    umma::fence_smem_to_async(); umma::fence_before_sync(); wg_sync(wg);
    if (m == 0) { umma::fence_after_sync(); umma::gemm_issue(tb + 0, E + E_RG, ROWS, Wb + W16(M_RGB0), 16, 16, 48); umma::commit(bar); }
    if (vrow < V) {   // weights = vis / (sum + 1e-8), mean / var of x over views, mean of weights -> F2
      // the V threads of sample t take every V-th PAIR of channels (packed fp32 pairs); stores stay coalesced over t
      const long long gs = (long long)tile * T + t;
      float* f2 = a.f2 + (size_t)tile * kF2 * T + t;
      float sum = 0.f;
#pragma unroll
      for (int vv = 0; vv < V; ++vv) sum += SF[SF_VIS2 * ROWS + vv * T + t];
      float wv[V], ws = 0.f;
#pragma unroll
      for (int vv = 0; vv < V; ++vv) { wv[vv] = SF[SF_VIS2 * ROWS + vv * T + t] / (sum + 1e-8f); ws += wv[vv]; }
      if (gs < p.total) {
#pragma unroll 2
        for (int c = 2 * vrow; c < 32; c += 2 * V) {
          float2 x[V];
#pragma unroll
          for (int vv = 0; vv < V; ++vv) x[vv] = make_float2(XF[c * ROWS + vv * T + t], XF[(c + 1) * ROWS + vv * T + t]);
          float2 mean_c = fmul2(x[0], make_float2(wv[0], wv[0]));
#pragma unroll
          for (int vv = 1; vv < V; ++vv) mean_c = ffma2(x[vv], make_float2(wv[vv], wv[vv]), mean_c);
          const float2 nm = make_float2(-mean_c.x, -mean_c.y);
          float2 var_c = make_float2(0.f, 0.f);
#pragma unroll
          for (int vv = 0; vv < V; ++vv) { const float2 d = fadd2(x[vv], nm); var_c = ffma2(make_float2(wv[vv], wv[vv]), fmul2(d, d), var_c); }
          f2[(size_t)c * T] = mean_c.x; f2[(size_t)(c + 1) * T] = mean_c.y;
          f2[(size_t)(32 + c) * T] = var_c.x; f2[(size_t)(33 + c) * T] = var_c.y;
        }
        if (vrow == 0) f2[(size_t)64 * T] = ws / (float)V;
      }
    }
    mbar_wait(bar, phase); phase ^= 1; umma::fence_after_sync();
    {   // rgb_fc.2 (16 -> 8, ELU) and rgb_fc.4 (8 -> 1) fused as fp32 register GEMVs
      float h8[8];
      epi_act_gemv<8>(tq + 0, Bias + B16(M_RGB0), nullptr, m, 2, ACT_ELU, Wsm + SMW(M_RGB1), 16, h8);
      float o = Wsm[SMB(M_RGB2)];
#pragma unroll
      for (int i = 0; i < 8; ++i) o = fmaf(elu1(h8[i] + Wsm[SMB(M_RGB1) + i]), Wsm[SMW(M_RGB2) + i], o);
      SF[SF_LOGIT * ROWS + m] = o;
    }
    SF[SF_R * ROWS + m] = rgb_in[0]; SF[SF_G * ROWS + m] = rgb_in[1]; SF[SF_B * ROWS + m] = rgb_in[2];
    umma::fence_before_sync();
    wg_sync(wg);
    // ------------------------------------------------------------ softmax over views, blend the raw colours -> F2
    if (m < T) {
      const long long gs = (long long)tile * T + m;
      if (gs < p.total) {
        float mx = -INFINITY;
#pragma unroll
        for (int vv = 0; vv < V; ++vv) mx = fmaxf(mx, SF[SF_LOGIT * ROWS + vv * T + m]);
        float den = 0.f, r = 0.f, gg = 0.f, b = 0.f;
#pragma unroll
        for (int vv = 0; vv < V; ++vv) {
            const float e = fast_exp(SF[SF_LOGIT * ROWS + vv * T + m] - mx);
            den += e;
            r += SF[SF_R * ROWS + vv * T + m] * e; gg += SF[SF_G * ROWS + vv * T + m] * e; b += SF[SF_B * ROWS + vv * T + m] * e;
        }
        float* f2 = a.f2 + (size_t)tile * kF2 * T + m;
        f2[(size_t)(F2_RGB + 0) * T] = r / den; f2[(size_t)(F2_RGB + 1) * T] = gg / den; f2[(size_t)(F2_RGB + 2) * T] = b / den;
      }
    }
    wg_sync(wg);   // SF / E / P are rewritten by the next tile
  }

  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_w16_blob_bytes(void) { return kW16Bytes; }
extern "C" int pgrf_w16_num_layers(void) { return kNumLayers16; }
extern "C" int pgrf_w16_layer_info(int i, char* name, int name_cap, int* Kpad, int* Npad, int* w_offset_bytes, int* b_offset_bytes,
                                   int* kmap, int* nmap, int* is_small) {
  PGRF_REQUIRE(i >= 0 && i < kNumLayers16, "w16 layer index %d out of range", i);
  snprintf(name, name_cap, "%s", kLayers16[i].name);
  *Kpad = kLayers16[i].Kpad; *Npad = kLayers16[i].Npad;
  const int sec = kLayers16[i].section;
  *is_small = is_small16(i) ? 1 : 0;
  if (is_small16(i)) {   // fp32 W[N][K] row-major, then bias[N]
    const int small_begin = sec16_begin(sec) + sec16_w_bytes(sec) + 4 * sec16_b_floats(sec);
    *w_offset_bytes = small_begin + 4 * small16_offset(i);
    *b_offset_bytes = small_begin + 4 * small16_bias_offset(i);
  } else {
    *w_offset_bytes = sec16_begin(sec) + w16_offset(i);
    *b_offset_bytes = sec16_begin(sec) + sec16_w_bytes(sec) + 4 * b16_offset(i);
  }
  for (int k = 0; k < kLayers16[i].Kpad; ++k) kmap[k] = w16_kmap(i, k);
  for (int n = 0; n < kLayers16[i].Npad; ++n) nmap[n] = w16_nmap(i, n);
  return PGRF_OK;
}

namespace pgrf {
static_assert(SM16_BYTES + 1024 <= 227 * 1024, "fused MLP kernel shared memory");
int launch_render_mlp_bf16(const pgrf_render_args& a, int V, int T, long long total, int n_tiles, int sms, cudaStream_t st) {
  PGRF_REQUIRE(a.weights16 != nullptr, "render: bf16 path needs weights16");
  PGRF_REQUIRE(total < (1ll << 31), "render: rn * dn = %lld samples per launch exceed 2^31 (lower rays_per_launch)", total);
  PGRF_REQUIRE(((uintptr_t)a.weights16 & 15) == 0, "render: weights16 must be 16-byte aligned");
  Render16Params p;
  p.a = a; p.V = V; p.T = T; p.M = V * T; p.total = total; p.n_tiles = n_tiles;
  const int grid = min((n_tiles + kWG - 1) / kWG, sms);
#define PGRF_LAUNCH_V(VV)                                                                                              \
  case VV: {                                                                                                           \
    static bool done = false;                                                                                          \
    if (!done) {                                                                                                       \
      PGRF_CUDA(cudaFuncSetAttribute(render_mlp_bf16_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM16_BYTES)); \
      done = true;                                                                                                     \
    }                                                                                                                  \
    render_mlp_bf16_kernel<VV><<<grid, kThreads16, SM16_BYTES, st>>>(p);                                               \
  } break;
  switch (V) {
    PGRF_LAUNCH_V(1) PGRF_LAUNCH_V(2) PGRF_LAUNCH_V(3) PGRF_LAUNCH_V(4)
    default: PGRF_REQUIRE(false, "render: rfn=%d source views unsupported (1..4)", V);
  }
#undef PGRF_LAUNCH_V
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}
}  // namespace pgrf

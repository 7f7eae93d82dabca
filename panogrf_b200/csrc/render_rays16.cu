// bf16 tensor-core variant of the per-ray stage (geometry_fc, ray transformer, compositing, fine resampling).
//
// Same decomposition as render_bf16.cu: CTA = 4 independent warpgroups, a warpgroup owns one tile of up to 128
// samples (whole rays), thread <-> sample.  geometry_fc.0/.2 and the fused q|k|v projection are tcgen05 MMAs
// (bf16 operands, fp32 TMEM accumulators); the 4-head attention over the samples of a ray runs per thread on fp32
// K/V tiles kept in shared memory as float4 per (head, token); fc + LayerNorm + out_geometry_fc are 16-wide
// register GEMVs; compositing and the inverse-CDF fine resampling are one warp per ray with the stated sequential
// fp32 accumulation order.  Reference: ibrnet.py:352-364,15-27,72-102; render_ops.py:145-153,413-473;
// renderer.py:210-219,302-304,472.
#include <type_traits>

#include "render_device.cuh"
#include "render_layout16.cuh"
#include "render_rays_common.cuh"
#include "umma.cuh"

#define W16(L) (std::integral_constant<int, pgrf::w16_offset(L)>::value)
#define B16(L) (std::integral_constant<int, pgrf::b16_offset(L)>::value)
#define WOFF(L) (std::integral_constant<int, pgrf::sec_off(L) - pgrf::sec_off(pgrf::L_AFC)>::value)
#define BOFF(L) (std::integral_constant<int, pgrf::bias_off(L) - pgrf::sec_off(pgrf::L_AFC)>::value)

namespace pgrf {

constexpr int kWGr = 4;
constexpr int kThreadsR = 128 * kWGr;
constexpr int RROWS = 128;
constexpr int RCH = RROWS * 16;

// per-warpgroup shared memory (bytes)
constexpr int R_A = 0;                       // 10 chunks: [mean 32 | var 32 | wmean, 0 x15]; later H64 (8 chunks)
constexpr int R_G16 = 10 * RCH;              // 2 chunks
constexpr int R_K4 = R_G16 + 2 * RCH;        // float4 [4 heads][128 tokens]
constexpr int R_V4 = R_K4 + 4 * RROWS * 16;
constexpr int R_RV = R_V4 + 4 * RROWS * 16;  // 6 x [256] floats
constexpr int R_RGB = R_RV + 6 * 256 * 4;    // 3 x [128] floats
constexpr int R_WG_BYTES = R_RGB + 3 * RROWS * 4;

constexpr int kSec1Bytes = sec16_bytes(1);
constexpr int kR3W32Begin = sec_off(L_AFC);                            // only fc, out_geometry_fc and layer norm stay fp32
constexpr int kR3W32 = section_floats(2) + 32 - kR3W32Begin;
constexpr int SMR_W16 = 0;
constexpr int SMR_W32 = (kSec1Bytes + 15) & ~15;
constexpr int SMR_PE = SMR_W32 + ((kR3W32 * 4 + 15) & ~15);
constexpr int SMR_WG = (SMR_PE + kMaxSamplesPerRay * 16 * 4 + 127) & ~127;
constexpr int SMR_BAR = SMR_WG + kWGr * R_WG_BYTES;
constexpr int SMR_BYTES = SMR_BAR + 128;

__device__ __forceinline__ void wgr_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }

static __device__ __noinline__ void epi_elu_store(uint32_t taddr, const float* __restrict__ bias, unsigned char* dst, int m, int nchunks) {
#pragma unroll 1
  for (int c = 0; c < nchunks; c += 2) {
    float v[16];
    umma::ld16(taddr + 8 * c, v);
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) q[i >> 1] = elu_pack(fadd2(make_float2(v[i], v[i + 1]), *reinterpret_cast<const float2*>(bias + 8 * c + i)));
    *reinterpret_cast<uint4*>(dst + ((size_t)c * RROWS + m) * 16) = make_uint4(q[0], q[1], q[2], q[3]);
    *reinterpret_cast<uint4*>(dst + ((size_t)(c + 1) * RROWS + m) * 16) = make_uint4(q[4], q[5], q[6], q[7]);
  }
}

struct Rays16Params {
  pgrf_render_args a;
  int V, T, log2T;
  long long total;
  int rays_per_tile, n_tiles;
};

#define RSTAGE_BEGIN()                                                      \
  umma::fence_smem_to_async(); umma::fence_before_sync(); wgr_sync(wg);     \
  if (m == 0) { umma::fence_after_sync();
#define RSTAGE_END()                                                        \
    umma::commit(bar); }                                                    \
  mbar_wait(bar, phase); phase ^= 1; umma::fence_after_sync();

// two-pass softmax(q.k) V with the exact running maximum (fallback of the bounded single pass below)
static __device__ __noinline__ void attention_exact(float q0, float q1, float q2, float q3, const float4* Kh, const float4* Vh, int dn,
                                                    float& den, float& o0, float& o1, float& o2, float& o3) {
  float mx = -INFINITY;
  for (int j = 0; j < dn; ++j) {
    const float4 k = Kh[j];
    mx = fmaxf(mx, q0 * k.x + q1 * k.y + q2 * k.z + q3 * k.w);
  }
  den = 0.f; o0 = 0.f; o1 = 0.f; o2 = 0.f; o3 = 0.f;
  for (int j = 0; j < dn; ++j) {
    const float4 k = Kh[j];
    const float4 vv = Vh[j];
    const float e = fast_exp(q0 * k.x + q1 * k.y + q2 * k.z + q3 * k.w - mx);
    den += e;
    o0 = fmaf(e, vv.x, o0); o1 = fmaf(e, vv.y, o1); o2 = fmaf(e, vv.z, o2); o3 = fmaf(e, vv.w, o3);
  }
}

__global__ void __launch_bounds__(kThreadsR, 1) render_rays_bf16_kernel(const Rays16Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x;
  const int wg = tid >> 7, m = tid & 127, wq = (tid >> 5) & 3, lane = tid & 31;
  unsigned char* Wb = smem + SMR_W16;
  const float* Bias = reinterpret_cast<const float*>(Wb + sec16_w_bytes(1));
  float* W32 = reinterpret_cast<float*>(smem + SMR_W32);
  float* PE = reinterpret_cast<float*>(smem + SMR_PE);
  unsigned char* G = smem + SMR_WG + wg * R_WG_BYTES;
  float4* K4 = reinterpret_cast<float4*>(G + R_K4);
  float4* V4 = reinterpret_cast<float4*>(G + R_V4);
  float* RV = reinterpret_cast<float*>(G + R_RV);
  float* RGB = reinterpret_cast<float*>(G + R_RGB);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SMR_BAR) + wg;
  uint64_t* tbar = reinterpret_cast<uint64_t*>(smem + SMR_BAR) + kWGr + wg;     // completion of the tile's bulk copy
  const unsigned char* f2_op = reinterpret_cast<const unsigned char*>(a.f2);    // bf16 operand tiles written by the MLP kernel

  {
    constexpr int sec1 = sec16_begin(1);
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.weights16) + sec1);
    uint4* dst = reinterpret_cast<uint4*>(Wb);
    for (int i = tid; i < kSec1Bytes / 16; i += kThreadsR) dst[i] = __ldg(src + i);
    constexpr int sec2 = section_begin(2);
    for (int i = tid; i < kR3W32; i += kThreadsR) W32[i] = __ldg(a.weights + sec2 + kR3W32Begin + i);
    for (int i = tid; i < a.dn * 16; i += kThreadsR) PE[i] = __ldg(a.weights + kPosencOffset + i);
  }
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 256);
  if (m == 0) { mbar_init(bar, 1); mbar_init(tbar, 1); }
  // chunk 9 of the A operand (K columns 72..79 of geometry_fc.0) is never written by a tile: zero it once
  *reinterpret_cast<uint4*>(G + R_A + ((size_t)9 * RROWS + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s + wg * 64;
  const uint32_t tq = tb + ((uint32_t)(wq * 32) << 16);
  uint32_t phase = 0, tphase = 0;
  constexpr int ln_rel = section_floats(2) - kR3W32Begin;
  const float* LNW = W32 + ln_rel;
  const float* LNB = LNW + 16;

  const int dn = a.dn, V = p.V, T = p.T;
  const int rpt = p.rays_per_tile;
  const int Mv = rpt * dn;
  const float4* f2_rgb = reinterpret_cast<const float4*>(f2_op + (size_t)p.n_tiles * kF2TileBytes);

  __shared__ int s_tile[kWGr];
  __shared__ float4 s_kn[kWGr][4];
  int static_tile = blockIdx.x * kWGr + wg;
#pragma unroll 1
  while (true) {
    int tile;
    if (a.sched) {   // dynamic tile scheduler
      if (m == 0) s_tile[wg] = atomicAdd(a.sched + 1, 1);
      wgr_sync(wg);
      tile = s_tile[wg];
    } else {
      tile = static_tile;
      static_tile += gridDim.x * kWGr;
    }
    if (tile >= p.n_tiles) break;
    const long long g0 = (long long)tile * Mv;
    long long g = g0 + min(m, Mv - 1);
    if (g >= p.total) g = p.total - 1;
    const int sidx = (int)((unsigned)g % (unsigned)dn);    // sample index inside its ray (total < 2^31, checked by the MLP launcher)
    // ---- pooled features of the tile: the MLP kernel wrote them as a ready bf16 A operand (9 k-chunks x 128 rows); ONE bulk copy
    //      (TMA) brings the tile in while the threads fetch their sample's blended colour
    if (m == 0) {
      mbar_expect_tx(tbar, kF2TileBytes);
      bulk_g2s(G + R_A, f2_op + (size_t)tile * kF2TileBytes, kF2TileBytes, tbar);
    }
    {
      const float4 c = __ldcs(f2_rgb + g);
      RGB[m] = c.x; RGB[RROWS + m] = c.y; RGB[2 * RROWS + m] = c.z;
    }
    mbar_wait(tbar, tphase); tphase ^= 1;
    // ---- geometry_fc 65 -> 64 -> 16 (+ positional code)
    RSTAGE_BEGIN() umma::gemm_issue(tb, G + R_A, RROWS, Wb + W16(M_GEO0), 64, 64, 80); RSTAGE_END()
    epi_elu_store(tq, Bias + B16(M_GEO0), G + R_A, m, 8);       // H64 overwrites A: its MMA is complete
    RSTAGE_BEGIN() umma::gemm_issue(tb, G + R_A, RROWS, Wb + W16(M_GEO1), 16, 16, 64); RSTAGE_END()
    float g16[16];
    {
      umma::ld16(tq, g16);
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 r = fadd2(elu_pair(fadd2(make_float2(g16[i], g16[i + 1]), *reinterpret_cast<const float2*>(Bias + B16(M_GEO1) + i))),
                               *reinterpret_cast<const float2*>(PE + sidx * 16 + i));
        g16[i] = r.x; g16[i + 1] = r.y;
      }
      umma::store_chunk(G + R_G16, RROWS, 0, m, g16);
      umma::store_chunk(G + R_G16, RROWS, 1, m, g16 + 8);
    }
    // ---- q | k | v projections
    RSTAGE_BEGIN() umma::gemm_issue(tb, G + R_G16, RROWS, Wb + W16(M_QKV), 48, 48, 16); RSTAGE_END()
    float q[16];
    {
      umma::ld16(tq, q);
#pragma unroll
      for (int i = 0; i < 16; ++i) q[i] *= 0.5f;            // q / temperature, temperature = sqrt(d_k) = 2
      float kv[16];
      umma::ld16(tq + 16, kv);
      float kn2[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        K4[h * RROWS + m] = make_float4(kv[4 * h], kv[4 * h + 1], kv[4 * h + 2], kv[4 * h + 3]);
        kn2[h] = kv[4 * h] * kv[4 * h] + kv[4 * h + 1] * kv[4 * h + 1] + kv[4 * h + 2] * kv[4 * h + 2] + kv[4 * h + 3] * kv[4 * h + 3];
      }
      // largest key norm of the tile, per head: |q||k|max bounds every score, so the softmax needs no separate max pass
      if (m >= Mv || g0 + m >= p.total) { kn2[0] = 0.f; kn2[1] = 0.f; kn2[2] = 0.f; kn2[3] = 0.f; }   // rows without a sample
#pragma unroll
      for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kn2[h] = fmaxf(kn2[h], __shfl_xor_sync(0xffffffffu, kn2[h], o));
      }
      if (lane == 0) s_kn[wg][wq] = make_float4(kn2[0], kn2[1], kn2[2], kn2[3]);
      umma::ld16(tq + 32, kv);
#pragma unroll
      for (int h = 0; h < 4; ++h) V4[h * RROWS + m] = make_float4(kv[4 * h], kv[4 * h + 1], kv[4 * h + 2], kv[4 * h + 3]);
    }
    umma::fence_before_sync();
    wgr_sync(wg);
    // ---- 4-head attention over the samples of this token's ray (mask rule ibrnet.py:359-360: < 2 views -> uniform)
    float ao[16];
    {
      const int r0 = (min(m, Mv - 1) / dn) * dn;
      const bool masked = !(V > 1);
      float kmax[4];
      {
        const float4 a0 = s_kn[wg][0], a1 = s_kn[wg][1], a2 = s_kn[wg][2], a3 = s_kn[wg][3];
        kmax[0] = sqrtf(fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)));
        kmax[1] = sqrtf(fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y)));
        kmax[2] = sqrtf(fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z)));
        kmax[3] = sqrtf(fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w)));
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float q0 = q[4 * h], q1 = q[4 * h + 1], q2 = q[4 * h + 2], q3 = q[4 * h + 3];
        const float4* Kh = K4 + h * RROWS + r0;
        const float4* Vh = V4 + h * RROWS + r0;
        float den = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
        if (masked) {                                   // all scores equal -> uniform weights
#pragma unroll 4
          for (int j = 0; j < dn; ++j) { const float4 vv = Vh[j]; o0 += vv.x; o1 += vv.y; o2 += vv.z; o3 += vv.w; }
          den = (float)dn;
        } else {
          // softmax(s) == softmax(s - c) for any c; c = |q||k|max >= max_j s_j (Cauchy-Schwarz) keeps exp() <= 1
          const float bound = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3) * kmax[h];
          const float2 qa = make_float2(q0, q1), qb = make_float2(q2, q3), c0 = make_float2(-bound, 0.f);
          float2 oa = make_float2(0.f, 0.f), ob = make_float2(0.f, 0.f);
#pragma unroll 4
          for (int j = 0; j < dn; ++j) {
            const float4 k = Kh[j];
            const float4 vv = Vh[j];
            float2 t = ffma2(qa, make_float2(k.x, k.y), c0);
            t = ffma2(qb, make_float2(k.z, k.w), t);
            const float e = fast_exp(t.x + t.y);
            den += e;
            const float2 ee = make_float2(e, e);
            oa = ffma2(ee, make_float2(vv.x, vv.y), oa);
            ob = ffma2(ee, make_float2(vv.z, vv.w), ob);
          }
          o0 = oa.x; o1 = oa.y; o2 = ob.x; o3 = ob.y;
          // the bound was more than ~46 above the true maximum (huge, misaligned q/k): small terms were flushed -> exact path
          if (!(den >= 1e-20f)) attention_exact(q0, q1, q2, q3, Kh, Vh, dn, den, o0, o1, o2, o3);
        }
        const float inv = 1.f / den;
        ao[4 * h] = o0 * inv; ao[4 * h + 1] = o1 * inv; ao[4 * h + 2] = o2 * inv; ao[4 * h + 3] = o3 * inv;
      }
    }
    // ---- fc + residual + LayerNorm(1e-6) + out_geometry_fc 16 -> 16 -> 1 (ReLU), fp32 register GEMVs
    {
      float o[16], x[16];
      reg_layer<16, 16, 16>(ao, W32 + WOFF(L_AFC), nullptr, o);
      float mean = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { x[i] = o[i] + g16[i]; mean += x[i]; }
      mean *= (1.f / 16.f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) var += (x[i] - mean) * (x[i] - mean);
      var *= (1.f / 16.f);
      const float rstd = rsqrtf(var + 1e-6f);
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = (x[i] - mean) * rstd * LNW[i] + LNB[i];
      float h1[16], s4[4];
      reg_layer<16, 16, 16>(x, W32 + WOFF(L_OG0), W32 + BOFF(L_OG0), h1);
#pragma unroll
      for (int i = 0; i < 16; i += 2) { const float2 r = elu_pair(make_float2(h1[i], h1[i + 1])); h1[i] = r.x; h1[i + 1] = r.y; }
      reg_layer<16, 1, 4>(h1, W32 + WOFF(L_OG1), W32 + BOFF(L_OG1), s4);
      RV[RV2_SIGMA * 256 + m] = fmaxf(s4[0], 0.f);
    }
    wgr_sync(wg);
    // ---- per ray: alpha compositing (+ fine resampling), one warp per ray
    for (int r = wq; r < rpt; r += 4) {
      const long long ray = (long long)tile * rpt + r;
      if (ray >= a.rn) continue;
      composite_ray(a, ray, r, dn, lane, RV, RGB, RROWS);
    }
    umma::fence_smem_to_async();
    wgr_sync(wg);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base_s, 256);
}

static_assert(SMR_BYTES <= 227 * 1024, "rays kernel shared memory");

int launch_render_rays_bf16(const pgrf_render_args& a, int V, int T, long long total, int sms, cudaStream_t st) {
  Rays16Params p;
  p.a = a; p.V = V; p.T = T; p.total = total;
  p.log2T = 0;
  while ((1 << p.log2T) < T) ++p.log2T;
  p.rays_per_tile = a.dn >= RROWS ? 1 : RROWS / a.dn;
  p.n_tiles = (a.rn + p.rays_per_tile - 1) / p.rays_per_tile;
  static bool attr_done[64] = {};   // per-device function attribute
  int dev = 0;
  PGRF_CUDA(cudaGetDevice(&dev));
  if (!attr_done[dev & 63]) {
    PGRF_CUDA(cudaFuncSetAttribute(render_rays_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMR_BYTES));
    attr_done[dev & 63] = true;
  }
  const int grid = min((p.n_tiles + kWGr - 1) / kWGr, sms);
  render_rays_bf16_kernel<<<grid, kThreadsR, SMR_BYTES, st>>>(p);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

}  // namespace pgrf

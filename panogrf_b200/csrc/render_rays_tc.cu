// Per-ray stage of the bf16 path with the ray transformer's attention ON THE TENSOR CORES (dn = 64 or 128 samples per ray).
//
// geometry_fc 65 -> 64 -> 16 (+ positional code), 4-head self-attention over the samples of a ray (ibrnet.py:15-27,52-102),
// fc + residual + LayerNorm, out_geometry_fc 16 -> 16 -> 1, alpha compositing and the inverse-CDF fine resampling.
// Same warpgroup / thread <-> sample decomposition as render_rays16.cu (which stays the path for every other dn), but
//   * the tile's pooled features arrive as a ready bf16 A operand through ONE bulk copy, prefetched one tile ahead (the copy of
//     tile i+1 is issued as soon as geometry_fc.0 of tile i has consumed the buffer; no generic-proxy access ever touches it);
//   * hidden activations (H64, g16) go back to tensor memory as the next MMA's A operand;
//   * attention = per head one  S = Q K^T  MMA and one  O += P V  MMA chain.  A tile holds RPT = 128/dn whole rays, and a row
//     must only see the keys of its own ray: for two rays the contraction is made block-diagonal — A row = [q | 0] (ray 0) or
//     [0 | q] (ray 1) against B row j = [k_ray0[j] | k_ray1[j]] (K = 8 + 8) — so ONE M = 128, N = 64, K = 16 MMA gives every row
//     its own ray's 64 scores; for one ray of 128 samples the keys are processed in two blocks of 64 that share the softmax
//     shift.  exp() runs on the accumulator in registers (thread = row), P is written back IN PLACE as bf16 (tcgen05.st) and
//     multiplied with V^T (head-masked copies, so all heads accumulate into one 16-column output per ray);
//   * softmax(s) = softmax(s - c): c = |q| max|k| (Cauchy-Schwarz, no max pass) while that bound is below 2^-60 safe; a tile in
//     which any row exceeds it takes an exact pre-pass (scores recomputed, true row maxima).
// Numerics: q, k, v, P rounded to bf16 (fp32 accumulation); everything else fp32.
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

constexpr int kWGt = 4;
constexpr int kThreadsT = 128 * kWGt;
constexpr int TROWS = 128;
constexpr int TCH = TROWS * 16;

// per-warpgroup shared memory (bytes)
constexpr int T_A = 0;                        // 10 chunks: F2 operand tile (9 chunks, bulk copy) + one zero chunk
constexpr int T_KB = 10 * TCH;                // [4 heads][2][64 rows][16 B]   keys:  B operand of S = Q K^T
constexpr int T_VB = T_KB + 4 * 2 * 64 * 16;  // [4 heads][2][8 chunks][16 rows][16 B]   values^T (head-masked): B operand of O += P V
constexpr int T_RV = T_VB + 4 * 2 * 8 * 16 * 16;
constexpr int T_RGB = T_RV + 6 * RRV * 4;
constexpr int T_WG_BYTES = T_RGB + 3 * TROWS * 4;

constexpr int kSec1BytesT = sec16_bytes(1);
constexpr int kT3W32Begin = sec_off(L_AFC);                            // fc, out_geometry_fc and layer norm stay fp32
constexpr int kT3W32 = section_floats(2) + 32 - kT3W32Begin;
constexpr int SMT_W16 = 0;
constexpr int SMT_W32 = (kSec1BytesT + 15) & ~15;
constexpr int SMT_WG = (SMT_W32 + kT3W32 * 4 + 127) & ~127;
constexpr int SMT_ZERO = SMT_WG + kWGt * T_WG_BYTES;                   // zero chunk above every operand
constexpr int SMT_BAR = SMT_ZERO + 64 * 16;
constexpr int SMT_BYTES = SMT_BAR + 128;
static_assert(SMT_BYTES + 1024 <= 227 * 1024, "tensor-core rays kernel shared memory");

// tensor-memory columns of a warpgroup (128)
constexpr int TC_R = 0;          // 64: geometry_fc.0 accumulator | [geometry_fc.2 16 | qkv 48] | scores S -> probabilities P (32)
constexpr int TC_QA = 64;        // 32: H64 operand of geometry_fc.2, then the four heads' query operands (8 columns each)
constexpr int TC_G16 = 96;       // 8: operand of the q|k|v projection
constexpr int TC_O = 96;         // 2 x 16: attention output per ray (all heads), after the projection has consumed TC_G16

constexpr float kLog2eT = 1.4426950408889634f;
constexpr float kSafeBound = 60.f;   // |q||k|max (log2 units) up to which 2^(s - bound) cannot underflow for the row's largest score

__device__ __forceinline__ void wgt_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
__device__ __forceinline__ bool elect_one_t() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void st16t(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

struct RaysTcParams {
  pgrf_render_args a;
  int V;
  long long total;
  int n_tiles;
};

// RPT = rays per tile: 2 (dn = 64) or 1 (dn = 128)
template <int RPT>
__global__ void __launch_bounds__(kThreadsT, 1) render_rays_tc_kernel(const RaysTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_tile[kWGt];
  __shared__ float4 s_kn[kWGt][4];
  __shared__ int s_exact[kWGt];
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x;
  const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);
  const int m = tid & 127, lane = tid & 31;
  const int wq = __shfl_sync(0xffffffffu, (tid >> 5) & 3, 0);
  constexpr int dn = 128 / RPT;
  constexpr int NBLK = 2 / RPT;                 // key blocks of 64 per ray
  constexpr int NR = 4 * NBLK;                  // attention rounds (head, key block)
  unsigned char* Wb = smem + SMT_W16;
  const float* Bias = reinterpret_cast<const float*>(Wb + sec16_w_bytes(1));
  float* W32 = reinterpret_cast<float*>(smem + SMT_W32);
  unsigned char* G = smem + SMT_WG + wg * T_WG_BYTES;
  float* RV = reinterpret_cast<float*>(G + T_RV);
  float* RGB = reinterpret_cast<float*>(G + T_RGB);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SMT_BAR) + wg;
  uint64_t* tbar = reinterpret_cast<uint64_t*>(smem + SMT_BAR) + kWGt + wg;     // completion of the tile's bulk copy
  const unsigned char* f2_op = reinterpret_cast<const unsigned char*>(a.f2);
  const float4* f2_rgb = reinterpret_cast<const float4*>(f2_op + (size_t)p.n_tiles * kF2TileBytes);

  {
    constexpr int sec1 = sec16_begin(1);
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.weights16) + sec1);
    uint4* dst = reinterpret_cast<uint4*>(Wb);
    for (int i = tid; i < kSec1BytesT / 16; i += kThreadsT) dst[i] = __ldg(src + i);
    constexpr int sec2 = section_begin(2);
    for (int i = tid; i < kT3W32; i += kThreadsT) W32[i] = __ldg(a.weights + sec2 + kT3W32Begin + i);
    if (tid < 64) reinterpret_cast<uint4*>(smem + SMT_ZERO)[tid] = make_uint4(0u, 0u, 0u, 0u);
    // zero once: chunk 9 of the A operand, the value operands (rows of the other heads stay zero for ever)
    *reinterpret_cast<uint4*>(G + T_A + ((size_t)9 * TROWS + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = m; i < 4 * 2 * 8 * 16; i += 128) reinterpret_cast<uint4*>(G + T_VB)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 512);
  if (m == 0) { mbar_init(bar, 1); mbar_init(tbar, 1); }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0) + wg * 128;
  const uint32_t tq = tb + ((uint32_t)(wq * 32) << 16);
  const uint32_t bar_addr = umma::smem_addr(bar);
  uint32_t phase = 0, tphase = 0;
  constexpr int ln_rel = section_floats(2) - kT3W32Begin;
  const float* LNW = W32 + ln_rel;
  const float* LNB = LNW + 16;
  const uint32_t sW = umma::smem_addr(Wb), sZ = umma::smem_addr(smem + SMT_ZERO), sA = umma::smem_addr(G + T_A);
  const uint32_t sKB = umma::smem_addr(G + T_KB), sVB = umma::smem_addr(G + T_VB);
  const bool masked = !(p.V > 1);               // ibrnet.py:359-360: < 2 views -> every key masked -> uniform attention
  const int rr = (RPT == 2) ? (wq >> 1) : 0;    // ray of this row inside the tile (warp-uniform)
  const int jk = m & 63;                        // key index inside its ray (RPT = 2) / key block (RPT = 1)
  const int kc = (RPT == 2) ? rr : (m >> 6);    // operand slot of this row's key / value: ray (RPT = 2) or key block (RPT = 1)

#define T_ISSUE_BEGIN() if (wq == 0) { if (elect_one_t()) { umma::fence_after_sync();
#define T_ISSUE_END() umma::commit(bar); } __syncwarp(); }
#define T_WAIT() { if (wq == 0) mbar_wait(bar, phase); wgt_sync(wg); phase ^= 1; umma::fence_after_sync(); }
#define T_SYNC_TMEM() { umma::wait_st(); umma::fence_before_sync(); wgt_sync(wg); }
#define T_SYNC_BOTH() { umma::wait_st(); umma::fence_smem_to_async(); umma::fence_before_sync(); wgt_sync(wg); }
  auto mma_ss = [&](uint32_t d, uint32_t aaddr, uint32_t albo, uint32_t baddr, uint32_t blbo, int N, uint32_t acc) {
    umma::mma_bf16(d, umma::smem_desc(aaddr, albo, 128), umma::smem_desc(baddr, blbo, 128), umma::instr_desc_bf16(128, N), acc);
  };
  auto mma_ts = [&](uint32_t d, uint32_t atmem, uint32_t baddr, uint32_t blbo, int N, uint32_t acc) {
    umma::mma_bf16_ts(d, atmem, umma::smem_desc(baddr, blbo, 128), umma::instr_desc_bf16(128, N), acc);
  };
  // scores of round (head h, key block c): S[128 x 64] = QA_h . KB[h]^T
  auto issue_scores = [&](int h, int c) {
    const uint32_t kb = sKB + (uint32_t)(h * 2) * 1024;
    if (RPT == 2) mma_ts(tb + TC_R, tb + TC_QA + 8 * h, kb, 1024, 64, 0u);                       // K = [ray0 8 | ray1 8]
    else mma_ts(tb + TC_R, tb + TC_QA + 8 * h, kb + c * 1024, sZ - (kb + c * 1024), 64, 0u);     // K = [block 8 | zero]
  };
  // O_ray += P . V^T of round (h, c)
  auto issue_pv = [&](int h, int c, uint32_t acc_first) {
    const uint32_t vb = sVB + (uint32_t)(h * 2) * 2048;
    if (RPT == 2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ts(tb + TC_O + 16 * r, tb + TC_R + 8 * k, vb + r * 2048 + k * 512, 256, 16, (k > 0) ? 1u : acc_first);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ts(tb + TC_O, tb + TC_R + 8 * k, vb + c * 2048 + k * 512, 256, 16, (k > 0) ? 1u : acc_first);
    }
  };

  int tile, next_tile;
  if (a.sched) {
    if (m == 0) s_tile[wg] = atomicAdd(a.sched + 1, 1);
    wgt_sync(wg);
    tile = s_tile[wg];
  } else {
    tile = blockIdx.x * kWGt + wg;
  }
  if (tile < p.n_tiles && m == 0) {       // first tile's operand
    mbar_expect_tx(tbar, kF2TileBytes);
    bulk_g2s(G + T_A, f2_op + (size_t)tile * kF2TileBytes, kF2TileBytes, tbar);
  }
#pragma unroll 1
  while (tile < p.n_tiles) {
    // next tile index one iteration ahead (its operand is prefetched below)
    if (a.sched) {
      if (m == 0) s_tile[wg] = atomicAdd(a.sched + 1, 1);
    }
    const long long g0 = (long long)tile * 128;
    long long g = g0 + m;
    const bool row_valid = g < p.total;
    if (!row_valid) g = p.total - 1;
    const int sidx = m & (dn - 1);                // sample index inside its ray
    {
      const float4 c = __ldcs(f2_rgb + g);
      RGB[m] = c.x; RGB[TROWS + m] = c.y; RGB[2 * TROWS + m] = c.z;
    }
    if (wq == 0) mbar_wait(tbar, tphase);
    tphase ^= 1;
    wgt_sync(wg);
    next_tile = a.sched ? s_tile[wg] : tile + gridDim.x * kWGt;
    // ---- geometry_fc.0: 65 -> 64
    T_ISSUE_BEGIN()
#pragma unroll
      for (int k = 0; k < 5; ++k) mma_ss(tb + TC_R, sA + 2 * k * TCH, TCH, sW + W16(M_GEO0) + 2 * k * 64 * 16, 64 * 16, 64, k > 0);
    T_ISSUE_END()
    T_WAIT()
    if (next_tile < p.n_tiles && m == 0) {   // the operand buffer is free (only the MMA read it): prefetch the next tile
      mbar_expect_tx(tbar, kF2TileBytes);
      bulk_g2s(G + T_A, f2_op + (size_t)next_tile * kF2TileBytes, kF2TileBytes, tbar);
    }
    {   // H64 = ELU(. + bias) -> tensor memory (operand of geometry_fc.2)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        umma::ld32(tq + TC_R + 32 * half, v);
        uint32_t h[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          h[i] = elu_pack(fadd2(make_float2(v[2 * i], v[2 * i + 1]), *reinterpret_cast<const float2*>(Bias + B16(M_GEO0) + 32 * half + 2 * i)));
        st16t(tq + TC_QA + 16 * half, h);
      }
    }
    T_SYNC_TMEM()
    T_ISSUE_BEGIN()
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ts(tb + TC_R, tb + TC_QA + 8 * k, sW + W16(M_GEO1) + 2 * k * 16 * 16, 16 * 16, 16, k > 0);
    T_ISSUE_END()
    T_WAIT()
    float g16[16];
    {
      umma::ld16(tq + TC_R, g16);
      const float* pe = a.weights + kPosencOffset + sidx * 16;
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        const float2 r = fadd2(elu_pair(fadd2(make_float2(g16[i], g16[i + 1]), *reinterpret_cast<const float2*>(Bias + B16(M_GEO1) + i))),
                               __ldg(reinterpret_cast<const float2*>(pe + i)));
        g16[i] = r.x; g16[i + 1] = r.y;
        h[i >> 1] = umma::pack2(r.x, r.y);
      }
      umma::st8(tq + TC_G16, h);
    }
    // ---- q | k | v projections (no bias)
    T_SYNC_TMEM()
    T_ISSUE_BEGIN()
      mma_ts(tb + TC_R + 16, tb + TC_G16, sW + W16(M_QKV), 48 * 16, 48, 0u);
    T_ISSUE_END()
    T_WAIT()
    float cb[4];          // softmax shift per head (log2 units)
    uint32_t qpk[8];      // the four heads' bf16 queries
    {
      float q[16], kv[16];
      umma::ld16(tq + TC_R + 16, q);
      umma::ld16(tq + TC_R + 32, kv);
      float kn2[4], qn2[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        // scores in log2 units: q / temperature (sqrt(d_k) = 2) * log2(e); operands are bf16: norms of the ROUNDED values
        const uint32_t q01 = umma::pack2(q[4 * h] * (0.5f * kLog2eT), q[4 * h + 1] * (0.5f * kLog2eT));
        const uint32_t q23 = umma::pack2(q[4 * h + 2] * (0.5f * kLog2eT), q[4 * h + 3] * (0.5f * kLog2eT));
        // rows without a sample hold whatever the workspace contained: their keys / values must be exact zeros, because the
        // block-diagonal contraction multiplies them with the zero half of the other ray's queries (0 x NaN = NaN)
        const uint32_t k01 = row_valid ? umma::pack2(kv[4 * h], kv[4 * h + 1]) : 0u, k23 = row_valid ? umma::pack2(kv[4 * h + 2], kv[4 * h + 3]) : 0u;
        // key row [k(4), 1, 0, 0, 0]: the 1 meets the query's 5th slot, which carries -shift (written once the shift is known)
        *reinterpret_cast<uint4*>(G + T_KB + ((size_t)(h * 2 + kc) * 64 + jk) * 16) = make_uint4(k01, k23, row_valid ? 0x00003F80u : 0u, 0u);
        qpk[2 * h] = q01; qpk[2 * h + 1] = q23;
        const float2 qa2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q01)), qb2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q23));
        const float2 ka2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&k01)), kb2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&k23));
        qn2[h] = qa2.x * qa2.x + qa2.y * qa2.y + qb2.x * qb2.x + qb2.y * qb2.y;
        kn2[h] = row_valid ? ka2.x * ka2.x + ka2.y * ka2.y + kb2.x * kb2.x + kb2.y * kb2.y : 0.f;
      }
      // largest key norm of the tile, per head: |q||k|max bounds every score, so the softmax needs no separate max pass
#pragma unroll
      for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) kn2[h] = fmaxf(kn2[h], __shfl_xor_sync(0xffffffffu, kn2[h], o));
      }
      if (lane == 0) s_kn[wg][wq] = make_float4(kn2[0], kn2[1], kn2[2], kn2[3]);
      umma::ld16(tq + TC_R + 48, kv);         // values -> head-masked V^T operands: element (n = 4h + d, key jk) of slot kc
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        unsigned char* vb = G + T_VB + (size_t)(h * 2 + kc) * 2048 + (size_t)(jk >> 3) * 256 + (jk & 7) * 2;
#pragma unroll
        for (int d = 0; d < 4; ++d)
          *reinterpret_cast<__nv_bfloat16*>(vb + (4 * h + d) * 16) = __float2bfloat16_rn(row_valid ? kv[4 * h + d] : 0.f);
      }
      if (m == 0) s_exact[wg] = 0;
      T_SYNC_BOTH()
      {
        const float4 a0 = s_kn[wg][0], a1 = s_kn[wg][1], a2 = s_kn[wg][2], a3 = s_kn[wg][3];
        cb[0] = sqrtf(qn2[0] * fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)));
        cb[1] = sqrtf(qn2[1] * fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y)));
        cb[2] = sqrtf(qn2[2] * fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z)));
        cb[3] = sqrtf(qn2[3] * fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w)));
      }
      if (!masked && fmaxf(fmaxf(cb[0], cb[1]), fmaxf(cb[2], cb[3])) > kSafeBound) s_exact[wg] = 1;
    }
    // query operands: [q(4), -shift, 0, 0, 0] in the K-half of the row's ray -> the MMA delivers  s - shift  directly
    auto write_queries = [&]() {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t qa[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        const uint32_t sh = umma::pack2(-cb[h], 0.f);
        if (RPT == 2 && rr == 1) { qa[4] = qpk[2 * h]; qa[5] = qpk[2 * h + 1]; qa[6] = sh; }
        else { qa[0] = qpk[2 * h]; qa[1] = qpk[2 * h + 1]; qa[2] = sh; }
        umma::st8(tq + TC_QA + 8 * h, qa);
      }
    };
    wgt_sync(wg);
    const bool exact = !masked && s_exact[wg] != 0;     // tile-uniform
    if (exact) {
      // rare: huge logits. Pre-pass with the true row maxima (scores are recomputed in the main rounds)
      cb[0] = 0.f; cb[1] = 0.f; cb[2] = 0.f; cb[3] = 0.f;
      write_queries();
      T_SYNC_TMEM()
      float mx0 = 0.f, mx1 = 0.f, mx2 = 0.f, mx3 = 0.f;
#pragma unroll 1
      for (int h = 0; h < 4; ++h) {
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < NBLK; ++c) {
          T_ISSUE_BEGIN() issue_scores(h, c); T_ISSUE_END()
          T_WAIT()
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float y[32];
            umma::ld32(tq + TC_R + 32 * half, y);
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, y[i]);
          }
          umma::fence_before_sync();
          wgt_sync(wg);
        }
        if (h == 0) mx0 = mx; else if (h == 1) mx1 = mx; else if (h == 2) mx2 = mx; else mx3 = mx;
      }
      cb[0] = mx0; cb[1] = mx1; cb[2] = mx2; cb[3] = mx3;
    }
    write_queries();
    T_SYNC_TMEM()
    // ---- attention rounds: S -> P = 2^(S - c) (in place, bf16) -> O += P V
    float den[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const int h = r / NBLK, c = r % NBLK;
      T_ISSUE_BEGIN()
        if (r > 0) issue_pv((r - 1) / NBLK, (r - 1) % NBLK, (r - 1) > 0 ? 1u : 0u);
        if (!masked) issue_scores(h, c);
      T_ISSUE_END()
      T_WAIT()
      float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t pk[16];
        if (masked) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0x3F803F80u;
          d2.x += 16.f; d2.y += 16.f;
        } else {
          float y[32];
          umma::ld32(tq + TC_R + 32 * half, y);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 e = make_float2(ex2_approx(y[2 * i]), ex2_approx(y[2 * i + 1]));
            pk[i] = umma::pack2(e.x, e.y);
            d2 = fadd2(d2, e);        // denominator from the unrounded terms (the bf16 rounding of P averages out: <= 2^-9 relative)
          }
        }
        st16t(tq + TC_R + 16 * half, pk);
      }
      den[h] += d2.x + d2.y;
      T_SYNC_TMEM()
    }
    T_ISSUE_BEGIN()
      issue_pv((NR - 1) / NBLK, (NR - 1) % NBLK, 1u);
    T_ISSUE_END()
    T_WAIT()
    float ao[16];
    umma::ld16(tq + TC_O + 16 * rr, ao);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float inv = 1.f / den[h];
#pragma unroll
      for (int d = 0; d < 4; ++d) ao[4 * h + d] *= inv;
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
      RV[RV2_SIGMA * RRV + m] = fmaxf(s4[0], 0.f);
    }
    umma::fence_before_sync();
    wgt_sync(wg);
    // ---- per ray: alpha compositing (+ fine resampling), one warp per ray
    for (int r = wq; r < RPT; r += 4) {
      const long long ray = (long long)tile * RPT + r;
      if (ray >= a.rn) continue;
      composite_ray(a, ray, r, dn, lane, RV, RGB, TROWS);
    }
    wgt_sync(wg);
    tile = next_tile;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem_base_s, 512);
}

int launch_render_rays_tc(const pgrf_render_args& a, int V, long long total, int sms, cudaStream_t st) {
  RaysTcParams p;
  p.a = a; p.V = V; p.total = total;
  const int rpt = 128 / a.dn;
  p.n_tiles = (a.rn + rpt - 1) / rpt;
  int dev = 0;
  PGRF_CUDA(cudaGetDevice(&dev));
  static bool done[64] = {};
  if (!done[dev & 63]) {
    PGRF_CUDA(cudaFuncSetAttribute(render_rays_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMT_BYTES));
    PGRF_CUDA(cudaFuncSetAttribute(render_rays_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMT_BYTES));
    done[dev & 63] = true;
  }
  const int grid = min((p.n_tiles + kWGt - 1) / kWGt, sms);
  if (rpt == 2) render_rays_tc_kernel<2><<<grid, kThreadsT, SMT_BYTES, st>>>(p);
  else render_rays_tc_kernel<1><<<grid, kThreadsT, SMT_BYTES, st>>>(p);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

}  // namespace pgrf

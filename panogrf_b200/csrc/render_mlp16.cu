// bf16 tensor-core variant of the per-(view,sample) and per-sample stages of the render path (K2 + K3a + K3b):
// projection + gathers + dist decoder + compute_prob + aggregation MLP, ONE persistent tcgen05 kernel.
//
// Round-2 structure:
//   * CTA = 4 independent warpgroups (16 warps, 128 registers); a warpgroup owns one tile of 128 (view,sample) rows at a time,
//     thread <-> row <-> TMEM lane for geometry, every epilogue and all element-wise state (registers);
//   * every hidden activation is written straight back to TENSOR MEMORY as the bf16 A operand of the next tcgen05.mma
//     (tcgen05.st, thread = lane): no shared-memory round trip, no proxy fence; shared memory only holds what other threads
//     need (cooperatively gathered features, per-view features for the cross-view pooling, pooled statistics, fp32 x);
//   * biases are an extra K-step against a constant "ones" chunk, odd K chunks are paired with a constant zero chunk through the
//     descriptor's leading-byte offset, ELU layers feeding another layer are pre-scaled by log2 e (render_layout16.cuh): the
//     SIMT side of a hidden layer is  tcgen05.ld -> 2 MUFU.EX2 + FFMA2 + 2 F2FP + 2 HMNMX2 per pair -> tcgen05.st;
//   * all epilogues are straight-line code with compile-time TMEM columns and shared-memory offsets.
// Reference semantics: render_rows_kernel + render_samples_kernel (render_kernels.cu header): render_ops.py:76-257,
// ops.py:32-52, dist_decoder.py:64-140, aggregate_net.py:41-89, ibrnet.py:315-351.
// Numerics: operands rounded to bf16, accumulation and all non-GEMM math in fp32 (north star: rtol 1e-2).
#include <type_traits>

#include "render_device.cuh"
#include "render_layout16.cuh"
#include "umma.cuh"

#define W16(L) (std::integral_constant<int, pgrf::w16_offset(L)>::value)
#define BC16(L) (std::integral_constant<int, pgrf::bias16_offset(L)>::value)
#define NP16(L) (std::integral_constant<int, pgrf::kLayers16[L].Npad>::value)
#define SMW(L) (std::integral_constant<int, pgrf::small16_offset(L)>::value)
#define SMB(L) (std::integral_constant<int, pgrf::small16_bias_offset(L)>::value)

namespace pgrf {

constexpr int kWG = 4;                      // warpgroups per CTA
constexpr int kThreads16 = 128 * kWG;
constexpr int ROWS = 128;                   // rows per tile == operand row pitch
constexpr int CH = ROWS * 16;               // bytes per k-chunk of an A operand

// ---- per-warpgroup shared memory map (bytes) ----
//  X early : RF = chunks 0..3 (ray_feats)
//  X mid   : pooled view statistics, mean = chunks 0..4, variance = chunks 5..9 (one weighting at a time)
//  X late  : x in fp32 [32][128] (chunks 0..7)
//  Y       : img_feats 0..3 | rgb 4  -> rgb_feat' (f' order) = per-row K-block of base_fc.0 and input of the pooling
//  S       : 2 x 128 footprint records during geometry / gather; SF float vectors afterwards
constexpr int X_BYTES = 10 * CH, Y_BYTES = 5 * CH, S_BYTES = 2 * CH;
constexpr int WG_BYTES = X_BYTES + Y_BYTES + S_BYTES;
enum { SF_W0 = 0, SF_VIS2, SF_LOGIT, SF_R, SF_G, SF_B };

constexpr int SM16_W = 0;
constexpr int SM16_WG = (kW16Sec0Bytes + 127) & ~127;
constexpr int SM16_ONES = SM16_WG + kWG * WG_BYTES;     // [128][8] bf16 = {1, 1, 0, 0, 0, 0, 0, 0}
constexpr int SM16_ZERO = SM16_ONES + CH;               // 2 KB of zeros, ABOVE every operand (LBO is an unsigned offset)
constexpr int SM16_BAR = SM16_ZERO + CH;
constexpr int SM16_BYTES = SM16_BAR + 64;
static_assert(SM16_BYTES + 1024 <= 227 * 1024, "fused MLP kernel shared memory");

// ---- per-warpgroup tensor-memory columns (128 of the CTA's 512) ----
constexpr int T_RDIN = 104;                                          // [dir_diff(4), 1, 1, 0 ...]: K = 16 step of ray_dir_fc.0
constexpr int C_RD0 = 112, T_HRD = 120;                              // ray_dir_fc.0 accumulator (16) -> its bf16 hidden (8 columns)
constexpr int C_MEAN0 = 0, C_VAR0 = 32, C_RD1 = 64;                  // stage 1
constexpr int T_H0 = 0, T_H1 = 16;
constexpr int C_MEAN1 = 32, C_VAR1 = 64, C_AW0 = 96;                 // stage 2
constexpr int T_H2 = 0;
constexpr int C_AW1 = 32, C_VIS0 = 64;                               // stage 3
constexpr int T_H3 = 0;
constexpr int C_VIS1 = 32;                                           // stage 4 (use_vis)
constexpr int C_PE0 = 0, T_HPE = 32, C_PE1 = 0, T_PEMB = 32;         // prob_embed
constexpr int T_HVT = 48;                                            // [hit', vis', 1, 1, 0 ...]: last K-step of prob_embed.0
constexpr int C_BASE0 = 64, C_NF0 = 48, T_H64 = 0;                   // base_fc.0 (N = 64), neuray_fc.0 (N = 16)
constexpr int C_BASE1 = 32, T_HV = 0;                                // base_fc.2
constexpr int C_VFC0 = 64, T_HV2 = 16, C_VFC1 = 64;                  // vis_fc
constexpr int T_HVP = 0, T_RGX = 16, T_RGT = 32;                     // x * vis | x | [vis2, dir_diff, 1, 1]
constexpr int C_V2 = 64, C_RGB0 = 96;                                // vis_fc2.0, rgb_fc.0

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + fast_exp(x)); }
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
// parity wait with a suspend-time hint: the hardware parks the warp instead of re-polling every few hundred cycles
__device__ __forceinline__ void mbar_wait_park(uint32_t bar_addr, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar_addr), "r"(phase), "r"(20000u) : "memory");
}

__device__ __forceinline__ void mbar_wait_poll(uint32_t bar_addr, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar_addr), "r"(phase) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- packed helpers ----
__device__ __forceinline__ uint32_t pack_relu2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// y = log2e * x (pre-scaled accumulator incl. bias) -> bf16x2 of log2e * ELU(x) = max(y, min(log2e * 2^y - log2e, 0))
__device__ __forceinline__ uint32_t elu2p_pack(float y0, float y1) {
  // ELU'(y) = relu(y) + min(g, 0) = relu(y) - relu(-g) with -g = log2e - log2e * 2^y.  Both relus are free in the packing conversion
  // (cvt.rn.relu.bf16x2.f32) and exactly one of the two terms is non-zero, so the packed subtraction is exact:
  // 2 MUFU + FFMA2 + 2 F2FP + HADD2 per pair (was 2 F2FP + 2 HMNMX2 after the FFMA2)
  const float2 ng = ffma2(make_float2(ex2_approx(y0), ex2_approx(y1)), make_float2(-kLog2e, -kLog2e), make_float2(kLog2e, kLog2e));
  const uint32_t yr = pack_relu2(y0, y1), r = pack_relu2(ng.x, ng.y);
  const __nv_bfloat162 o = __hsub2(*reinterpret_cast<const __nv_bfloat162*>(&yr), *reinterpret_cast<const __nv_bfloat162*>(&r));
  return *reinterpret_cast<const uint32_t*>(&o);
}
// same in fp32 (the value feeds a register GEMV): log2e * ELU(x)
__device__ __forceinline__ float2 elu2p_f32(float y0, float y1) {
  const float2 g = ffma2(make_float2(ex2_approx(fminf(y0, 0.f)), ex2_approx(fminf(y1, 0.f))), make_float2(kLog2e, kLog2e),
                         make_float2(-kLog2e, -kLog2e));
  return make_float2(fmaxf(y0, g.x), fmaxf(y1, g.y));
}
// plain ELU of an un-scaled pair: max(x, 2^(log2e * min(x, 0)) - 1)
__device__ __forceinline__ float2 elu_plain2(float x0, float x1) {
  const float2 t = fmul2(make_float2(fminf(x0, 0.f), fminf(x1, 0.f)), make_float2(kLog2e, kLog2e));
  const float2 g = fadd2(make_float2(ex2_approx(t.x), ex2_approx(t.y)), make_float2(-1.f, -1.f));
  return make_float2(fmaxf(x0, g.x), fmaxf(x1, g.y));
}
__device__ __forceinline__ float elu_plain1(float x) { return fmaxf(x, ex2_approx(fminf(x, 0.f) * kLog2e) - 1.f); }

// ---- TMEM access (thread = lane) ----
__device__ __forceinline__ void ld32f(uint32_t taddr, float (&v)[32]) { umma::ld32(taddr, v); }
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// accumulator columns [col, col+32) -> bf16( log2e * ELU ) -> A-operand columns [hcol, hcol+16)
__device__ __forceinline__ void epi_elu2p_tmem(uint32_t tq, int col, int hcol) {
  float y[32];
  ld32f(tq + col, y);
  uint32_t h[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) h[i] = elu2p_pack(y[2 * i], y[2 * i + 1]);
  st16(tq + hcol, h);
}
// accumulator columns [col, col+32) -> log2e * ELU in fp32 -> NOUT dot products with fp32 rows Wsm[j][32] (ln 2 folded in)
template <int NOUT>
__device__ __forceinline__ void epi_elu2p_gemv(uint32_t tq, int col, const float* __restrict__ Wsm, float (&out)[NOUT]) {
  float y[32];
  ld32f(tq + col, y);
  float2 acc[NOUT];
#pragma unroll
  for (int j = 0; j < NOUT; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 a = elu2p_f32(y[4 * i], y[4 * i + 1]), b = elu2p_f32(y[4 * i + 2], y[4 * i + 3]);
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
      const float4 w = *reinterpret_cast<const float4*>(Wsm + j * 32 + 4 * i);
      acc[j] = ffma2(a, make_float2(w.x, w.y), acc[j]);
      acc[j] = ffma2(b, make_float2(w.z, w.w), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < NOUT; ++j) out[j] = acc[j].x + acc[j].y;
}

__device__ __forceinline__ void zero_chunk(unsigned char* dst, int chunk, int m) {
  *reinterpret_cast<uint4*>(dst + ((size_t)chunk * ROWS + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
}

// Footprint record of one row for the cooperative gathers: everything the 8 lanes of a row would otherwise each recompute (neighbour
// offsets resolved from the border flags, the four bilinear weights) is computed once by the row's own thread.  32 bytes = 2 LDS.128.
struct __align__(16) FootRec {
  int off;      // texel index of the north-west tap inside the stacked (rfn*h*w) map
  int oe, os;   // float4 offsets of the east / south neighbour (0 when the neighbour is outside the map: its weight is 0 there)
  int pad;
  float w0, w1, w2, w3;   // (1-tx)(1-ty), tx(1-ty), (1-tx)ty, tx ty
};
__device__ __forceinline__ FootRec make_foot_rec(int view_base, const Footprint& f, int map_w) {
  FootRec r;
  r.off = view_base + f.off;
  r.oe = f.dx ? 8 : 0;
  r.os = f.dy ? map_w * 8 : 0;
  r.pad = 0;
  const float tx1 = 1.f - f.tx, ty1 = 1.f - f.ty;
  r.w0 = tx1 * ty1; r.w1 = f.tx * ty1; r.w2 = tx1 * f.ty; r.w3 = f.tx * f.ty;
  return r;
}
__device__ __forceinline__ float4 tap4_rec(const float4* __restrict__ base, const FootRec& f) {
  const float4 nw = ldg4(base), ne = ldg4(base + f.oe), sw = ldg4(base + f.os), se = ldg4(base + f.os + f.oe);
  // same products and accumulation order (nw, ne, sw, se) as the scalar form, two channels per instruction
  const float2 w0 = make_float2(f.w0, f.w0), w1 = make_float2(f.w1, f.w1), w2 = make_float2(f.w2, f.w2), w3 = make_float2(f.w3, f.w3);
  float2 lo = fmul2(make_float2(nw.x, nw.y), w0), hi = fmul2(make_float2(nw.z, nw.w), w0);
  lo = ffma2(make_float2(ne.x, ne.y), w1, lo); hi = ffma2(make_float2(ne.z, ne.w), w1, hi);
  lo = ffma2(make_float2(sw.x, sw.y), w2, lo); hi = ffma2(make_float2(sw.z, sw.w), w2, hi);
  lo = ffma2(make_float2(se.x, se.y), w3, lo); hi = ffma2(make_float2(se.z, se.w), w3, hi);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// ---- MMA issue (ONE thread).  All operands K-major, no swizzle, SBO = 128 B; a K = 16 step reads the two 8-element chunks at
// `addr` and `addr + lbo` — an odd trailing chunk is paired with the shared zero chunk by choosing lbo = zero - addr.
// The issuing thread is the critical path of every stage, so a descriptor costs ONE add: its low word is
//   (addr >> 4) | (lbo >> 4) << 16  =  per-buffer base (computed once per kernel) + compile-time immediate,
// with two bases per buffer: `n` for chunk pairs inside the buffer (constant lbo) and `z` for a chunk paired with the zero chunk
// (lbo = zero - addr, i.e. base16 + ((z16 - base16) << 16), immediate off16 - (off16 << 16)). ----
struct Issuer {
  uint32_t wn, wz;        // weight section
  uint32_t xn, xz, yn, yz;   // the warpgroup's X / Y operand regions
  uint32_t ones_lo;       // A = [ones | zero]
  static constexpr uint64_t kHi = (uint64_t)0x4008 << 32;     // SBO = 128 B, descriptor version 1, no swizzle
  __device__ __forceinline__ void init(uint32_t sW, uint32_t sX, uint32_t sY, uint32_t sO, uint32_t sZ) {
    const uint32_t z16 = sZ >> 4;
    wn = sW >> 4; wz = wn + ((z16 - wn) << 16);
    xn = sX >> 4; xz = xn + ((z16 - xn) << 16);
    yn = sY >> 4; yz = yn + ((z16 - yn) << 16);
    ones_lo = (sO >> 4) + ((z16 - (sO >> 4)) << 16);
  }
  static __device__ __forceinline__ constexpr uint32_t imm_n(uint32_t off, uint32_t pitch) { return (off >> 4) | ((pitch >> 4) << 16); }
  static __device__ __forceinline__ constexpr uint32_t imm_z(uint32_t off) { return (off >> 4) - ((off >> 4) << 16); }
  __device__ __forceinline__ void ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, int N, uint32_t acc) const {
    umma::mma_bf16(d, kHi | a_lo, kHi | b_lo, umma::instr_desc_bf16(128, N), acc);
  }
  __device__ __forceinline__ void ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo, int N, uint32_t acc) const {
    umma::mma_bf16_ts(d, a_tmem, kHi | b_lo, umma::instr_desc_bf16(128, N), acc);
  }
  // A = NCH consecutive chunks of the X (AX = true) or Y operand region starting at byte offset `a_off`, B = chunks
  // [kc0, kc0 + NCH) of the layer at byte offset `w_off` of the weight section (Npad = N rows)
  template <int NCH, bool AX>
  __device__ __forceinline__ void smem_chunks(uint32_t d, uint32_t a_off, uint32_t w_off, int kc0, int N, uint32_t acc_first) const {
    const uint32_t wp = (uint32_t)N * 16;
    const uint32_t an = AX ? xn : yn, az = AX ? xz : yz;
#pragma unroll
    for (int c = 0; c < NCH; c += 2) {
      const uint32_t ao = a_off + c * CH, bo = w_off + (kc0 + c) * wp;
      if (c + 1 < NCH) ss(d, an + imm_n(ao, CH), wn + imm_n(bo, wp), N, (c > 0) ? 1u : acc_first);
      else ss(d, az + imm_z(ao), wz + imm_z(bo), N, (c > 0) ? 1u : acc_first);
    }
  }
  // A = NCH chunks held in tensor memory (4 columns per chunk) starting at column address `a_tmem`
  template <int NCH>
  __device__ __forceinline__ void tmem_chunks(uint32_t d, uint32_t a_tmem, uint32_t w_off, int kc0, int N, uint32_t acc_first) const {
    static_assert(NCH % 2 == 0, "tensor-memory operands are written in whole K = 16 steps");
    const uint32_t wp = (uint32_t)N * 16;
#pragma unroll
    for (int c = 0; c < NCH; c += 2) ts(d, a_tmem + 4 * c, wn + imm_n(w_off + (kc0 + c) * wp, wp), N, (c > 0) ? 1u : acc_first);
  }
  // tensor-memory operand whose SECOND chunk is absent (K = 8 real columns): B pairs the weight chunk with the zero chunk
  __device__ __forceinline__ void tmem_half(uint32_t d, uint32_t a_tmem, uint32_t w_off, int N, uint32_t acc) const {
    ts(d, a_tmem, wz + imm_z(w_off), N, acc);
  }
  // + bias: A = [ones | zero], B = [bias chunk | zero]
  __device__ __forceinline__ void bias(uint32_t d, uint32_t b_off, int N) const { ss(d, ones_lo, wz + imm_z(b_off), N, 1u); }
};

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
  int Mv;                 // samples per tile of the rays kernel (whole rays): rows of one F2 operand tile
  int wait_mode;          // 0: every warp polls the MMA mbarrier; 1: the issuing warp polls, the others wait in a named barrier
  unsigned dn_magic, dn_shift, mv_magic, mv_shift;   // n / a.dn and n / Mv as multiply-high + shift (fastdiv_gen)
};

// Unsigned 32-bit division by a launch constant without the ~20-instruction udiv sequence: q = (t + ((n - t) >> 1)) >> shift,
// t = mulhi(magic, n)  (the "branch-free" form of division by invariant integers; exact for every n < 2^32, d >= 2).
static void fastdiv_gen(unsigned d, unsigned* magic, unsigned* shift) {
  int L = 31;
  while (!((d >> L) & 1u)) --L;                                    // floor(log2 d)
  if ((d & (d - 1u)) == 0u) { *magic = 0u; *shift = (unsigned)(L - 1); return; }
  const unsigned long long num = 1ull << (32 + L);
  unsigned long long m = num / d;
  const unsigned long long rem = num - m * d;
  m += m;
  const unsigned long long twice_rem = rem + rem;
  if (twice_rem >= d) m += 1;
  *magic = (unsigned)(m + 1);
  *shift = (unsigned)L;
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned magic, unsigned shift) {
  const unsigned t = __umulhi(magic, n);
  return (t + ((n - t) >> 1)) >> shift;
}

template <int V>
__global__ void __launch_bounds__(kThreads16, 1) render_mlp_bf16_kernel(const Render16Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_tile[kWG];
  const pgrf_render_args& a = p.a;
  const int tid = threadIdx.x;
  // warp-uniform indices go through a shuffle so that the compiler keeps everything derived from them (shared-memory and
  // tensor-memory addresses of the MMA descriptors) in uniform registers: no per-MMA R2UR waterfall in the issuing warp
  const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);            // warpgroup
  const int m = tid & 127;                                         // row inside the warpgroup's tile
  const int wq = __shfl_sync(0xffffffffu, (tid >> 5) & 3, 0);      // warp inside the warpgroup -> TMEM lane quadrant
  unsigned char* Wb = smem + SM16_W;
  const float* Wsm = reinterpret_cast<const float*>(Wb + kW16SmallBegin0);     // fp32 tiny output layers
  unsigned char* X = smem + SM16_WG + wg * WG_BYTES;
  unsigned char* Y = X + X_BYTES;
  float* SF = reinterpret_cast<float*>(Y + Y_BYTES);
  FootRec* FP = reinterpret_cast<FootRec*>(X + 4 * CH);   // [2][128] records of 32 bytes in X chunks 4..7: free between the tile's start and the view pooling
  float* XF = reinterpret_cast<float*>(X);        // x in fp32 [32][128]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM16_BAR) + wg;
  const uint32_t bar_addr = umma::smem_addr(bar);

  // one-time setup: weights -> smem, constant chunks, TMEM allocation (512 columns: 128 per warpgroup), mbarriers
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.weights16);
    uint4* dst = reinterpret_cast<uint4*>(Wb);
    for (int i = tid; i < kW16Sec0Bytes / 16; i += kThreads16) dst[i] = __ldg(src + i);
    if (tid < ROWS) {
      const uint32_t one2 = 0x3F803F80u;   // bf16 {1, 1}
      reinterpret_cast<uint4*>(smem + SM16_ONES)[tid] = make_uint4(one2, 0u, 0u, 0u);
      reinterpret_cast<uint4*>(smem + SM16_ZERO)[tid] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (tid < 32) umma::tmem_alloc(&tmem_base_s, 512);
  if (m == 0) mbar_init(bar, 1);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base_s, 0) + wg * 128;   // this warpgroup's TMEM columns
  const uint32_t tq = tb + ((uint32_t)(wq * 32) << 16);       // + this warp's lane quadrant
  uint32_t phase = 0;
  Issuer is;
  is.init(__shfl_sync(0xffffffffu, umma::smem_addr(Wb), 0), __shfl_sync(0xffffffffu, umma::smem_addr(X), 0),
          __shfl_sync(0xffffffffu, umma::smem_addr(Y), 0), __shfl_sync(0xffffffffu, umma::smem_addr(smem + SM16_ONES), 0),
          __shfl_sync(0xffffffffu, umma::smem_addr(smem + SM16_ZERO), 0));

  const int T = p.T, M = p.M;
  const int v = min(m / T, V - 1), t = m % T;
  const int vrow = m / T;             // unclamped: rows beyond the V*T valid ones take no share of the pooling work
  const float wgt = 1.f / ((float)V + 1e-8f);

  int static_tile = blockIdx.x * kWG + wg;
  // per-thread constants of the whole launch (the view of a row never changes: v = m / T)
  const float inv_wm1 = 1.f / (float)(a.img_w - 1), inv_hm1 = 1.f / (float)(a.img_h - 1);
  const float q_nn = -1.f / a.que_near, q_inv = 1.f / (-1.f / a.que_far - q_nn);
  const float r_nn = -1.f / __ldg(a.ref_depth_range + 2 * v), r_inv = 1.f / (-1.f / __ldg(a.ref_depth_range + 2 * v + 1) - r_nn);
  unsigned char* f2_op = reinterpret_cast<unsigned char*>(a.f2);                  // bf16 operand tiles of the rays kernel
  float4* f2_rgb = reinterpret_cast<float4*>(f2_op + (size_t)((p.total + p.Mv - 1) / p.Mv) * kF2TileBytes);

#define ISSUE_BEGIN() if (wq == 0) { if (elect_one()) { umma::fence_after_sync();
#define ISSUE_END() umma::commit(bar); } __syncwarp(); }
  // completion of the committed MMAs: either every warp polls the mbarrier, or (wait_mode 1) the issuing warp polls and the other
  // three sleep in the warpgroup's named barrier (no issue slots spent on polling)
#define WAIT_MMA() { if (p.wait_mode == 0 || wq == 0) mbar_wait_poll(bar_addr, phase); if (p.wait_mode) wg_sync(wg); phase ^= 1; umma::fence_after_sync(); }
  // publish operand writes (shared memory: generic -> async proxy; tensor memory: wait::st) and sync the warpgroup
#define SYNC_SMEM() { umma::fence_smem_to_async(); umma::fence_before_sync(); wg_sync(wg); }
#define SYNC_TMEM() { umma::wait_st(); umma::fence_before_sync(); wg_sync(wg); }
#define SYNC_BOTH() { umma::wait_st(); umma::fence_smem_to_async(); umma::fence_before_sync(); wg_sync(wg); }

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
    // keep the descriptor bases opaque per tile: otherwise the compiler hoists all ~140 "base + immediate" sums out of the tile
    // loop and spills them (they cost one add each where they are used)
    asm volatile("" : "+r"(is.wn), "+r"(is.wz), "+r"(is.xn), "+r"(is.xz), "+r"(is.yn), "+r"(is.yz), "+r"(is.ones_lo));
    long long g = (long long)tile * T + t;
    const bool row_valid = (m < M) && (g < p.total);
    if (g >= p.total) g = p.total - 1;
    const int ray = (int)fastdiv((unsigned)g, p.dn_magic, p.dn_shift), s = (int)((unsigned)g - (unsigned)ray * (unsigned)a.dn);   // total < 2^31

    // ------------------------------------------------------------ geometry (thread = row)
    const RowGeom rg = row_geometry<true>(a, v, g);
    {
      Footprint f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, a.rf_h == a.img_h && a.rf_w == a.img_w, a.rf_h, a.rf_w);
      FP[m] = make_foot_rec(v * a.rf_h * a.rf_w, f, a.rf_w);
      f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, a.if_h == a.img_h && a.if_w == a.img_w, a.if_h, a.if_w);
      FP[ROWS + m] = make_foot_rec(v * a.if_h * a.if_w, f, a.if_w);
    }
    {   // input of ray_dir_fc.0, one K = 16 step in tensor memory: [dir_diff(4), 1, 1, 0 x10] (the two ones carry the bias)
      const uint32_t h[8] = {umma::pack2(rg.dirdiff[0], rg.dirdiff[1]), umma::pack2(rg.dirdiff[2], rg.dirdiff[3]), 0x3F803F80u, 0u, 0u, 0u, 0u, 0u};
      umma::st8(tq + T_RDIN, h);
    }
    // own-row colour taps (fp32, kept in registers for the final blend)
    float rgb_in[3];
    {
      const Footprint f = border_footprint_r(rg.px, rg.py, inv_wm1, inv_hm1, true, a.img_h, a.img_w);
      const float4 c = tap4(reinterpret_cast<const float4*>(a.imgs_cl) + (size_t)v * a.img_h * a.img_w + f.off, f, 1, a.img_w);
      rgb_in[0] = c.x; rgb_in[1] = c.y; rgb_in[2] = c.z;
      if (a.wo_appearance) { rgb_in[0] = 0.f; rgb_in[1] = 0.f; rgb_in[2] = 0.f; }      // aggregate_net.py:79-81
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
    // ------------------------------------------------------------ stage 0: ray_dir_fc.0 (runs under the gathers)
    SYNC_TMEM()      // also publishes the footprint records to the warpgroup
    ISSUE_BEGIN()
      is.tmem_half(tb + C_RD0, tb + T_RDIN, W16(M_RD0), 16, 0u);
    ISSUE_END()

    // ------------------------------------------------------------ cooperative gathers (lane = (row, float4 group))
    // each row's two bilinear footprints were computed once by its own thread (FP records in smem); here 8 lanes
    // per row fetch the 4 x 128-byte taps of both feature maps and blend
#pragma unroll 2
    for (int it = m; it < ROWS * 8; it += 128) {
      const int r = it >> 3, cg = it & 7;
      const FootRec f1 = FP[r], f2 = FP[ROWS + r];
      const float4* b1 = reinterpret_cast<const float4*>(a.ray_feats_cl) + (size_t)f1.off * 8 + cg;
      const float4* b2 = reinterpret_cast<const float4*>(a.img_feats_cl) + (size_t)f2.off * 8 + cg;
      const float4 rf = tap4_rec(b1, f1);
      const float4 imf = tap4_rec(b2, f2);
      uint2 q;   // 4 channels = half a chunk: chunk cg/2, 8-byte half cg%2
      q.x = umma::pack2(rf.x, rf.y); q.y = umma::pack2(rf.z, rf.w);
      *reinterpret_cast<uint2*>(X + ((size_t)(cg >> 1) * ROWS + r) * 16 + (cg & 1) * 8) = q;
      q.x = umma::pack2(imf.x, imf.y); q.y = umma::pack2(imf.z, imf.w);
      if (a.wo_appearance) q = make_uint2(0u, 0u);
      *reinterpret_cast<uint2*>(Y + ((size_t)(cg >> 1) * ROWS + r) * 16 + (cg & 1) * 8) = q;
    }
    WAIT_MMA()
    {   // ray_dir_fc.0 hidden (16, ELU) -> tensor memory
      float y[16];
      umma::ld16(tq + C_RD0, y);
      uint32_t h[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = elu2p_pack(y[2 * i], y[2 * i + 1]);
      umma::st8(tq + T_HRD, h);
    }
    // ------------------------------------------------------------ stage 1: mean_decoder.0, var_decoder.0, ray_dir_fc.2
    SYNC_BOTH()
    ISSUE_BEGIN()
      is.smem_chunks<4, true>(tb + C_MEAN0, 0, W16(M_MEAN0), 0, 32, 0u); is.bias(tb + C_MEAN0, BC16(M_MEAN0), 32);
      is.smem_chunks<4, true>(tb + C_VAR0, 0, W16(M_VAR0), 0, 32, 0u);   is.bias(tb + C_VAR0, BC16(M_VAR0), 32);
      is.tmem_chunks<2>(tb + C_RD1, tb + T_HRD, W16(M_RD1), 0, 48, 0u); is.bias(tb + C_RD1, BC16(M_RD1), 48);
    ISSUE_END()
    WAIT_MMA()
    epi_elu2p_tmem(tq, C_MEAN0, T_H0);
    epi_elu2p_tmem(tq, C_VAR0, T_H1);
    // direction feature (f' order: img_feats 0..31, rgb 32..34): rgb_feat = [img_feats, rgb] + ELU(ray_dir_fc)
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      float df[8], x[8];
      umma::ld8(tq + C_RD1 + 8 * c, df);
      if (c < 4) {
        umma::load_chunk(Y, ROWS, c, m, x);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = i < 3 ? rgb_in[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float2 d = elu_plain2(df[i], df[i + 1]);
        const float2 xs = fadd2(make_float2(x[i], x[i + 1]), d);
        x[i] = (c < 4 || i < 3) ? xs.x : 0.f;
        x[i + 1] = (c < 4 || i + 1 < 3) ? xs.y : 0.f;
      }
      umma::store_chunk(Y, ROWS, c, m, x);
    }
    // ------------------------------------------------------------ stage 2: mean_decoder.2, var_decoder.2, aw_decoder.0
    SYNC_BOTH()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_MEAN1, tb + T_H0, W16(M_MEAN1), 0, 32, 0u); is.bias(tb + C_MEAN1, BC16(M_MEAN1), 32);
      is.tmem_chunks<4>(tb + C_VAR1, tb + T_H1, W16(M_VAR1), 0, 32, 0u);   is.bias(tb + C_VAR1, BC16(M_VAR1), 32);
      is.smem_chunks<4, true>(tb + C_AW0, 0, W16(M_AW0), 0, 32, 0u);            is.bias(tb + C_AW0, BC16(M_AW0), 32);
    ISSUE_END()
    WAIT_MMA()
    float mean[2], var[2], aw, visd = 1.f;
    // aw_decoder.0's hidden first: it is the only input of stage 3, whose MMAs then run under the two output-layer epilogues below
    epi_elu2p_tmem(tq, C_AW0, T_H2);
    // ------------------------------------------------------------ stage 3: aw_decoder.2 (+ vis_decoder.0)
    SYNC_TMEM()
    // without the vis decoder the accumulator takes aw_decoder.0's columns (just consumed): mean/var accumulators stay readable
    const int c_aw1 = a.use_vis ? C_AW1 : C_AW0;
    if (!a.use_vis) {
      ISSUE_BEGIN()
        is.tmem_chunks<4>(tb + C_AW0, tb + T_H2, W16(M_AW1), 0, 32, 0u); is.bias(tb + C_AW0, BC16(M_AW1), 32);
      ISSUE_END()
    }
    {
      float o2[2];
      epi_elu2p_gemv<2>(tq, C_MEAN1, Wsm + SMW(M_MEAN2), o2);
      mean[0] = softplus_fast(o2[0] + Wsm[SMB(M_MEAN2)]); mean[1] = softplus_fast(o2[1] + Wsm[SMB(M_MEAN2) + 1]);
      epi_elu2p_gemv<2>(tq, C_VAR1, Wsm + SMW(M_VAR2), o2);
      var[0] = softplus_fast(o2[0] + Wsm[SMB(M_VAR2)]) + a.bias_val; var[1] = softplus_fast(o2[1] + Wsm[SMB(M_VAR2) + 1]) + a.bias_val;
    }
    if (a.use_vis) {
      umma::fence_before_sync();
      wg_sync(wg);                 // every thread has read the mean / var accumulators that stage 3 overwrites
      ISSUE_BEGIN()
        is.tmem_chunks<4>(tb + C_AW1, tb + T_H2, W16(M_AW1), 0, 32, 0u); is.bias(tb + C_AW1, BC16(M_AW1), 32);
        is.smem_chunks<4, true>(tb + C_VIS0, 0, W16(M_VIS0), 0, 32, 0u); is.bias(tb + C_VIS0, BC16(M_VIS0), 32);
      ISSUE_END()
    }
    WAIT_MMA()
    {
      float o1[1];
      epi_elu2p_gemv<1>(tq, c_aw1, Wsm + SMW(M_AW2), o1);
      aw = sigmoidf(o1[0] + Wsm[SMB(M_AW2)]);
    }
    if (a.use_vis) {   // 4th decoder
      epi_elu2p_tmem(tq, C_VIS0, T_H3);
      SYNC_TMEM()
      ISSUE_BEGIN()
        is.tmem_chunks<4>(tb + C_VIS1, tb + T_H3, W16(M_VIS1), 0, 32, 0u); is.bias(tb + C_VIS1, BC16(M_VIS1), 32);
      ISSUE_END()
      WAIT_MMA()
      float o1[1];
      epi_elu2p_gemv<1>(tq, C_VIS1, Wsm + SMW(M_VIS2), o1);
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
      // last K-step of prob_embed.0 in tensor memory: [hit', vis', 1, 1, 0 x12] (the two ones carry the bias)
      const uint32_t h[8] = {umma::pack2((hp - 0.5f) * 2.f, (visibility - 0.5f) * 2.f), 0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u};
      umma::st8(tq + T_HVT, h);
    }

    // ------------------------------------------------------------ prob_embed 34 -> 32 (ReLU) -> 32
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.smem_chunks<4, true>(tb + C_PE0, 0, W16(M_PE0), 0, 32, 0u);
      is.tmem_half(tb + C_PE0, tb + T_HVT, W16(M_PE0) + 4 * 32 * 16, 32, 1u);
    ISSUE_END()
    WAIT_MMA()
    {
      float y[32];
      ld32f(tq + C_PE0, y);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = pack_relu2(y[2 * i], y[2 * i + 1]);
      st16(tq + T_HPE, h);
    }
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_PE1, tb + T_HPE, W16(M_PE1), 0, 32, 0u); is.bias(tb + C_PE1, BC16(M_PE1), 32);
    ISSUE_END()
    WAIT_MMA()
    {   // prob_embedding (no activation) -> tensor memory: operand of base_fc.0 and neuray_fc.0
      float y[32];
      ld32f(tq + C_PE1, y);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = a.wo_geometry ? 0u : umma::pack2(y[2 * i], y[2 * i + 1]);   // aggregate_net.py:60-62
      st16(tq + T_PEMB, h);
    }
    // ------------------------------------------------------------ base_fc.0 = per-row K-blocks [rgb_feat' | prob_embedding] + bias,
    // then the pooled K-blocks in two accumulating slices (uniform weights first: they do not need neuray_fc)
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_NF0, tb + T_PEMB, W16(M_NF0), 0, 16, 0u); is.bias(tb + C_NF0, BC16(M_NF0), 16);
      is.smem_chunks<5, false>(tb + C_BASE0, 0, W16(M_BASE0), 20, 64, 0u);
      is.tmem_chunks<4>(tb + C_BASE0, tb + T_PEMB, W16(M_BASE0), 25, 64, 1u);
      is.bias(tb + C_BASE0, BC16(M_BASE0), 64);
    ISSUE_END()
    float w0n[V];
#pragma unroll
    for (int vv = 0; vv < V; ++vv) w0n[vv] = wgt;
    pool_views<V>(Y, X, t, T, vrow, w0n);          // X early (RF, HV) is dead: prob_embed.0 completed
    WAIT_MMA()
    {   // neuray_fc 32 -> 8 (ELU) -> 1 (sigmoid): weight of this view in the first pooling
      float h8[8];
      umma::ld8(tq + C_NF0, h8);
      float o = Wsm[SMB(M_NF1)];
#pragma unroll
      for (int i = 0; i < 8; ++i) o = fmaf(elu_plain1(h8[i]), Wsm[SMW(M_NF1) + i], o);
      SF[SF_W0 * ROWS + m] = sigmoidf(o);
    }
    SYNC_SMEM()
    ISSUE_BEGIN()
      is.smem_chunks<5, true>(tb + C_BASE0, 0, W16(M_BASE0), 10, 64, 1u);              // mean1 (uniform)
      is.smem_chunks<5, true>(tb + C_BASE0, 5 * CH, W16(M_BASE0), 15, 64, 1u);     // var1
    ISSUE_END()
#pragma unroll
    for (int vv = 0; vv < V; ++vv) w0n[vv] = SF[SF_W0 * ROWS + vv * T + t] * wgt;
    WAIT_MMA()
    pool_views<V>(Y, X, t, T, vrow, w0n);
    SYNC_SMEM()
    ISSUE_BEGIN()
      is.smem_chunks<5, true>(tb + C_BASE0, 0, W16(M_BASE0), 0, 64, 1u);               // mean0 (neuray-weighted)
      is.smem_chunks<5, true>(tb + C_BASE0, 5 * CH, W16(M_BASE0), 5, 64, 1u);      // var0
    ISSUE_END()
    WAIT_MMA()
    epi_elu2p_tmem(tq, C_BASE0, T_H64);
    epi_elu2p_tmem(tq, C_BASE0 + 32, T_H64 + 16);

    // ------------------------------------------------------------ base_fc.2 -> x (fp32, registers)
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<8>(tb + C_BASE1, tb + T_H64, W16(M_BASE1), 0, 32, 0u); is.bias(tb + C_BASE1, BC16(M_BASE1), 32);
    ISSUE_END()
    WAIT_MMA()
    float x[32];
    {
      ld32f(tq + C_BASE1, x);
      uint32_t h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 xv = elu_plain2(x[2 * i], x[2 * i + 1]);
        x[2 * i] = xv.x; x[2 * i + 1] = xv.y;
        const float2 sv = fmul2(xv, make_float2(wgt, wgt));
        h[i] = umma::pack2(sv.x, sv.y);
      }
      st16(tq + T_HV, h);
    }
    // ------------------------------------------------------------ vis_fc(x * weight) 32 -> 32 -> 33
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_VFC0, tb + T_HV, W16(M_VFC0), 0, 32, 0u); is.bias(tb + C_VFC0, BC16(M_VFC0), 32);
    ISSUE_END()
    WAIT_MMA()
    epi_elu2p_tmem(tq, C_VFC0, T_HV2);
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_VFC1, tb + T_HV2, W16(M_VFC1), 0, 48, 0u); is.bias(tb + C_VFC1, BC16(M_VFC1), 48);
    ISSUE_END()
    WAIT_MMA()
    {
      float vr, vr1;
      umma::ld2(tq + C_VFC1 + 32, vr, vr1);
      const float vis1 = sigmoidf(elu_plain1(vr));   // vis = sigmoid(vis) * mask
#pragma unroll
      for (int half = 0; half < 2; ++half) {       // two halves of 16 columns: x[32] stays live, keep the rest of the working set small
        float r[16];
        umma::ld16(tq + C_VFC1 + 16 * half, r);
        uint32_t h[8], hx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = 16 * half + 2 * i;
          const float2 xv = fadd2(make_float2(x[k], x[k + 1]), elu_plain2(r[2 * i], r[2 * i + 1]));   // x = x + x_res
          x[k] = xv.x; x[k + 1] = xv.y;
          XF[k * ROWS + m] = xv.x;
          XF[(k + 1) * ROWS + m] = xv.y;
          const float2 sv = fmul2(xv, make_float2(vis1, vis1));
          h[i] = umma::pack2(sv.x, sv.y);
          hx[i] = umma::pack2(xv.x, xv.y);
        }
        umma::st8(tq + T_HVP + 8 * half, h);
        umma::st8(tq + T_RGX + 8 * half, hx);
      }
    }
    // ------------------------------------------------------------ vis_fc2(x * vis) 32 -> 32 -> 1 and the x K-block of rgb_fc.0
    SYNC_TMEM()
    ISSUE_BEGIN()
      is.tmem_chunks<4>(tb + C_V2, tb + T_HVP, W16(M_V2_0), 0, 32, 0u); is.bias(tb + C_V2, BC16(M_V2_0), 32);
      is.tmem_chunks<4>(tb + C_RGB0, tb + T_RGX, W16(M_RGB0), 0, 16, 0u);
    ISSUE_END()
    WAIT_MMA()
    {
      float o1[1];
      epi_elu2p_gemv<1>(tq, C_V2, Wsm + SMW(M_V2_1), o1);
      const float vis2 = sigmoidf(o1[0] + Wsm[SMB(M_V2_1)]);
      SF[SF_VIS2 * ROWS + m] = vis2;
      // last K-chunk of rgb_fc.0: [vis, ray_diff(4), 1, 1, 0] (the ones carry the bias); the absent second chunk of the K = 16 step is
      // zero on the weight side, the tensor-memory columns beyond are written as zeros so no stale NaN pattern can be read
      uint32_t h[8] = {umma::pack2(vis2, rg.dirdiff[0]), umma::pack2(rg.dirdiff[1], rg.dirdiff[2]), umma::pack2(rg.dirdiff[3], 1.f),
                       umma::pack2(1.f, 0.f), 0u, 0u, 0u, 0u};
      umma::st8(tq + T_RGT, h);
    }
    // ------------------------------------------------------------ rgb_fc.0 tail, overlapped with view pooling #2
    SYNC_TMEM()      // also publishes XF and SF_VIS2 to the warpgroup (generic reads: no proxy fence)
    ISSUE_BEGIN()
      is.tmem_half(tb + C_RGB0, tb + T_RGT, W16(M_RGB0) + 4 * 16 * 16, 16, 1u);
    ISSUE_END()
    if (vrow < V) {   // weights = vis / (sum + 1e-8), mean / var of x over views, mean of weights -> the rays kernel's A operand
      const long long gs = (long long)tile * T + t;
      if (gs < p.total) {
        const unsigned rt = fastdiv((unsigned)gs, p.mv_magic, p.mv_shift), ri = (unsigned)gs - rt * (unsigned)p.Mv;
        unsigned char* dst = f2_op + (size_t)rt * kF2TileBytes + (size_t)ri * 16;
        float sum = 0.f;
#pragma unroll
        for (int vv = 0; vv < V; ++vv) sum += SF[SF_VIS2 * ROWS + vv * T + t];
        const float isum = 1.f / (sum + 1e-8f);
        float wv[V], ws = 0.f;
#pragma unroll
        for (int vv = 0; vv < V; ++vv) { wv[vv] = SF[SF_VIS2 * ROWS + vv * T + t] * isum; ws += wv[vv]; }
#pragma unroll 1
        for (int c = vrow; c < 4; c += V) {      // the V threads of a sample take every V-th chunk of 8 channels
          uint32_t mu[4], vr[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 xx[V];
#pragma unroll
            for (int vv = 0; vv < V; ++vv)
              xx[vv] = make_float2(XF[(8 * c + 2 * i) * ROWS + vv * T + t], XF[(8 * c + 2 * i + 1) * ROWS + vv * T + t]);
            float2 mean_c = fmul2(xx[0], make_float2(wv[0], wv[0]));
#pragma unroll
            for (int vv = 1; vv < V; ++vv) mean_c = ffma2(xx[vv], make_float2(wv[vv], wv[vv]), mean_c);
            const float2 nm = make_float2(-mean_c.x, -mean_c.y);
            float2 var_c = make_float2(0.f, 0.f);
#pragma unroll
            for (int vv = 0; vv < V; ++vv) { const float2 d = fadd2(xx[vv], nm); var_c = ffma2(make_float2(wv[vv], wv[vv]), fmul2(d, d), var_c); }
            mu[i] = umma::pack2(mean_c.x, mean_c.y);
            vr[i] = umma::pack2(var_c.x, var_c.y);
          }
          __stcs(reinterpret_cast<uint4*>(dst + (size_t)c * CH), make_uint4(mu[0], mu[1], mu[2], mu[3]));
          __stcs(reinterpret_cast<uint4*>(dst + (size_t)(4 + c) * CH), make_uint4(vr[0], vr[1], vr[2], vr[3]));
        }
        if (vrow == 0) __stcs(reinterpret_cast<uint4*>(dst + (size_t)8 * CH), make_uint4(umma::pack2(ws / (float)V, 0.f), 0u, 0u, 0u));
      }
    }
    WAIT_MMA()
    {   // rgb_fc.0 (ELU) -> rgb_fc.2 (16 -> 8, ELU) -> rgb_fc.4 (8 -> 1) as fp32 register GEMVs
      float y[16];
      umma::ld16(tq + C_RGB0, y);
      float2 acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 e0 = elu2p_f32(y[4 * i], y[4 * i + 1]), e1 = elu2p_f32(y[4 * i + 2], y[4 * i + 3]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(Wsm + SMW(M_RGB1) + j * 16 + 4 * i);
          acc[j] = ffma2(e0, make_float2(w.x, w.y), acc[j]);
          acc[j] = ffma2(e1, make_float2(w.z, w.w), acc[j]);
        }
      }
      float o = Wsm[SMB(M_RGB2)];
#pragma unroll
      for (int j = 0; j < 8; ++j) o = fmaf(elu_plain1(acc[j].x + acc[j].y + Wsm[SMB(M_RGB1) + j]), Wsm[SMW(M_RGB2) + j], o);
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
        const float id = 1.f / den;
        __stcs(f2_rgb + gs, make_float4(r * id, gg * id, b * id, 0.f));
      }
    }
    wg_sync(wg);   // SF / X / Y are rewritten by the next tile
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
    const int small_begin = sec16_begin(sec) + sec16_w_bytes(sec) + sec16_b_bytes(sec);
    *w_offset_bytes = small_begin + 4 * small16_offset(i);
    *b_offset_bytes = small_begin + 4 * small16_bias_offset(i);
  } else {
    *w_offset_bytes = sec16_begin(sec) + w16_offset(i);
    *b_offset_bytes = (kLayers16[i].bias == BIAS_CHUNK || kLayers16[i].bias == BIAS_F32) ? sec16_begin(sec) + bias16_offset(i) : -1;
  }
  for (int k = 0; k < kLayers16[i].Kpad; ++k) kmap[k] = w16_kmap(i, k);
  for (int n = 0; n < kLayers16[i].Npad; ++n) nmap[n] = w16_nmap(i, n);
  return PGRF_OK;
}
extern "C" int pgrf_w16_layer_info2(int i, int* bias_kind, int* in_ln2, int* out_log2e) {
  PGRF_REQUIRE(i >= 0 && i < kNumLayers16, "w16 layer index %d out of range", i);
  *bias_kind = kLayers16[i].bias; *in_ln2 = kLayers16[i].in_ln2; *out_log2e = kLayers16[i].out_log2e;
  return PGRF_OK;
}

namespace pgrf {
int g_mlp_wait_mode = 1;
int launch_render_mlp_bf16(const pgrf_render_args& a, int V, int T, long long total, int n_tiles, int Mv, int sms, cudaStream_t st) {
  PGRF_REQUIRE(a.weights16 != nullptr, "render: bf16 path needs weights16");
  PGRF_REQUIRE(total < (1ll << 31), "render: rn * dn = %lld samples per launch exceed 2^31 (lower rays_per_launch)", total);
  PGRF_REQUIRE(((uintptr_t)a.weights16 & 15) == 0, "render: weights16 must be 16-byte aligned");
  Render16Params p;
  p.a = a; p.V = V; p.T = T; p.M = V * T; p.total = total; p.n_tiles = n_tiles; p.Mv = Mv; p.wait_mode = g_mlp_wait_mode;
  PGRF_REQUIRE(a.dn >= 2 && Mv >= 2, "render: the bf16 path needs at least 2 samples per ray (dn=%d)", a.dn);
  fastdiv_gen((unsigned)a.dn, &p.dn_magic, &p.dn_shift);
  fastdiv_gen((unsigned)Mv, &p.mv_magic, &p.mv_shift);
  const int grid = min((n_tiles + kWG - 1) / kWG, sms);
  int dev = 0;
  PGRF_CUDA(cudaGetDevice(&dev));
#define PGRF_LAUNCH_V(VV)                                                                                              \
  case VV: {                                                                                                           \
    static bool done[64] = {};   /* the >48 KB opt-in is a per-device function attribute */                            \
    if (!done[dev & 63]) {                                                                                             \
      PGRF_CUDA(cudaFuncSetAttribute(render_mlp_bf16_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM16_BYTES)); \
      done[dev & 63] = true;                                                                                           \
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

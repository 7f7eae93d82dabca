// Weight layout of the bf16 tensor-core render path (fused rows+samples kernel csrc/render_mlp16.cu, rays kernel
// csrc/render_rays16.cu).
//
// Every Linear layer that runs on the tensor cores is stored as a tcgen05 B operand: bf16, k-chunk-major
// [Kpad/8][Npad][8] (see umma.cuh), K padded to a multiple of 8 (one chunk; an odd last chunk is paired with a shared zero
// chunk by the kernel, so no K=16 padding is stored) and N to a multiple of 16.
//
// Section 0 (fused MLP kernel):
//   * biases never touch the SIMT epilogues: a layer's bias is the extra K-step  A = [1, 1, 0 ...] (constant "ones" chunk)
//     x B = [bias_hi, bias_lo, 0 ...]  (bias split into two bf16 so the sum is fp32-accurate) — BIAS_CHUNK layers store that
//     B chunk ([Npad][8] bf16) right after the section's weights; BIAS_INLINE layers have a spare pair of K columns in a chunk
//     the kernel writes anyway (it writes the two ones itself) and carry the bias there (kmap = -2 / -3);
//   * ELU layers whose output only feeds another Linear layer are stored pre-scaled by log2(e) (`out_log2e`): the epilogue
//     evaluates  a' = log2e * ELU(x) = max(y, min(log2e * 2^y - log2e, 0)),  y = log2e * x  with one MUFU.EX2 and no multiply,
//     and the consuming layer's weights are stored multiplied by ln 2 (`in_ln2`);
//   * tiny output layers stay fp32 ([N][K] row-major + bias[N], W16_SMALL) and are evaluated as register GEMVs inside the
//     epilogue that produces their input.
//   Feature permutations:
//   XB  (input of base_fc.0, K = 232): [mean0 40 | var0 40 | mean1 40 | var1 40 | rgb_feat' 40 | prob_embedding 32]
//        where every 40-block is  f' = [img_feats 0..31, rgb 0..2, 5 zero]  (reference order: [rgb, img_feats]), so that
//        chunk c of a pooled block depends only on chunk c of the per-view features.
//   ray_dir_fc.2 output (N = 48): same f' order, so it can be added to the gathered features in place.
// Section 1 (rays kernel): plain K/N padding to 16, fp32 biases [Npad] after the section's weights.
// Python packs by asking pgrf_w16_layer_info / pgrf_w16_layer_info2 for the explicit k / n index maps and scales.
#pragma once

namespace pgrf {

enum W16Kind : int {
  W16_PLAIN = 0,     // kmap[k] = k < K ? k : -1 ; nmap[n] = n < N ? n : -1
  W16_BASE0 = 1,     // XB permutation (see above)
  W16_RD2 = 2,       // output permutation f'
  W16_SMALL = 3,     // tiny output layer kept in fp32 ([N][K] row-major + bias[N]), register GEMV in the producing epilogue
  W16_INLINE = 4,    // plain K order, then the two bias columns (kmap -2 = bias_hi, -3 = bias_lo) at k = K, K + 1
};
enum W16Bias : int { BIAS_F32 = 0, BIAS_CHUNK = 1, BIAS_INLINE = 2, BIAS_NONE = 3 };
constexpr int KMAP_ZERO = -1, KMAP_BIAS_HI = -2, KMAP_BIAS_LO = -3;

struct Layer16 {
  const char* name;
  int K, N;          // reference in/out features (of the whole torch weight)
  int Kpad, Npad;
  int kind;
  int section;       // 0 = fused rows+samples kernel, 1 = rays kernel
  int bias;          // W16Bias
  int in_ln2;        // weights stored * ln 2 (the producer's activation is stored * log2 e)
  int out_log2e;     // weights and bias stored * log2 e (the epilogue's ELU works on the pre-scaled value)
};

constexpr int kNumLayers16 = 30;
constexpr Layer16 kLayers16[kNumLayers16] = {
    {"{dd}.mean_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1}, {"{dd}.mean_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 1, 1},
    {"{dd}.mean_decoder.4", 32, 2, 32, 2, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{dd}.var_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1},  {"{dd}.var_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 1, 1},
    {"{dd}.var_decoder.4", 32, 2, 32, 2, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{dd}.aw_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1},   {"{dd}.aw_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 1, 1},
    {"{dd}.aw_decoder.4", 32, 1, 32, 1, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{dd}.vis_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1},  {"{dd}.vis_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 1, 1},
    {"{dd}.vis_decoder.4", 32, 1, 32, 1, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{agg}.prob_embed.0", 34, 32, 40, 32, W16_INLINE, 0, BIAS_INLINE, 0, 0},  {"{agg}.prob_embed.2", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 0},
    {"{agg}.agg_impl.ray_dir_fc.0", 4, 16, 8, 16, W16_INLINE, 0, BIAS_INLINE, 0, 1},
    {"{agg}.agg_impl.ray_dir_fc.2", 16, 35, 16, 48, W16_RD2, 0, BIAS_CHUNK, 1, 0},
    {"{agg}.agg_impl.neuray_fc.0", 32, 8, 32, 16, W16_PLAIN, 0, BIAS_CHUNK, 0, 0},
    {"{agg}.agg_impl.neuray_fc.2", 8, 1, 8, 1, W16_SMALL, 0, BIAS_F32, 0, 0},
    {"{agg}.agg_impl.base_fc.0", 207, 64, 232, 64, W16_BASE0, 0, BIAS_CHUNK, 0, 1},
    {"{agg}.agg_impl.base_fc.2", 64, 32, 64, 32, W16_PLAIN, 0, BIAS_CHUNK, 1, 0},
    {"{agg}.agg_impl.vis_fc.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1},
    {"{agg}.agg_impl.vis_fc.2", 32, 33, 32, 48, W16_PLAIN, 0, BIAS_CHUNK, 1, 0},
    {"{agg}.agg_impl.vis_fc2.0", 32, 32, 32, 32, W16_PLAIN, 0, BIAS_CHUNK, 0, 1},
    {"{agg}.agg_impl.vis_fc2.2", 32, 1, 32, 1, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{agg}.agg_impl.rgb_fc.0", 37, 16, 40, 16, W16_INLINE, 0, BIAS_INLINE, 0, 1},
    {"{agg}.agg_impl.rgb_fc.2", 16, 8, 16, 8, W16_SMALL, 0, BIAS_F32, 1, 0},
    {"{agg}.agg_impl.rgb_fc.4", 8, 1, 8, 1, W16_SMALL, 0, BIAS_F32, 0, 0},
    // ---- rays kernel ----
    {"{agg}.agg_impl.geometry_fc.0", 65, 64, 80, 64, W16_PLAIN, 1, BIAS_F32, 0, 0},
    {"{agg}.agg_impl.geometry_fc.2", 64, 16, 64, 16, W16_PLAIN, 1, BIAS_F32, 0, 0},
    {"{agg}.agg_impl.ray_attention.qkv", 16, 48, 16, 48, W16_PLAIN, 1, BIAS_F32, 0, 0},   // [w_qs | w_ks | w_vs] stacked along n, no bias
};
enum : int {
  M_MEAN0 = 0, M_MEAN1, M_MEAN2, M_VAR0, M_VAR1, M_VAR2, M_AW0, M_AW1, M_AW2, M_VIS0, M_VIS1, M_VIS2,
  M_PE0, M_PE1, M_RD0, M_RD1, M_NF0, M_NF1, M_BASE0, M_BASE1, M_VFC0, M_VFC1, M_V2_0, M_V2_1, M_RGB0, M_RGB1, M_RGB2,
  M_GEO0, M_GEO1, M_QKV,
};

// F2 hand-off (fused MLP kernel -> rays kernel), bf16 path: per tile of the RAYS kernel (whole rays, <= 128 samples) the pooled
// per-sample features as a ready tcgen05 A operand [9 k-chunks][128 rows][8] bf16 = [mean 32 | var 32 | mean-of-weights, 0 x7]
// (one bulk copy per tile), followed by one float4 (blended r, g, b, 0) per sample.
constexpr int kF2TileBytes = 9 * 128 * 16;
// samples per tile of the fused MLP kernel: 128 (view, sample) rows
constexpr int tile_samples16(int rfn) { return rfn >= 1 && rfn <= 4 ? 128 / rfn : 0; }

// feature order f' -> reference order of a 35-vector [rgb(3), img_feats(32)]
__host__ __device__ constexpr int fprime_to_ref(int f) { return f < 32 ? f + 3 : f - 32; }

constexpr bool is_small16(int j) { return kLayers16[j].kind == W16_SMALL; }
// section blob = [bf16 B operands of the MMA layers | bias region | fp32 small layers (W[N][K], b[N])]
//   bias region: section 0 = bf16 bias chunks [Npad][8] of the BIAS_CHUNK layers; section 1 = fp32 biases [Npad]
constexpr int sec16_w_bytes(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec && !is_small16(j)) o += kLayers16[j].Kpad * kLayers16[j].Npad * 2;
  return o;
}
constexpr int bias16_bytes_of(int j) {
  return is_small16(j) ? 0 : kLayers16[j].bias == BIAS_CHUNK ? kLayers16[j].Npad * 16 : kLayers16[j].bias == BIAS_F32 ? kLayers16[j].Npad * 4 : 0;
}
constexpr int sec16_b_bytes(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec) o += bias16_bytes_of(j);
  return o;
}
constexpr int small16_floats_of(int j) { return ((kLayers16[j].K * kLayers16[j].N + 3) & ~3) + ((kLayers16[j].N + 3) & ~3); }
constexpr int sec16_small_floats(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec && is_small16(j)) o += small16_floats_of(j);
  return o;
}
constexpr int sec16_bytes(int sec) { return sec16_w_bytes(sec) + sec16_b_bytes(sec) + 4 * sec16_small_floats(sec); }
constexpr int sec16_begin(int sec) { return sec == 0 ? 0 : sec16_bytes(0); }
// byte offset of MMA layer i's bf16 weights RELATIVE TO ITS SECTION (what the kernels index their smem copy with)
constexpr int w16_offset(int i) {
  int o = 0;
  for (int j = 0; j < i; ++j)
    if (kLayers16[j].section == kLayers16[i].section && !is_small16(j)) o += kLayers16[j].Kpad * kLayers16[j].Npad * 2;
  return o;
}
// byte offset of MMA layer i's bias (chunk or fp32 vector) relative to its section
constexpr int bias16_offset(int i) {
  int o = sec16_w_bytes(kLayers16[i].section);
  for (int j = 0; j < i; ++j)
    if (kLayers16[j].section == kLayers16[i].section) o += bias16_bytes_of(j);
  return o;
}
// float index of MMA layer i's fp32 bias inside its section's bias array (section 1)
constexpr int b16_offset(int i) { return (bias16_offset(i) - sec16_w_bytes(kLayers16[i].section)) / 4; }
// float index of small layer i's W[N][K] inside its section's small region (bias follows at + round4(K*N))
constexpr int small16_offset(int i) {
  int o = 0;
  for (int j = 0; j < i; ++j)
    if (kLayers16[j].section == kLayers16[i].section && is_small16(j)) o += small16_floats_of(j);
  return o;
}
constexpr int small16_bias_offset(int i) { return small16_offset(i) + ((kLayers16[i].K * kLayers16[i].N + 3) & ~3); }
constexpr int kW16Sec0Bytes = sec16_bytes(0);
constexpr int kW16WeightBytes = sec16_w_bytes(0);
constexpr int kW16SmallBegin0 = sec16_w_bytes(0) + sec16_b_bytes(0);   // byte offset of section 0's small region
constexpr int kW16Bytes = sec16_bytes(0) + sec16_bytes(1);
static_assert(sec16_w_bytes(0) % 16 == 0 && sec16_bytes(0) % 16 == 0 && sec16_w_bytes(1) % 16 == 0 && kW16SmallBegin0 % 16 == 0,
              "16-byte aligned regions");

inline int w16_kmap(int layer, int k) {
  const Layer16& L = kLayers16[layer];
  if (L.kind == W16_BASE0) {
    if (k < 160) return (k % 40) < 35 ? (k / 40) * 35 + fprime_to_ref(k % 40) : KMAP_ZERO;   // pooled blocks
    if (k < 200) return (k - 160) < 35 ? 140 + fprime_to_ref(k - 160) : KMAP_ZERO;           // rgb_feat'
    return 175 + (k - 200);                                                                  // prob_embedding
  }
  if (L.kind == W16_INLINE) return k < L.K ? k : k == L.K ? KMAP_BIAS_HI : k == L.K + 1 ? KMAP_BIAS_LO : KMAP_ZERO;
  return k < L.K ? k : KMAP_ZERO;
}
inline int w16_nmap(int layer, int n) {
  const Layer16& L = kLayers16[layer];
  if (L.kind == W16_RD2) return n < 35 ? fprime_to_ref(n) : -1;
  return n < L.N ? n : -1;
}

}  // namespace pgrf

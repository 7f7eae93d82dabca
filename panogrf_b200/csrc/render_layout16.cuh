// Weight layout of the bf16 tensor-core render path (fused rows+samples kernel, csrc/render_bf16.cu).
//
// Every Linear layer of dist_decoder / prob_embed / ray_dir_fc / neuray_fc / base_fc / vis_fc / vis_fc2 /
// rgb_fc is stored as a tcgen05 B operand: bf16, k-chunk-major [Kpad/8][Npad][8] (see umma.cuh), K padded
// to a multiple of 16 and N to a multiple of 16, followed (in a second region) by its fp32 bias [Npad].
// Some layers permute / pad their input or output features to make the operands chunk aligned:
//   XB  (input of base_fc.0, K = 240): [mean0 40 | var0 40 | mean1 40 | var1 40 | rgb_feat' 40 | neuray 32 | 8 zero]
//        where every 40-block is  f' = [img_feats 0..31, rgb 0..2, 5 zero]  (reference order: [rgb, img_feats]), so that
//        chunk c of a pooled block depends only on chunk c of the per-view features.
//   ray_dir_fc.2 output (N = 48): same f' order, so it can be added to the gathered features in place.
// Python packs by asking pgrf_w16_layer_info for the explicit k / n index maps (-1 = zero padding).
#pragma once

namespace pgrf {

enum W16Kind : int {
  W16_PLAIN = 0,     // kmap[k] = k < K ? k : -1 ; nmap[n] = n < N ? n : -1
  W16_BASE0 = 1,     // XB permutation (see above)
  W16_RD2 = 2,       // output permutation f'
  W16_SMALL = 3,     // tiny output layer kept in fp32 ([N][K] row-major + bias[N]) and evaluated as a register GEMV
                     // inside the previous layer's epilogue (no MMA stage, no bf16 rounding of its input)
};

struct Layer16 {
  const char* name;
  int K, N;          // reference in/out features (of the whole torch weight)
  int Kpad, Npad;
  int kind;
  int section;       // 0 = fused rows+samples kernel, 1 = rays kernel; blob = [W sec0 | bias sec0 | W sec1 | bias sec1]
};

constexpr int kNumLayers16 = 30;
constexpr Layer16 kLayers16[kNumLayers16] = {
    {"{dd}.mean_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0}, {"{dd}.mean_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{dd}.mean_decoder.4", 32, 2, 32, 2, W16_SMALL, 0},
    {"{dd}.var_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0},  {"{dd}.var_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{dd}.var_decoder.4", 32, 2, 32, 2, W16_SMALL, 0},
    {"{dd}.aw_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0},   {"{dd}.aw_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{dd}.aw_decoder.4", 32, 1, 32, 1, W16_SMALL, 0},
    {"{dd}.vis_decoder.0", 32, 32, 32, 32, W16_PLAIN, 0},  {"{dd}.vis_decoder.2", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{dd}.vis_decoder.4", 32, 1, 32, 1, W16_SMALL, 0},
    {"{agg}.prob_embed.0", 34, 32, 48, 32, W16_PLAIN, 0},  {"{agg}.prob_embed.2", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{agg}.agg_impl.ray_dir_fc.0", 4, 16, 4, 16, W16_SMALL, 0},
    {"{agg}.agg_impl.ray_dir_fc.2", 16, 35, 16, 48, W16_RD2, 0},
    {"{agg}.agg_impl.neuray_fc.0", 32, 8, 32, 8, W16_SMALL, 0},
    {"{agg}.agg_impl.neuray_fc.2", 8, 1, 8, 1, W16_SMALL, 0},
    {"{agg}.agg_impl.base_fc.0", 207, 64, 240, 64, W16_BASE0, 0},
    {"{agg}.agg_impl.base_fc.2", 64, 32, 64, 32, W16_PLAIN, 0},
    {"{agg}.agg_impl.vis_fc.0", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{agg}.agg_impl.vis_fc.2", 32, 33, 32, 48, W16_PLAIN, 0},
    {"{agg}.agg_impl.vis_fc2.0", 32, 32, 32, 32, W16_PLAIN, 0},
    {"{agg}.agg_impl.vis_fc2.2", 32, 1, 32, 1, W16_SMALL, 0},
    {"{agg}.agg_impl.rgb_fc.0", 37, 16, 48, 16, W16_PLAIN, 0},
    {"{agg}.agg_impl.rgb_fc.2", 16, 8, 16, 8, W16_SMALL, 0},
    {"{agg}.agg_impl.rgb_fc.4", 8, 1, 8, 1, W16_SMALL, 0},
    // ---- rays kernel ----
    {"{agg}.agg_impl.geometry_fc.0", 65, 64, 80, 64, W16_PLAIN, 1},
    {"{agg}.agg_impl.geometry_fc.2", 64, 16, 64, 16, W16_PLAIN, 1},
    {"{agg}.agg_impl.ray_attention.qkv", 16, 48, 16, 48, W16_PLAIN, 1},   // [w_qs | w_ks | w_vs] stacked along n, no bias
};
enum : int {
  M_MEAN0 = 0, M_MEAN1, M_MEAN2, M_VAR0, M_VAR1, M_VAR2, M_AW0, M_AW1, M_AW2, M_VIS0, M_VIS1, M_VIS2,
  M_PE0, M_PE1, M_RD0, M_RD1, M_NF0, M_NF1, M_BASE0, M_BASE1, M_VFC0, M_VFC1, M_V2_0, M_V2_1, M_RGB0, M_RGB1, M_RGB2,
  M_GEO0, M_GEO1, M_QKV,
};

// feature order f' -> reference order of a 35-vector [rgb(3), img_feats(32)]
__host__ __device__ constexpr int fprime_to_ref(int f) { return f < 32 ? f + 3 : f - 32; }

constexpr bool is_small16(int j) { return kLayers16[j].kind == W16_SMALL; }
// section blob = [bf16 B operands of the MMA layers | fp32 biases of the MMA layers | fp32 small layers (W[N][K], b[N])]
constexpr int sec16_w_bytes(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec && !is_small16(j)) o += kLayers16[j].Kpad * kLayers16[j].Npad * 2;
  return o;
}
constexpr int sec16_b_floats(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec && !is_small16(j)) o += kLayers16[j].Npad;
  return o;
}
constexpr int small16_floats_of(int j) { return ((kLayers16[j].K * kLayers16[j].N + 3) & ~3) + ((kLayers16[j].N + 3) & ~3); }
constexpr int sec16_small_floats(int sec) {
  int o = 0;
  for (int j = 0; j < kNumLayers16; ++j)
    if (kLayers16[j].section == sec && is_small16(j)) o += small16_floats_of(j);
  return o;
}
constexpr int sec16_bytes(int sec) { return sec16_w_bytes(sec) + 4 * sec16_b_floats(sec) + 4 * sec16_small_floats(sec); }
constexpr int sec16_begin(int sec) { return sec == 0 ? 0 : sec16_bytes(0); }
// byte offset of MMA layer i's bf16 weights RELATIVE TO ITS SECTION (what the kernels index their smem copy with)
constexpr int w16_offset(int i) {
  int o = 0;
  for (int j = 0; j < i; ++j)
    if (kLayers16[j].section == kLayers16[i].section && !is_small16(j)) o += kLayers16[j].Kpad * kLayers16[j].Npad * 2;
  return o;
}
// float index of MMA layer i's bias inside its section's bias array (which follows the section's weights)
constexpr int b16_offset(int i) {
  int o = 0;
  for (int j = 0; j < i; ++j)
    if (kLayers16[j].section == kLayers16[i].section && !is_small16(j)) o += kLayers16[j].Npad;
  return o;
}
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
constexpr int kW16SmallBegin0 = sec16_w_bytes(0) + 4 * sec16_b_floats(0);   // byte offset of section 0's small region
constexpr int kW16Bytes = sec16_bytes(0) + sec16_bytes(1);
static_assert(sec16_w_bytes(0) % 16 == 0 && sec16_bytes(0) % 16 == 0 && sec16_w_bytes(1) % 16 == 0 && kW16SmallBegin0 % 16 == 0,
              "16-byte aligned regions");

inline int w16_kmap(int layer, int k) {
  const Layer16& L = kLayers16[layer];
  if (L.kind == W16_BASE0) {
    if (k < 160) return (k % 40) < 35 ? (k / 40) * 35 + fprime_to_ref(k % 40) : -1;   // pooled blocks
    if (k < 200) return (k - 160) < 35 ? 140 + fprime_to_ref(k - 160) : -1;           // rgb_feat'
    if (k < 232) return 175 + (k - 200);                                              // neuray
    return -1;
  }
  return k < L.K ? k : -1;
}
inline int w16_nmap(int layer, int n) {
  const Layer16& L = kLayers16[layer];
  if (L.kind == W16_RD2) return n < 35 ? fprime_to_ref(n) : -1;
  return n < L.N ? n : -1;
}

}  // namespace pgrf

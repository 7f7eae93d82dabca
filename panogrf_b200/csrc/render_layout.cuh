// Weight-blob layout and intermediate row formats of the render-path kernels.
//
// One blob per network pair (dist_decoder + agg_net, or the fine_* pair).  Every Linear layer is
// stored k-major (Wt[k][n] = torch weight[n][k]) with n padded to a multiple of 4, followed by its
// bias (padded the same way), so that the kernels read weights with 128-bit shared-memory loads.
// Python packs by NAME through pgrf_weight_layer_info(); this table is the single source of truth.
#pragma once

namespace pgrf {

struct LayerDesc {
  const char* name;  // reference state_dict prefix relative to the renderer (without .weight/.bias)
  int K, N, Npad;
  int has_bias;
  int section;       // 0 = rows kernel (R1), 1 = samples kernel (R2), 2 = rays kernel (R3)
  int k_begin;       // first input column of the torch weight this entry covers (base_fc.0 is split)
};

// NOTE: names use '{dd}' for [fine_]dist_decoder and '{agg}' for [fine_]agg_net, resolved by the host.
constexpr int kNumLayers = 34;
constexpr LayerDesc kLayers[kNumLayers] = {
    // ---- section 0: per (view,sample) row networks ------------------------------------------
    {"{dd}.mean_decoder.0", 32, 32, 32, 1, 0, 0},
    {"{dd}.mean_decoder.2", 32, 32, 32, 1, 0, 0},
    {"{dd}.mean_decoder.4", 32, 2, 4, 1, 0, 0},
    {"{dd}.var_decoder.0", 32, 32, 32, 1, 0, 0},
    {"{dd}.var_decoder.2", 32, 32, 32, 1, 0, 0},
    {"{dd}.var_decoder.4", 32, 2, 4, 1, 0, 0},
    {"{dd}.aw_decoder.0", 32, 32, 32, 1, 0, 0},
    {"{dd}.aw_decoder.2", 32, 32, 32, 1, 0, 0},
    {"{dd}.aw_decoder.4", 32, 1, 4, 1, 0, 0},
    {"{dd}.vis_decoder.0", 32, 32, 32, 1, 0, 0},   // zeros when use_vis is false
    {"{dd}.vis_decoder.2", 32, 32, 32, 1, 0, 0},
    {"{dd}.vis_decoder.4", 32, 1, 4, 1, 0, 0},
    {"{agg}.prob_embed.0", 34, 32, 32, 1, 0, 0},
    {"{agg}.prob_embed.2", 32, 32, 32, 1, 0, 0},
    {"{agg}.agg_impl.ray_dir_fc.0", 4, 16, 16, 1, 0, 0},
    {"{agg}.agg_impl.ray_dir_fc.2", 16, 35, 36, 1, 0, 0},
    {"{agg}.agg_impl.neuray_fc.0", 32, 8, 8, 1, 0, 0},
    {"{agg}.agg_impl.neuray_fc.2", 8, 1, 4, 1, 0, 0},
    // ---- section 1: per-sample cross-view network --------------------------------------------
    {"{agg}.agg_impl.base_fc.0", 140, 64, 64, 0, 1, 0},     // columns 0..139: [mean0,var0,mean1,var1]
    {"{agg}.agg_impl.base_fc.0", 67, 64, 64, 1, 1, 140},    // columns 140..206: [rgb_feat, neuray_feat]
    {"{agg}.agg_impl.base_fc.2", 64, 32, 32, 1, 1, 0},
    {"{agg}.agg_impl.vis_fc.0", 32, 32, 32, 1, 1, 0},
    {"{agg}.agg_impl.vis_fc.2", 32, 33, 36, 1, 1, 0},
    {"{agg}.agg_impl.vis_fc2.0", 32, 32, 32, 1, 1, 0},
    {"{agg}.agg_impl.vis_fc2.2", 32, 1, 4, 1, 1, 0},
    {"{agg}.agg_impl.rgb_fc.0", 37, 16, 16, 1, 1, 0},
    {"{agg}.agg_impl.rgb_fc.2", 16, 8, 8, 1, 1, 0},
    {"{agg}.agg_impl.rgb_fc.4", 8, 1, 4, 1, 1, 0},
    // ---- section 2: per-ray geometry head + ray transformer ----------------------------------
    {"{agg}.agg_impl.geometry_fc.0", 65, 64, 64, 1, 2, 0},
    {"{agg}.agg_impl.geometry_fc.2", 64, 16, 16, 1, 2, 0},
    {"{agg}.agg_impl.ray_attention.qkv", 16, 48, 48, 0, 2, 0},  // [w_qs | w_ks | w_vs] concatenated along n
    {"{agg}.agg_impl.ray_attention.fc", 16, 16, 16, 0, 2, 0},
    {"{agg}.agg_impl.out_geometry_fc.0", 16, 16, 16, 1, 2, 0},
    {"{agg}.agg_impl.out_geometry_fc.2", 16, 1, 4, 1, 2, 0},
};

constexpr int layer_floats(int i) { return kLayers[i].K * kLayers[i].Npad + (kLayers[i].has_bias ? kLayers[i].Npad : 0); }
constexpr int layer_offset(int i) {  // offset of layer i inside the blob (floats)
  int o = 0;
  for (int j = 0; j < i; ++j) o += layer_floats(j);
  return o;
}
constexpr int section_begin(int s) {
  int o = 0;
  for (int j = 0; j < kNumLayers; ++j) {
    if (kLayers[j].section == s) return o;
    o += layer_floats(j);
  }
  return o;
}
constexpr int section_floats(int s) {
  int o = 0;
  for (int j = 0; j < kNumLayers; ++j)
    if (kLayers[j].section == s) o += layer_floats(j);
  return o;
}
// layer_norm weight/bias (16+16) are appended to section 2, then the positional table [dn][16]
constexpr int kLnOffset = layer_offset(kNumLayers - 1) + layer_floats(kNumLayers - 1);
constexpr int kPosencOffset = kLnOffset + 32;
constexpr int kMaxSamplesPerRay = 128;
constexpr int kBlobFloats = kPosencOffset + kMaxSamplesPerRay * 16;

// Offset of layer i relative to the start of ITS section (what the kernels index smem with)
constexpr int sec_off(int i) { return layer_offset(i) - section_begin(kLayers[i].section); }
constexpr int bias_off(int i) { return sec_off(i) + kLayers[i].K * kLayers[i].Npad; }

// layer indices
enum : int {
  L_MEAN0 = 0, L_MEAN1, L_MEAN2, L_VAR0, L_VAR1, L_VAR2, L_AW0, L_AW1, L_AW2, L_VIS0, L_VIS1, L_VIS2,
  L_PE0, L_PE1, L_RD0, L_RD1, L_NF0, L_NF1,
  L_BASE0G, L_BASE0R, L_BASE1, L_VFC0, L_VFC1, L_VFC2_0, L_VFC2_1, L_RGB0, L_RGB1, L_RGB2,
  L_GEO0, L_GEO1, L_QKV, L_AFC, L_OG0, L_OG1,
};

// ---- intermediate formats (HBM) --------------------------------------------------------------
// F1: rows kernel -> samples kernel.  One block per tile of T samples: [kF1][kTileRows] floats,
// column m = v*T + t (view-major), feature-major so a block is one contiguous TMA bulk copy.
constexpr int kTileRows = 128;
constexpr int kF1 = 76;
constexpr int F1_RGBFEAT = 0;    // 35: [rgb(3), img_feats(32)] + ray_dir_fc output
constexpr int F1_NEURAY = 35;    // 32: prob_embed output
constexpr int F1_W0 = 67;        // sigmoid(neuray_fc)
constexpr int F1_DIRDIFF = 68;   // 4
constexpr int F1_RGBRAW = 72;    // 3 (+1 pad)
// F2: samples kernel -> rays kernel.  Per tile of T samples: [kF2][T] floats.
constexpr int kF2 = 68;          // mean(32), var(32), mean-of-weights(1), rgb_out(3)
constexpr int F2_RGB = 65;

// samples per tile for V source views: T*V <= 128, T a power of two <= 64
__host__ __device__ constexpr int tile_samples(int V) { return V <= 2 ? 64 : (V <= 4 ? 32 : (V <= 8 ? 16 : 0)); }

}  // namespace pgrf

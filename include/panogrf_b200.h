/*
 * panogrf_b200 — C ABI of the B200-native PanoGRF render-time hot path.
 *
 * The reference (thucz/PanoGRF) is pure Python/PyTorch and has no FFI layer of its own; its
 * boundary for this path is the Python function / nn.Module API.  This header is
 * the boundary a maintainer binds to replace that path: plain pointers and sizes, no torch types.
 * `panogrf_b200/_lib.py` is the ctypes binding; `INTEGRATION.md` shows the reference-side stub.
 *
 * Conventions
 *   - every entry point returns 0 on success, a negative PGRF_E* code on failure;
 *     `pgrf_last_error()` returns a thread-local human-readable message for the last failure.
 *   - `*_fwd` entry points take DEVICE pointers (fp32, contiguous, 16-byte aligned) and a
 *     `cudaStream_t` passed as `void*` (NULL = legacy default stream); they only enqueue work.
 *   - `*_host` entry points take HOST pointers, do H2D, the kernels and D2H themselves and
 *     synchronise before returning (this is what `bench.py`'s `e2e` number times).
 *   - no entry point ever falls back to a CPU implementation.
 */
#ifndef PANOGRF_B200_H_
#define PANOGRF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PGRF_API __attribute__((visibility("default")))
#else
#define PGRF_API
#endif

#define PGRF_OK 0
#define PGRF_EINVAL (-1)   /* bad shape / unsupported configuration */
#define PGRF_ECUDA (-2)    /* CUDA runtime error (message in pgrf_last_error) */
#define PGRF_ERANGE (-3)   /* device-side range assertion ("Wrong UV mapping") */

/* dataset_name -> id (helpers/my_torch_helpers.py:33-58, network/spt_utils.py:45-86) */
#define PGRF_DS_M3D 0
#define PGRF_DS_REPLICA_TEST 1
#define PGRF_DS_RESIDENTIAL 2
#define PGRF_DS_COFFEEAREA 3

/* cost_type -> id (models/spherical_cost_volume.py:212-217) */
#define PGRF_COST_ABS_DIFF 0
#define PGRF_COST_DOT 1
#define PGRF_COST_NONE 2

/* physical layout of the cost volume written by pgrf_cost_volume_fwd */
#define PGRF_CV_BDCHW 0   /* storage of the reference's stack(dim=1): (B,D,C,H,W)  (:340)            */
#define PGRF_CV_BDHWC 1   /* channels-last contiguous (B,D,H,W,C)                                   */
#define PGRF_CV_BCDHW 2   /* what the 3-D regulariser consumes (pipeline3_model.py:847); with        */
                          /* groups>0 this is the group-wise mean (B,G,D,H,W) (pipeline3_model.py:849-853) */
#define PGRF_CV_BDHWC_BF16 3 /* channels-last (B,D,H,W,C) stored as bf16 (`out` points to 2-byte elements, C % 16 == 0): the */
                          /* operand layout of pgrf_conv3d_fwd — the regulariser consumes the sweep without a conversion pass */

PGRF_API const char* pgrf_last_error(void);
PGRF_API int pgrf_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
PGRF_API int64_t pgrf_launch_count(void);

/*
 * K1 — fused spherical-sweep cost volume.
 * Replaces models/spherical_cost_volume.py:231-341 (calculate_cost_volume_erp),
 *          models/spherical_cost_volume.py:135-230 (get_cv_per_depth),
 *          models/spherical_cost_volume_mv.py:219-347 (calculate_cost_volume_erp_multiview).
 *
 *   images        (B,S,H,W,C) channels-last; C in {4,8,16,32,64}
 *   depths        (D) scalar hypotheses, used when depth_volume == NULL
 *   depth_volume  (B,D,H,W) per-pixel hypotheses (args["contain_dnet"]) or NULL
 *   rots (B,S,3,3), trans (B,S,3)   world->camera of every view
 *   ref_idx       reference view (1 for the 2-view call, curr_idx for multi-view)
 *   src_views[n_src]  HOST array of the views swept and summed (n_src <= 8)
 *   divisor       each per-view cost is divided by this before summation (seq_len-2 for MV, 0 = none)
 *   out           layout per `layout`/`groups`; err_flag: device int, OR-ed with 1 when a uv falls
 *                 outside [-1,1] (the reference's assert at :191); caller checks it once.
 */
PGRF_API int pgrf_cost_volume_fwd(const float* images, int B, int S, int H, int W, int C,
                         const float* depths, const float* depth_volume, int D,
                         const float* rots, const float* trans,
                         int ref_idx, const int* src_views, int n_src, float divisor,
                         int dataset, int cost_type, int layout, int groups,
                         float* out, int* err_flag, void* stream);

/* tcgen05/TMEM self-test: out[128,N] = bf16(A[128,K]) * bf16(W[N,K])^T, fp32 accumulate (device pointers) */
PGRF_API int pgrf_umma_selftest(const float* A, const float* W, float* out, int K, int N, int variant, void* stream);

/* tuning knobs for experiments ("cv_jb": gathers batched per lane 2|4|8, "cv_dchunk": depths per CTA, 0 = heuristic) */
PGRF_API int pgrf_debug_set(const char* key, int value);

/* Same computation from/to HOST buffers (H2D + kernel + D2H + sync). Returns PGRF_ERANGE if flagged. */
PGRF_API int pgrf_cost_volume_host(const float* images, int B, int S, int H, int W, int C,
                          const float* depths, const float* depth_volume, int D,
                          const float* rots, const float* trans,
                          int ref_idx, const int* src_views, int n_src, float divisor,
                          int dataset, int cost_type, int layout, int groups,
                          float* out);

/* K1 backward: gradient w.r.t. the feature maps (the reference: autograd through grid_sample + abs / mul,
 * models/spherical_cost_volume.py:135-230; hypotheses and poses carry no gradient, pipeline3_model.py:647,671).
 *   grad_out     (B,D,H,W,C) contiguous upstream gradient of the cost volume
 *   grad_images  (B,S,H,W,C) ACCUMULATED into with vector atomics: zero it before the call
 * Other arguments as pgrf_cost_volume_fwd. */
PGRF_API int pgrf_cost_volume_bwd(const float* grad_out, const float* images, int B, int S, int H, int W, int C,
                                  const float* depths, const float* depth_volume, int D, const float* rots, const float* trans,
                                  int ref_idx, const int* src_views, int n_src, float divisor, int dataset, int cost_type,
                                  float* grad_images, void* stream);


/* ------------------------------------------------------------------------------------------------
 * K2 + K3 + K4 — one render pass (coarse or fine) over a batch of rays.
 * Replaces, per ray batch, network/renderer.py:223-317 (render_by_depth) and everything it calls:
 *   render_ops.py:76-106 (depth2points_spherical), :110-122 (depth2inv_dists), :158-257
 *   (project_points_dict), ops.py:32-52 (interpolate_feats), renderer.py:120-136
 *   (predict_proj_ray_prob) + dist_decoder.py:99-140, renderer.py:180-188 (get_img_feats),
 *   aggregate_net.py:41-89 + ibrnet.py:315-373, renderer.py:210-219 (network_rendering),
 *   render_ops.py:145-153 (alpha_values2hit_prob), renderer.py:302-304 (render_depth) and, when
 *   `fine_depth` is set, render_ops.py:413-473 (sample_fine_depth) + the sort of renderer.py:470-472.
 * All pointers are DEVICE pointers to contiguous fp32 unless stated.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pgrf_render_args {
  int dataset;                 /* PGRF_DS_* (cfg["dataset_name"]) */
  int H, W;                    /* ERP size of the spherical convention (cfg["height"], cfg["width"]) */
  int rfn;                     /* source views, 1..4 */
  int rn;                      /* rays in this batch */
  int dn;                      /* samples per ray, 3..128 */
  int use_vis;                 /* dist_decoder_cfg.use_vis of the COARSE decoder (renderer.py:129) */
  float bias_val;              /* dist_decoder_cfg.bias_val (0.05) */
  const float* coords;         /* (rn,2) query pixel (x,y); truncated like .long() */
  const float* depth;          /* (rn,dn) sample depths, or (dn) shared by all rays when depth_ray_stride==0 */
  int depth_ray_stride;        /* dn or 0 */
  const float* que_c2w;        /* (3,4) query camera-to-world */
  float que_near, que_far;     /* que_imgs_info["depth_range"] */
  const float* ref_w2c;        /* (rfn,3,4) */
  const float* ref_depth_range;/* (rfn,2) */
  const float* imgs_cl;        /* (rfn,img_h,img_w,4) channels-last rgb, 4th channel padding */
  int img_h, img_w;
  const float* img_feats_cl;   /* (rfn,if_h,if_w,32) channels-last */
  int if_h, if_w;
  const float* ray_feats_cl;   /* (rfn,rf_h,rf_w,32) channels-last */
  int rf_h, rf_w;
  const float* weights;        /* packed blob, pgrf_weight_blob_floats() floats, see pgrf_weight_layer_info */
  float* f1;                   /* workspace, sizes from pgrf_render_workspace */
  float* f2;
  float* pixel_colors;         /* (rn,3)  out */
  float* render_depth;         /* (rn)    out, optional */
  float* hit_prob;             /* (rn,dn) out, optional */
  float* density;              /* (rn,dn) out, optional */
  float* colors;               /* (rn,dn,3) out, optional */
  float* fine_depth;           /* (rn, fine_dn [+ dn]) out, optional: sorted fine samples */
  int fine_dn;
  const float* fine_u;         /* (fine_dn) the deterministic u table of render_ops.py:442-445 */
  int fine_use_all;            /* cfg["fine_depth_use_all"] */
  int use_disp;                /* cfg["use_disp"] (inv_mode of sample_fine_depth) */
  int* fine_inds;              /* (rn,fine_dn) searchsorted bin indices, optional (parity tests) */
  float* prob_dbg;             /* (rfn,rn*dn,3) alpha, vis, hit_prob of prj_dict, optional */
  float* prj_dbg;              /* (rfn,rn*dn,6) pts(2), depth, dir(3) of prj_dict, optional */
  float* feat_dbg;             /* (rfn,rn*dn,67) ray_feats(32), rgb(3), img_feats(32) of prj_dict, optional */
  int stage_mask;              /* 0 = all kernels; else bit0 rows, bit1 samples, bit2 rays (profiling; bf16: bit0|bit1 = fused MLP kernel) */
  int mlp_bf16;                /* 1 = bf16 tcgen05 MLP path (rtol 1e-2), 0 = fp32 SIMT parity path (rtol 1e-4) */
  int* sched;                  /* optional device int[2]: dynamic tile counters of the bf16 kernels (zeroed by the call) */
  const void* weights16;       /* bf16 blob (pgrf_w16_blob_bytes bytes, layout from pgrf_w16_layer_info); needed when mlp_bf16 */
  /* ---- optional per-row INPUTS replacing the fused producers (fp32 path only; for callers of the reference's module-level
   * API: predict_proj_ray_prob renderer.py:120-136, DefaultAggregationNet.forward aggregate_net.py:41-89, network_rendering
   * renderer.py:210-219).  Same record layouts as the *_dbg outputs above, (rfn, rn*dn, .) ---- */
  const float* prj_in;         /* px,py,depth,dir[3]: skips ray construction + projection; needs que_dir_in and interval_in;
                                  coords / que_c2w / ref_w2c may then be NULL */
  const float* feat_in;        /* ray_feats[32] rgb[3] img_feats[32]: skips the three gathers; the maps may then be NULL */
  const float* prob_in;        /* alpha, vis, hit_prob: skips the dist decoder + compute_prob */
  const float* que_dir_in;     /* (rn*dn,3) unit query ray directions */
  const float* interval_in;    /* (rn*dn) que_dists of depth2inv_dists (render_ops.py:110-122); `depth` may then be NULL
                                  (no render_depth / fine sampling) */
  float* dec_dbg;              /* optional OUTPUT (rfn, rn*dn, 6): mean[2], var[2], vis, aw of the dist decoder */
  /* ablation switches of DefaultAggregationNet (network/aggregate_net.py:60-62, 79-81) */
  int wo_geometry;             /* agg_net_cfg.wo_geometry: prob_embedding = 0 */
  int wo_appearance;           /* agg_net_cfg.wo_appearance: [rgb, img_feats] of every view = 0 (the blended colours are then 0) */
  /* perspective / cube query rays (is_perspec, network/render_ops.py:37-74): optional (rn,3) WORLD-space ray directions
   * (unnormalised, the reference's `directions`) replacing the ERP pixel -> ray table; the origin stays que_c2w[:,3] (= -R^T t);
   * `coords` is then unused (may be NULL) */
  const float* ray_dirs;
} pgrf_render_args;
/* Module-level entry ("agg_mlp_fwd"): the aggregation network + ray transformer + compositing on
 * caller-provided per-row inputs (prj_in, feat_in, prob_in, que_dir_in required).  fp32. */
PGRF_API int pgrf_agg_mlp_fwd(const pgrf_render_args* args, void* stream);

PGRF_API int pgrf_render_pass_fwd(const pgrf_render_args* args, void* stream);
/* workspace sizes (floats) for `n_samples` = rn*dn samples and rfn views */
PGRF_API int pgrf_render_workspace(int rfn, long long n_samples, long long* f1_floats, long long* f2_floats);

/* Whole view = the reference's ray-batch loop (network/renderer.py:647-683) + coarse->fine hand-off
 * (renderer.py:600-631, 435-524).  `pass` describes the COARSE pass over ALL rays of the view
 * (depth = the shared (dn) table, depth_ray_stride = 0, outputs sized for pass.rn rays, f1/f2 sized
 * for rays_per_launch rays of max(dn, fine_total) samples); rays are processed rays_per_launch at a time. */
typedef struct pgrf_render_view_args {
  pgrf_render_args pass;
  int hierarchical;            /* cfg["use_hierarchical_sampling"] */
  const float* weights_fine;   /* blob of fine_dist_decoder + fine_agg_net (== pass.weights for cfg["one_mlp"]) */
  const void* weights16_fine;  /* bf16 blob of the fine nets (when pass.mlp_bf16) */
  float bias_val_fine;
  int rays_per_launch;
  float* fine_depth_ws;        /* (rays_per_launch, fine_total) workspace, used when que_depth_fine == NULL */
  float* pixel_colors_fine;    /* (rn,3) */
  float* render_depth_fine;    /* (rn) optional */
  float* hit_prob_fine;        /* (rn,fine_total) optional */
  float* density_fine;         /* (rn,fine_total) optional */
  float* colors_fine;          /* (rn,fine_total,3) optional */
  float* que_depth_fine;       /* (rn,fine_total) optional: the sorted fine sample depths */
} pgrf_render_view_args;

PGRF_API int pgrf_render_view_fwd(const pgrf_render_view_args* args, void* stream);
/* HOST-pointer variant: all pointers are host memory, the three maps are NCHW like the reference's
 * ref_imgs_info (imgs: 3 channels); does H2D, layout conversion, all passes, D2H and synchronises. */
PGRF_API int pgrf_render_view_host(const pgrf_render_view_args* args);
/* (N,C,H,W) -> channels-last (N,H,W,Cpad), zero padded */
PGRF_API int pgrf_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int Cpad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone operators (the reference's functional API, network/render_ops.py); device pointers.
 * ---------------------------------------------------------------------------------------------- */
/* project_points_dict (render_ops.py:234-257) + get_img_feats (renderer.py:180-188): pts (pn,3) world points ->
 * per view: pixel (rfn,pn,2), depth (rfn,pn), dir (rfn,pn,3), ray_feats (rfn,pn,32), rgb (rfn,pn,3), img_feats (rfn,pn,32, optional).
 * Maps channels-last like pgrf_render_args. */
PGRF_API int pgrf_project_gather_fwd(const float* pts, long long pn, const float* w2c, int rfn, int dataset, int H, int W,
                                     const float* imgs_cl, int img_h, int img_w, const float* img_feats_cl, int if_h, int if_w,
                                     const float* ray_feats_cl, int rf_h, int rf_w, float* out_pix, float* out_depth,
                                     float* out_dir, float* out_ray_feats, float* out_rgb, float* out_img_feats, void* stream);
/* alpha_values2hit_prob (render_ops.py:145-153) [+ renderer.py:214-218,302-304]: pass `alpha` (rn,dn), or `density` (rn,dn) for
 * alpha = 1-exp(-relu(density)); optional colors (rn,dn,3) -> pixel_colors (rn,3); optional depth -> render_depth (rn). */
PGRF_API int pgrf_composite_fwd(const float* density, const float* alpha, const float* colors, const float* depth,
                                int depth_ray_stride, int rn, int dn, float* hit_prob, float* pixel_colors, float* render_depth,
                                void* stream);
/* sample_fine_depth (render_ops.py:413-473) with the deterministic u table; fine_out (rn, fine_dn [+dn if use_all]);
 * sort_out = 1 also applies the sort of renderer.py:470-472; inds_out (rn,fine_dn) optional searchsorted indices. */
PGRF_API int pgrf_fine_sample_fwd(const float* depth, int depth_ray_stride, const float* hit_prob, const float* u_table,
                                  float near_depth, float far_depth, int inv_mode, int rn, int dn, int fine_dn, int sort_out,
                                  int use_all, float* fine_out, int* inds_out, void* stream);
/* sample_3sigma + sample_pdf (network/sample_utils.py:6-60, det = True) and its use in fine_render_impl (network/renderer.py:438-470).
 * select = 1: ft (rn, ft_stride >= 3) rows [marker, low, high]; rays with marker >= min_valid get io[ray] = sort(n samples
 * [++ coarse_depth[ray] (dn) when coarse_depth != NULL]), the other rows of io (rn, n [+ dn]) are left as they are.
 * select = 0: ft rows [low, high]; io (rn, n) = the unsorted samples of every ray (the functional form).
 * t_table (n) = linspace(0,1,n), gauss (n-1) = N(0,1) density at linspace(-3,3,n-1), both built by the host with torch. */
PGRF_API int pgrf_sample_3sigma_fwd(const float* ft, int ft_stride, float min_valid, const float* t_table, const float* gauss, int n,
                                    float near_depth, float far_depth, const float* coarse_depth, int coarse_ray_stride, int dn,
                                    int select, int rn, float* io, void* stream);
/* depth hypotheses of the MVS net (pipeline3_model.py:723-733,774-815): out (B,n_mono+n_linear,h,w) = per-pixel sorted
 * [clamp(ref_mu + k_sigma[i], min, max)] ++ linear[]; k_sigma and linear ascending (device arrays, computed by the host
 * with the reference's own ops so they are bit-identical). */
PGRF_API int pgrf_depth_hypotheses_fwd(const float* ref_mu, int B, int h, int w, const float* k_sigma, int n_mono,
                                       const float* linear, int n_linear, float min_depth, float max_depth, float* out, void* stream);
/* every variant of the builder (pipeline3_model.py:717-733, 774-815):
 *   mono list  clamp(ref_mu + s * k_i): mono_mode 0 = k_table holds float32(k_i * fixed_sigma); 1 = s = max(ref_sigma, basic_sigma)
 *              (mono_uncertainty, :731); 2 = (ref_sigma * k_i) * relaxation (:729); n_mono <= 16, 0 = none (n_samples == 0)
 *   centres    centers_mode 0 = table `centers` (n_centers; linear :802 or inverse-linear :804, built by the host with the reference's
 *              torch ops); 1 = per-pixel `revise_range` with `fixed_dist` (:784-799); 2 = none (`wo_hdh`)
 *   sort_out   1: per-pixel ascending merge (:815); 0: mono list in k order (the `wo_hdh` branch does not sort)
 * out (B, n_mono + n_centers, h, w). */
PGRF_API int pgrf_depth_hypotheses2_fwd(const float* ref_mu, const float* ref_sigma, int B, int h, int w, const float* k_table,
                                        int n_mono, int mono_mode, float basic_sigma, float relaxation, const float* centers,
                                        int n_centers, int centers_mode, float fixed_dist, float min_depth, float max_depth,
                                        int sort_out, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Depth-prior sample placement ("diner" branch of render_impl, network/renderer.py:570-600, :318-355):
 *   sample_depth(n_candidates, linear)            network/render_ops.py:292-339
 *   project_points_dict_diner                     network/render_ops.py:260-290
 *   sample_depthguided / fill_up_uniform_samples  network/original_depth_guided_sample.py:45-297 / 333-366
 * fused into one kernel: every candidate depth of a ray is projected into every source panorama, the MVS depth /
 * variance / normal priors are gathered (bilinear, border), the surface likelihood is taken as the max over views, the
 * n_samples - n_gaussian most likely candidates are kept (ties: lower candidate first), n_gaussian samples are drawn
 * around the occlusion-aware mean, empty slots are filled uniformly, the optional uniform samples are appended, and the
 * result is sorted.  The reference's two random draws are explicit inputs (fill_rand, gauss).  Device pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pgrf_diner_args {
  int dataset, H, W;             /* ERP convention, cfg height / width */
  int rfn;
  long long rn;
  int n_candidates, n_samples, n_gaussian, n_uniform;
  int include_norm;              /* cfg backface_culling: keep candidates with dot(R_v * ray, normal) <= 0 only */
  int sigma_is_var;              /* 1: sigma = sqrt(uncert) (reference default var=True), 0: sigma = uncert */
  float diner_sigma;             /* > 0: constant sigma (cfg diner_sigma) */
  float cand_step;               /* fp32((max_depth - min_depth) / n_candidates), computed by the host */
  float min_depth, max_depth, depth_diff_max;
  const float* coords;           /* (rn,2) pixel (x,y) */
  const float* cand_depth;       /* (n_candidates) shared by all rays (cand_ray_stride = 0) or (rn,n_candidates) */
  long long cand_ray_stride;
  const float* que_c2w;          /* (3,4) */
  const float* ref_w2c;          /* (rfn,3,4) */
  const float* mvs_depth;        /* (rfn,1,map_h,map_w) NCHW, as ref_imgs_info['mvs_depth'] */
  const float* mvs_uncert;       /* (rfn,1,map_h,map_w) */
  const float* mvs_normal;       /* (rfn,3,map_h,map_w) or NULL when include_norm = 0 */
  int map_h, map_w, img_h, img_w;
  const float* fill_rand;        /* (rn,n_samples) U[0,1) */
  const float* gauss;            /* (rn,n_gaussian) N(0,1), NULL when n_gaussian = 0 */
  const float* uniform_depth;    /* (n_uniform) appended to every ray (cfg contain_uniform), NULL when n_uniform = 0 */
  float* out_depth;              /* (rn, n_samples + n_uniform), ascending */
  float* likelihood;             /* optional (rn,n_candidates): max-over-views likelihood of every candidate */
  /* dict variant (sample_depthguided called on a precomputed project_points_dict_diner result); all NULL for the fused path */
  const float* prj_mu;           /* (rfn,rn,n_candidates) */
  const float* prj_uncert;       /* (rfn,rn,n_candidates) */
  const float* prj_depth;        /* (rfn,rn,n_candidates) */
  const float* prj_normal;       /* (rfn,rn,n_candidates,3) */
  const float* que_dir;          /* (rn,n_candidates,3) */
} pgrf_diner_args;
PGRF_API int pgrf_depth_guided_sample_fwd(const pgrf_diner_args* args, void* stream);
/* project_points_dict_diner (render_ops.py:260-290): pts (pn,3) -> pixel (rfn,pn,2), depth (rfn,pn), gathered mvs depth
 * (rfn,pn), variance (rfn,pn), normal (rfn,pn,3; optional) */
PGRF_API int pgrf_project_gather_diner_fwd(const float* pts, long long pn, const float* w2c, int rfn, int dataset, int H, int W,
                                           const float* mvs_depth, const float* mvs_uncert, const float* mvs_normal, int map_h,
                                           int map_w, int img_h, int img_w, float* out_pix, float* out_depth, float* out_mu,
                                           float* out_uncert, float* out_normal, void* stream);

/* depth2normal (network/orig_diner_depth2normal.py:7-110): prior normals of the depth-guided placement (cfg backface_culling):
 * mvs_depth (N,1,H,W) at the ERP size of the spherical convention -> normal (N,3,H,W); zero rows above / below, longitude wrap,
 * the reference's hole "cleaning" (lookup shifted away from neighbours whose x coordinate is 0), zero normal where depth == 0.
 * raw_ws: N*H*W*3 floats, off_ws: N*H*W*2 bytes of workspace. */
PGRF_API int pgrf_depth2normal_fwd(const float* mvs_depth, int N, int H, int W, int dataset, float* raw_ws, signed char* off_ws,
                                   float* out_normal, void* stream);

/* Equirectangular -> cubemap resampling of the MVS input path: replaces Equirec2Cube.run (UniFuse datasets/util.py:74-100) as called
 * by e2c_process (network/omni_mvsnet/pipeline3_model.py:262-283; scipy map_coordinates order 1, mode 'wrap', on the CPU).
 * equ (n_img,H,W,C) channels-last panoramas; coor_x / coor_y (face_w, 6*face_w) the reference's sampling coordinates (float32, built by
 * the host with the reference's numpy expressions); cube (n_img, face_w, 6*face_w, C) in [F R B L U D] order. */
PGRF_API int pgrf_e2c_fwd(const float* equ, int n_img, int H, int W, int C, const float* coor_x, const float* coor_y, int face_w,
                          float* cube, void* stream);

/* 3-D cost regulariser: the Conv3DBlockv2 / UNet2 stack (models/common_blocks.py:187-242, 366-503) applied to the cost
 * volume by network/omni_mvsnet/pipeline3_model.py:847-855.  Activations are bf16 CHANNELS-LAST (B,D,H,W,C) with C padded to a
 * multiple of 16; weights are packed by the host (panogrf_b200/regulariser.py: pack_conv) per (output-channel tile, tap, channel chunk).
 *  - pgrf_conv3d_to_bf16_cl: fp32 (B,C,D,H,W) with arbitrary ELEMENT strides (so every cost-volume layout feeds it without a copy)
 *    -> bf16 channels-last with Cpad channels (zeros above C);
 *  - pgrf_conv3d_fwd: Conv3d(k=3) over WrapPadding3D(1) (zeros along D and H, wrap along W) + bias (+ LeakyReLU 0.01 if act) of
 *    the channel concatenation [xa (Ca) | xb (Cb, may be NULL/0)]; tcgen05 implicit GEMM.  Exactly one output: y = bf16 channels-last
 *    with Cout channels, or yf = the first cout_real (of Cout padded) channels as fp32 planar (B,cout_real,D,H,W) — the
 *    single-channel head of the last decoder keeps fp32.  Layers whose grid cannot fill the GPU split K over the tap rows and need
 *    ws: pgrf_conv3d_workspace gives the number of floats (0 = none) for the same arguments;
 *  - pgrf_conv3d_cout1_fwd: the same convolution with ONE output channel on the fp32 pipes (SIMT), fp32 output (B,D,H,W); input either the bf16 pair or a
 *    single-channel fp32 volume xf; w is fp32 [27][Cin] (tap-major);
 *  - pgrf_avgpool3d2_fwd: AvgPool3d(2); pgrf_upsample3d2_fwd: F.interpolate(scale_factor=2, mode='trilinear') (align_corners False). */
PGRF_API int pgrf_conv3d_to_bf16_cl(const float* x, long long sb, long long sc, long long sd, long long sh, long long sw, int B, int C,
                                    int D, int H, int W, int Cpad, void* y, void* stream);
PGRF_API int pgrf_conv3d_workspace(int Ca, int Cb, int Cout, int B, int D, int H, int W, long long* ws_floats);
PGRF_API int pgrf_conv3d_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, float* yf,
                             int cout_real, int Cout, int B, int D, int H, int W, int act, float* ws, long long ws_floats, void* stream);
/* pgrf_conv3d_fwd with an optional residual input `res` (bf16 channels-last, the shape of y; added after bias / activation:
 * ResidualBlock, network/ops.py:61-115) and the choice of the width padding: wrap = 1 WrapPadding, 0 zeros on every side. */
PGRF_API int pgrf_conv3d_ex_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, float* yf,
                                int cout_real, int Cout, int B, int D, int H, int W, int act, float* ws, long long ws_floats,
                                const void* res, int wrap, void* stream);
/* DefaultVisEncoder (network/vis_encoder.py:6-33) around the convolutions: pgrf_feats_to_bf16_cl = cat(F.interpolate(img_feats, (h,w),
 * 'bilinear'), ray_feats) of two fp32 NCHW maps -> bf16 channels-last (N,h,w,Ci+Cr); pgrf_instnorm_relu_fwd = nn.InstanceNorm2d
 * (affine, biased variance, eps) + ReLU on a bf16 channels-last map (N,HW,C); stats_ws: 2*N*C doubles. */
PGRF_API int pgrf_feats_to_bf16_cl(const float* img, int Ci, int hi, int wi, const float* ray, int Cr, int N, int h, int w, void* out,
                                   void* stream);
PGRF_API int pgrf_instnorm_relu_fwd(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                                    void* y, void* stream);
/* Parts of the image encoder ResUNetLight (network/ops.py:126-455): pgrf_instnorm_act_fwd = act(InstanceNorm2d(x) [+ res]) with act
 * 0 none / 1 ReLU / 2 ELU (BasicBlock: relu(bn2(conv2) + identity); conv module: elu(bn(conv))), C up to 256;
 * pgrf_patch7x7_s2_fwd = the 7x7xCin input patch (WrapPadding(3) or zeros, stride 2) of every output pixel of conv1 as a bf16
 * channels-last row (k = c*49 + ky*7 + kx, padded to Kpad), which turns the 7x7 convolution into a pointwise GEMM;
 * pgrf_subsample2_fwd = y[yo][xo] = x[2yo][2xo] (stride 2 of a stride-1 result / input of a 1x1 stride-2 convolution);
 * pgrf_upsample2d2_ac_fwd = F.interpolate(scale_factor=2, 'bilinear', align_corners=True) (upconv).  All on bf16 channels-last. */
PGRF_API int pgrf_instnorm_act_fwd(const void* x, int N, int HW, int C, const float* gamma, const float* beta, float eps, double* stats_ws,
                                   const void* res, int act, void* y, void* stream);
PGRF_API int pgrf_patch7x7_s2_fwd(const float* x, int N, int Cin, int H, int W, int Kpad, int wrap, void* out, void* stream);
PGRF_API int pgrf_subsample2_fwd(const void* x, int N, int h, int w, int C, void* y, void* stream);
PGRF_API int pgrf_upsample2d2_ac_fwd(const void* x, int N, int h, int w, int C, void* y, void* stream);
/* A 3x3x3 convolution with ONE output channel over many input channels (the 128 -> 1 head of the last decoder) in two steps that read
 * every voxel's channels once instead of once per tap row: pgrf_conv3d_pointwise_fwd = 1x1x1 convolution through the tensor-core
 * pipeline (only the centre tap of the packed weights is walked; here with the 27 taps as output channels, fp32 planar
 * (B,27,D,H,W) output), then pgrf_conv3d_tapsum_fwd = out[v] = act(bias + sum_tap z[tap][neighbour(v, tap)]) (zeros along D / H,
 * wrap along W). */
PGRF_API int pgrf_conv3d_pointwise_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y,
                                       float* yf, int cout_real, int Cout, int B, int D, int H, int W, int act, void* stream);
PGRF_API int pgrf_conv3d_tapsum_fwd(const float* z, float bias, int B, int D, int H, int W, int act, float* out, void* stream);
/* 1 -> 1 channel 3x3x3 convolution of an fp32 scalar volume (B,D,H,W); w27_host = the 27 taps (kd, kh, kw order) in HOST memory. */
PGRF_API int pgrf_conv3d_scalar_fwd(const float* x, const float* w27_host, float bias, int B, int D, int H, int W, int act, float* out,
                                    void* stream);
PGRF_API int pgrf_conv3d_cout1_fwd(const void* xa, int Ca, const void* xb, int Cb, const float* xf, const float* w, float bias, int B,
                                   int D, int H, int W, int act, float* out, void* stream);
PGRF_API int pgrf_avgpool3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream);
PGRF_API int pgrf_upsample3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream);
/* 2-D heads after the regulariser (models/test_models.py:147-205, pipeline3_model.py:866-905).  A (B,1,H,W,C) volume is a 2-D feature
 * map: pgrf_conv3d_fwd with D == 1 is Conv2d(3x3) over WrapPadding (zeros along H, wrap along W; common_blocks.py:258-293) and walks
 * only the three kd == 1 tap rows.  pgrf_upsample2d2_fwd: F.interpolate(scale_factor=2, 'bilinear', align_corners=False) on bf16
 * channels-last (Upscale, common_blocks.py:245-256).  pgrf_channel_dot_upsample_fwd: decoders1 = 1x1 convolution over the C channels
 * of a strided fp32 (B,C,H,W) map (+ bias), bilinear x`scale` upsampling and the depth rectification (0 none, 1 clamp(min=0),
 * 2 1/(clamp(min=0)+1e-10)) -> (B, H*scale, W*scale) fp32. */
PGRF_API int pgrf_upsample2d2_fwd(const void* x, int B, int H, int W, int C, void* y, void* stream);
PGRF_API int pgrf_channel_dot_upsample_fwd(const float* x, long long sb, long long sc, long long sh, long long sw, int B, int C, int H,
                                           int W, const float* w, float bias, int scale, int rectify, float* out, void* stream);

/* MixtureLogisticsDistDecoder.compute_prob (dist_decoder.py:113-140) with get_near_far_points(is_ref=True) (:6-51):
 * depth (rfn,n), interval (n) shared by all views or (rfn,n) when interval_per_view, mean/var (rfn,n,2), vis (rfn,n) or NULL
 * (use_vis=False), aw (rfn,n), depth_range (rfn,2); n = rays*dn -> alpha, visibility, hit_prob (rfn,n) */
PGRF_API int pgrf_compute_prob_fwd(const float* depth, const float* interval, int interval_per_view, const float* mean,
                                   const float* var, const float* vis, const float* aw, const float* depth_range, int rfn,
                                   long long n, int dn, float* alpha, float* visibility, float* hit_prob, void* stream);
/* compute_prob with is_ref=False (dist_decoder.py:37-45,109-140): the query rays' own hit probability (training: depth loss).  depth,
 * interval (qn,n); mean / var (qn,n,2), vis (qn,n) or NULL, aw (qn,n) already broadcast over the dn samples of a ray; depth_range (qn,2). */
PGRF_API int pgrf_compute_prob_que_fwd(const float* depth, const float* interval, const float* mean, const float* var, const float* vis,
                                       const float* aw, const float* depth_range, int qn, long long n, int dn, float* alpha,
                                       float* visibility, float* hit_prob, void* stream);
/* interpolate_feature_map (render_ops.py:126-143 -> ops.py:32-52): feats (rfn,C,fh,fw) NCHW, pix (rfn,pn,2) in full-res
 * (h,w) pixel units -> out (rfn,pn,C); bilinear, padding 'border', align_corners = (fh==h && fw==w) */
PGRF_API int pgrf_interpolate_feature_map_fwd(const float* feats, int rfn, int C, int fh, int fw, const float* pix, long long pn,
                                              int h, int w, float* out, void* stream);
/* depth2points_spherical (render_ops.py:76-106): coords (rn,2), depth (rn,dn) or (dn) shared (stride 0), c2w (3,4) ->
 * pts (rn,dn,3) world points, dir (rn,dn,3) = -ray/|ray| */
PGRF_API int pgrf_depth2points_fwd(const float* coords, const float* depth, int depth_ray_stride, const float* c2w, int dataset,
                                   int H, int W, long long rn, int dn, float* pts, float* dir, void* stream);

/* Backward of pgrf_composite_fwd (the reference: autograd through cumprod / exp): upstream gradients g_hit_prob (rn,dn),
 * g_pixel_colors (rn,3), g_render_depth (rn) (each optional) -> grad w.r.t. density (or alpha, whichever was the input) (rn,dn)
 * and colors (rn,dn,3; optional) */
PGRF_API int pgrf_composite_bwd(const float* density, const float* alpha, const float* colors, const float* depth, int depth_ray_stride,
                                int rn, int dn, const float* g_hit_prob, const float* g_pixel_colors, const float* g_render_depth,
                                float* grad_density_or_alpha, float* grad_colors, void* stream);
/* Backward of pgrf_interpolate_feature_map_fwd w.r.t. the map: grad_feats (rfn,C,fh,fw) is ACCUMULATED into (zero it first) */
PGRF_API int pgrf_interpolate_feature_map_bwd(const float* grad_out, int rfn, int C, int fh, int fw, const float* pix, long long pn,
                                              int h, int w, float* grad_feats, void* stream);

/* Weight blob layout: one entry per (slice of a) Linear layer of [fine_]dist_decoder / [fine_]agg_net.
 * name uses "{dd}" / "{agg}" placeholders; the weight slice [N, k_begin:k_begin+K] is stored
 * transposed (k-major) with rows padded to Npad at w_offset, the bias (Npad) at b_offset. */
PGRF_API int pgrf_weight_blob_floats(void);
PGRF_API int pgrf_weight_num_layers(void);
PGRF_API int pgrf_weight_layer_info(int i, char* name, int name_cap, int* K, int* N, int* Npad, int* has_bias,
                                    int* k_begin, int* w_offset, int* b_offset);
/* bf16 tensor-core blob: layer i stored as [Kpad/8][Npad][8] bf16 at w_offset_bytes, its bias at b_offset_bytes (see layer_info2);
 * kmap[Kpad] / nmap[Npad] give the reference input / output feature of every padded slot (-1 = zero) */
PGRF_API int pgrf_w16_blob_bytes(void);
PGRF_API int pgrf_w16_num_layers(void);
PGRF_API int pgrf_w16_layer_info(int i, char* name, int name_cap, int* Kpad, int* Npad, int* w_offset_bytes, int* b_offset_bytes,
                                 int* kmap, int* nmap, int* is_small);
/* is_small = 1: tiny output layer kept as fp32 W[N][K] row-major at w_offset_bytes and bias[N] at b_offset_bytes (Kpad = K, Npad = N) */
/* kmap values: >= 0 reference input feature, -1 zero column, -2 / -3 the layer's bias as a bf16 (hi, lo) pair (bias_kind 2).
 * bias_kind: 0 = fp32 [Npad] at b_offset_bytes, 1 = bf16 chunk [Npad][8] = (hi, lo, 0 x6) at b_offset_bytes (B operand of a
 * K-step against a constant ones chunk), 2 = inside the weight (kmap -2 / -3; b_offset_bytes = -1), 3 = none.
 * in_ln2: store the weights multiplied by ln 2; out_log2e: store weights and bias multiplied by log2(e) (the kernel evaluates
 * the ELU between two such layers on pre-scaled values). */
PGRF_API int pgrf_w16_layer_info2(int i, int* bias_kind, int* in_ln2, int* out_log2e);
/* layer-norm (weight 16, bias 16) offset, positional table offset ([max_samples][16]) */
PGRF_API int pgrf_weight_aux_offsets(int* layer_norm_offset, int* posenc_offset, int* max_samples);

#ifdef __cplusplus
}
#endif
#endif /* PANOGRF_B200_H_ */

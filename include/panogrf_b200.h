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

/* Same computation from/to HOST buffers (H2D + kernel + D2H + sync). Returns PGRF_ERANGE if flagged. */
PGRF_API int pgrf_cost_volume_host(const float* images, int B, int S, int H, int W, int C,
                          const float* depths, const float* depth_volume, int D,
                          const float* rots, const float* trans,
                          int ref_idx, const int* src_views, int n_src, float divisor,
                          int dataset, int cost_type, int layout, int groups,
                          float* out);

#ifdef __cplusplus
}
#endif
#endif /* PANOGRF_B200_H_ */

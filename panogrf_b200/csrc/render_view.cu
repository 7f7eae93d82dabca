// Whole-view driver of the render path: the reference's ray-batch loop (network/renderer.py:647-683)
// and coarse -> fine hand-off (renderer.py:600-631, 435-524) as a C-ABI call, plus the layout
// conversion of the source maps and the HOST-buffer entry point used for end-to-end timing.
#include <vector>

#include "common.cuh"

namespace pgrf {

// (N,C,H,W) -> (N,H,W,Cpad) through a 32x33 shared-memory tile: coalesced on both sides.
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int C, long long HW, int Cpad) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long pix = p0 + tx;
    tile[i][tx] = (c < C && pix < HW) ? __ldg(src + ((size_t)n * C + c) * HW + pix) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const long long pix = p0 + i;
    const int c = c0 + tx;
    if (pix < HW && c < Cpad) dst[((size_t)n * HW + pix) * Cpad + c] = tile[tx][i];
  }
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_nchw_to_nhwc(const float* src, float* dst, int N, int C, int H, int W, int Cpad, void* stream) {
  PGRF_REQUIRE(src && dst && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C, "nchw_to_nhwc: bad arguments");
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), (unsigned)N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, C, HW, Cpad);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_render_view_fwd(const pgrf_render_view_args* va, void* stream) {
  PGRF_REQUIRE(va != nullptr, "render_view: null args");
  const pgrf_render_args& base = va->pass;
  PGRF_REQUIRE(va->rays_per_launch >= 1, "render_view: rays_per_launch=%d", va->rays_per_launch);
  PGRF_REQUIRE(base.depth_ray_stride == 0 || base.depth_ray_stride == base.dn, "render_view: depth_ray_stride must be 0 or dn");
  const int dn = base.dn;
  const int fine_total = va->hierarchical ? base.fine_dn + (base.fine_use_all ? dn : 0) : 0;
  if (va->hierarchical) {
    PGRF_REQUIRE(!base.mlp_bf16 || va->weights16_fine, "render_view: bf16 fine pass needs weights16_fine");
    PGRF_REQUIRE(va->weights_fine && va->pixel_colors_fine && base.fine_u, "render_view: fine pass needs weights_fine, "
                 "pixel_colors_fine and fine_u");
    PGRF_REQUIRE(va->que_depth_fine || va->fine_depth_ws, "render_view: need que_depth_fine or fine_depth_ws");
  }
  for (int r0 = 0; r0 < base.rn; r0 += va->rays_per_launch) {
    const int n = base.rn - r0 < va->rays_per_launch ? base.rn - r0 : va->rays_per_launch;
    pgrf_render_args c = base;
    c.rn = n;
    c.coords = base.coords ? base.coords + 2 * (size_t)r0 : nullptr;
    c.ray_dirs = base.ray_dirs ? base.ray_dirs + 3 * (size_t)r0 : nullptr;
    c.depth = base.depth + (size_t)r0 * base.depth_ray_stride;
    c.pixel_colors = base.pixel_colors + 3 * (size_t)r0;
    if (base.render_depth) c.render_depth = base.render_depth + r0;
    if (base.hit_prob) c.hit_prob = base.hit_prob + (size_t)r0 * dn;
    if (base.density) c.density = base.density + (size_t)r0 * dn;
    if (base.colors) c.colors = base.colors + (size_t)r0 * dn * 3;
    c.prob_dbg = nullptr; c.prj_dbg = nullptr; c.feat_dbg = nullptr; c.fine_inds = nullptr;
    float* fine_depth = nullptr;
    if (va->hierarchical)
      fine_depth = va->que_depth_fine ? va->que_depth_fine + (size_t)r0 * fine_total : va->fine_depth_ws;
    c.fine_depth = fine_depth;
    int rc = pgrf_render_pass_fwd(&c, stream);
    if (rc != PGRF_OK) return rc;
    if (!va->hierarchical) continue;
    pgrf_render_args f = c;
    f.dn = fine_total;
    f.depth = fine_depth;
    f.depth_ray_stride = fine_total;
    f.weights = va->weights_fine;
    f.weights16 = va->weights16_fine;
    f.bias_val = va->bias_val_fine;
    f.fine_depth = nullptr;
    f.pixel_colors = va->pixel_colors_fine + 3 * (size_t)r0;
    f.render_depth = va->render_depth_fine ? va->render_depth_fine + r0 : nullptr;
    f.hit_prob = va->hit_prob_fine ? va->hit_prob_fine + (size_t)r0 * fine_total : nullptr;
    f.density = va->density_fine ? va->density_fine + (size_t)r0 * fine_total : nullptr;
    f.colors = va->colors_fine ? va->colors_fine + (size_t)r0 * fine_total * 3 : nullptr;
    rc = pgrf_render_pass_fwd(&f, stream);
    if (rc != PGRF_OK) return rc;
  }
  return PGRF_OK;
}

// HOST-buffer variant: every pointer of `hv` (inputs, weights, outputs) is a HOST pointer, the
// three source maps are in the reference's NCHW layout (imgs has 3 channels).  Workspace pointers
// (f1, f2, fine_depth_ws) are ignored and allocated here.
extern "C" int pgrf_render_view_host(const pgrf_render_view_args* hv) {
  PGRF_REQUIRE(hv != nullptr, "render_view_host: null args");
  const pgrf_render_args& h = hv->pass;
  PGRF_REQUIRE(h.coords && h.depth && h.que_c2w && h.ref_w2c && h.ref_depth_range && h.imgs_cl && h.img_feats_cl &&
                   h.ray_feats_cl && h.weights && h.pixel_colors,
               "render_view_host: null pointer argument");
  const int rfn = h.rfn, rn = h.rn, dn = h.dn;
  const int fine_total = hv->hierarchical ? h.fine_dn + (h.fine_use_all ? dn : 0) : 0;
  const int chunk = hv->rays_per_launch < rn ? hv->rays_per_launch : rn;
  long long f1n = 0, f2n = 0, f1b = 0, f2b = 0;
  int rc = pgrf_render_workspace(rfn, (long long)chunk * dn, &f1n, &f2n);
  if (rc != PGRF_OK) return rc;
  if (hv->hierarchical) {
    rc = pgrf_render_workspace(rfn, (long long)chunk * fine_total, &f1b, &f2b);
    if (rc != PGRF_OK) return rc;
    if (f1b > f1n) f1n = f1b;
    if (f2b > f2n) f2n = f2b;
  }
  cudaStream_t st;
  PGRF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  {  // keep freed blocks in the stream-ordered pool between calls (default threshold 0 returns them to the OS at every sync)
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  std::vector<void*> allocs;
  auto dalloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMallocAsync(&p, bytes ? bytes : 16, st) != cudaSuccess) return nullptr;
    allocs.push_back(p);
    return p;
  };
  auto up = [&](const void* src, size_t bytes) -> float* {
    void* d = dalloc(bytes);
    if (d && cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    return (float*)d;
  };
  const int blob = pgrf_weight_blob_floats();
  auto body = [&]() -> int {
    pgrf_render_view_args d = *hv;
    pgrf_render_args& p = d.pass;
    const size_t n_img = (size_t)rfn * 3 * h.img_h * h.img_w, n_if = (size_t)rfn * 32 * h.if_h * h.if_w,
                 n_rf = (size_t)rfn * 32 * h.rf_h * h.rf_w;
    p.coords = up(h.coords, (size_t)rn * 2 * 4);
    p.depth = up(h.depth, (h.depth_ray_stride ? (size_t)rn * dn : (size_t)dn) * 4);
    p.que_c2w = up(h.que_c2w, 12 * 4);
    p.ref_w2c = up(h.ref_w2c, (size_t)rfn * 12 * 4);
    p.ref_depth_range = up(h.ref_depth_range, (size_t)rfn * 2 * 4);
    p.weights = up(h.weights, (size_t)blob * 4);
    if (h.mlp_bf16) p.weights16 = up(h.weights16, (size_t)pgrf_w16_blob_bytes());
    float* imgs_nchw = up(h.imgs_cl, n_img * 4);
    float* if_nchw = up(h.img_feats_cl, n_if * 4);
    float* rf_nchw = up(h.ray_feats_cl, n_rf * 4);
    float* imgs_cl = (float*)dalloc((size_t)rfn * h.img_h * h.img_w * 4 * 4);
    float* if_cl = (float*)dalloc(n_if * 4);
    float* rf_cl = (float*)dalloc(n_rf * 4);
    if (hv->hierarchical) {
      d.weights_fine = up(hv->weights_fine, (size_t)blob * 4);
      if (h.mlp_bf16) d.weights16_fine = up(hv->weights16_fine, (size_t)pgrf_w16_blob_bytes());
      p.fine_u = up(h.fine_u, (size_t)h.fine_dn * 4);
      d.fine_depth_ws = (float*)dalloc((size_t)chunk * fine_total * 4);
      d.que_depth_fine = nullptr;
    }
    p.f1 = h.mlp_bf16 ? nullptr : (float*)dalloc((size_t)f1n * 4);
    p.sched = h.mlp_bf16 ? (int*)dalloc(16) : nullptr;
    p.f2 = (float*)dalloc((size_t)f2n * 4);
    // device outputs for whatever the caller asked for
    struct Out { float** dev; float* host; size_t n; };
    std::vector<Out> outs;
    auto want = [&](float** slot, float* host_ptr, size_t n) {
      *slot = nullptr;
      if (host_ptr) { *slot = (float*)dalloc(n * 4); outs.push_back({slot, host_ptr, n}); }
    };
    want(&p.pixel_colors, h.pixel_colors, (size_t)rn * 3);
    want(&p.render_depth, h.render_depth, (size_t)rn);
    want(&p.hit_prob, h.hit_prob, (size_t)rn * dn);
    want(&p.density, h.density, (size_t)rn * dn);
    want(&p.colors, h.colors, (size_t)rn * dn * 3);
    want(&d.pixel_colors_fine, hv->pixel_colors_fine, (size_t)rn * 3);
    want(&d.render_depth_fine, hv->render_depth_fine, (size_t)rn);
    want(&d.hit_prob_fine, hv->hit_prob_fine, (size_t)rn * fine_total);
    want(&d.density_fine, hv->density_fine, (size_t)rn * fine_total);
    want(&d.colors_fine, hv->colors_fine, (size_t)rn * fine_total * 3);
    for (void* al : allocs) PGRF_REQUIRE(al != nullptr, "render_view_host: device allocation / upload failed");
    p.imgs_cl = imgs_cl; p.img_feats_cl = if_cl; p.ray_feats_cl = rf_cl;
    p.prob_dbg = nullptr; p.prj_dbg = nullptr; p.feat_dbg = nullptr; p.fine_inds = nullptr; p.fine_depth = nullptr;
    int r = pgrf_nchw_to_nhwc(imgs_nchw, imgs_cl, rfn, 3, h.img_h, h.img_w, 4, st);
    if (r == PGRF_OK) r = pgrf_nchw_to_nhwc(if_nchw, if_cl, rfn, 32, h.if_h, h.if_w, 32, st);
    if (r == PGRF_OK) r = pgrf_nchw_to_nhwc(rf_nchw, rf_cl, rfn, 32, h.rf_h, h.rf_w, 32, st);
    if (r == PGRF_OK) r = pgrf_render_view_fwd(&d, st);
    if (r != PGRF_OK) return r;
    for (const Out& o : outs) PGRF_CUDA(cudaMemcpyAsync(o.host, *o.dev, o.n * 4, cudaMemcpyDeviceToHost, st));
    PGRF_CUDA(cudaStreamSynchronize(st));
    return PGRF_OK;
  };
  rc = body();
  for (void* a : allocs) if (a) cudaFreeAsync(a, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return rc;
}

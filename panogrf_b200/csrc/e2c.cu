// Equirectangular -> cubemap resampling on the GPU (first slice).
//
// Replaces Equirec2Cube.run (UniFuse-Unidirectional-Fusion/UniFuse/datasets/util.py:74-100) as called per batch item and view by
// e2c_process (network/omni_mvsnet/pipeline3_model.py:262-283): there every panorama is copied to the host, resampled channel by
// channel with scipy's map_coordinates(order=1, mode='wrap') and copied back.  Here: one launch for all (batch, view) images,
// thread = (image, cube pixel), all channels.
//   * sampling coordinates (coor_x, coor_y) per cube pixel are a table built once per size by the host with the reference's own
//     numpy expressions (float32), so they are bit-identical;
//   * the reference appends two rows below the image before sampling (last row and first row, each rolled by W/2: the other side
//     of the pole) — rows H and H+1 are synthesised on the fly;
//   * 'wrap' of scipy's legacy mode: period n-1 on the COORDINATE (first and last sample coincide), then linear interpolation in
//     double precision between floor(c) and floor(c)+1 — kept in fp64 (the op is 0.1 M pixels) so the result equals scipy's.
#include "common.cuh"

namespace pgrf {

__device__ __forceinline__ double wrap_coord(double c, int n) {
  const double sz = (double)(n - 1);
  if (c < 0.0) c += sz * (trunc(-c / sz) + 1.0);
  else if (c > sz) c -= sz * trunc(c / sz);
  return c;
}

__global__ void __launch_bounds__(256) e2c_kernel(const float* __restrict__ equ, const float* __restrict__ coor_x,
                                                  const float* __restrict__ coor_y, int n_img, int H, int W, int C, int face_w,
                                                  float* __restrict__ cube) {
  const long long npix = (long long)face_w * face_w * 6;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * n_img) return;
  const long long img = i / npix, p = i % npix;
  const float* e = equ + img * (long long)H * W * C;
  const int Hp = H + 2;
  const double cy = wrap_coord((double)__ldg(coor_y + p), Hp), cx = wrap_coord((double)__ldg(coor_x + p), W);
  const int y0 = (int)floor(cy), x0 = (int)floor(cx);
  const double ty = cy - y0, tx = cx - x0;
  const int y1 = min(y0 + 1, Hp - 1), x1 = min(x0 + 1, W - 1);
  // padded row r >= H: H = last row rolled by W/2, H+1 = first row rolled by W/2  (np.roll(row, W/2): out[x] = row[(x - W/2) mod W])
  auto at = [&](int r, int x, int c) -> double {
    if (r >= H) { x = (x - W / 2 + W) % W; r = (r == H) ? H - 1 : 0; }
    return (double)__ldg(e + ((long long)r * W + x) * C + c);
  };
  float* o = cube + (img * npix + p) * C;
  for (int c = 0; c < C; ++c) {
    const double v = at(y0, x0, c) * (1.0 - ty) * (1.0 - tx) + at(y0, x1, c) * (1.0 - ty) * tx + at(y1, x0, c) * ty * (1.0 - tx) +
                     at(y1, x1, c) * ty * tx;
    o[c] = (float)v;
  }
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_e2c_fwd(const float* equ, int n_img, int H, int W, int C, const float* coor_x, const float* coor_y, int face_w,
                            float* cube, void* stream) {
  PGRF_REQUIRE(equ && coor_x && coor_y && cube, "e2c: null pointer argument");
  PGRF_REQUIRE(n_img >= 1 && H >= 2 && W >= 2 && C >= 1 && face_w >= 1, "e2c: bad sizes n=%d H=%d W=%d C=%d face=%d", n_img, H, W, C, face_w);
  const long long total = (long long)face_w * face_w * 6 * n_img;
  e2c_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(equ, coor_x, coor_y, n_img, H, W, C, face_w, cube);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

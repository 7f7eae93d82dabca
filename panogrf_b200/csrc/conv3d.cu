// 3-D cost regulariser: the Conv3DBlockv2 / UNet2 stack of models/common_blocks.py:187-242, 366-503 as consumed by
// network/omni_mvsnet/pipeline3_model.py:847-855, on the 5th-generation tensor cores.
//
//   conv3d_igemm_kernel / conv3d_igemm_row_kernel
//        3x3x3 convolution (+ bias + LeakyReLU 0.01) as an implicit GEMM on tcgen05: M = 128 output voxels per CTA (voxel <-> TMEM
//        lane), N = a tile of output channels (16..128, fp32 accumulators in tensor memory), K = 27 taps x input channels, walked in
//        pipeline stages of one tap (per-tap variant) or one (kd, kh) tap row = three taps (row variant) of a <=64-channel chunk.
//        A operand: activations are bf16 CHANNELS-LAST (B,D,H,W,C), so a voxel's channel chunk is one contiguous run: eight gather
//        warps copy the runs with 16-byte cp.async straight into the k-chunk-major stage buffer — WrapPadding3D (common_blocks.py:
//        448-503: zeros along depth / height, wrap along width) is folded into the addresses / zero-fill — and hand completion to the
//        stage's mbarrier (cp.async.mbarrier.arrive), so nobody waits for its own loads.  B operand: the layer's weights pre-packed
//        per (n-tile, tap, chunk) as ready k-chunk-major blocks, one bulk copy (TMA) each.  A ninth warp issues the MMAs and commits
//        them to the stage's `free` barrier.  A second input pointer makes the U-Net's torch.cat((upsampled, skip), 1) free; layers
//        whose grid cannot fill the GPU split K over the nine tap rows (conv3d_reduce_kernel sums the partial volumes).
//   conv3d_cout1_kernel  single-output-channel layer on the fp32 pipes (the 1 -> 1 head of the last decoder), fp32 output.
//   avgpool / trilinear  AvgPool3d(2) and the x2 trilinear upsampling (align_corners=False) of UNet2.forward on bf16 channels-last.
// Numerics: bf16 operands, fp32 accumulation and activation (the reference's cuDNN convolutions run in TF32 on the same GPUs).
#include <cuda_bf16.h>

#include "render_device.cuh"
#include "umma.cuh"

namespace pgrf {

constexpr int kConvRows = 128;          // output voxels per CTA == TMEM lanes
constexpr int kGatherThreads = 256;     // 8 gather / epilogue warps
constexpr int kConvThreads = 288;       // + 1 MMA-issuing warp
// Bytes appended to each k-chunk plane of the A stage.  The gather writes 16 bytes per lane with 8 consecutive lanes = the 8 k-chunks
// of one row, i.e. one quarter-warp phase hits 8 planes at the same row offset: a plane pitch of 16 (mod 128) bytes spreads them over
// all 32 banks (ncu: the LSU shared-memory wavefronts were the busiest unit, 80 % of peak, with a 4-way conflict per phase).
constexpr int kConvPad = 16;

struct ConvParams {
  const __nv_bfloat16* xa; int Ca;     // first input, channels-last, channel count (multiple of 16)
  const __nv_bfloat16* xb; int Cb;     // second input of a concatenation (or null / 0)
  const unsigned char* wpk;            // packed weights [n_tiles][27][n_cc][KC/8][NT][8] bf16
  const float* bias;                   // [Cout] (padded)
  __nv_bfloat16* y; int Cout;          // bf16 channels-last output with Cout (multiple of 16) channels, or
  float* yf; int cout_real;            // fp32 planar (B, cout_real, D, H, W) output of the first cout_real channels
  int B, D, H, W, KC, n_cc, act;
  long long n_vox;
  int splits;                          // split-K over the 9 (kd, kh) tap rows: blockIdx.z accumulates 9/splits of them into
  float* ws;                           // fp32 partial sums [splits][n_vox][Cout] (bias / activation applied by the reduction)
  const __nv_bfloat16* res;            // optional residual (same shape as y) added after bias / activation (ResidualBlock, ops.py:61-115)
  int wrap;                            // 1: wrap along width (WrapPadding), 0: zeros on every side (plain zero padding)
  int tap_lo, tap_cnt;                 // taps walked by the per-tap variant (= 3 * the tap rows below, or the centre tap alone for a
                                       // pointwise / 1x1x1 convolution)
  int krow_lo, krow_cnt;               // (kd, kh) tap rows walked: all 9, or rows 3..5 when D == 1 (a 2-D convolution: the kd != 1
                                       // taps only ever see the zero padding)
};

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// mbarrier arrive triggered by the completion of all cp.async operations this thread issued before it (counts as one arrival)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// wait with back-off: the epilogue warps of the persistent kernel idle for a whole tile; a tight try_wait loop would take issue slots
// from the gather warps (ncu: ALU was the busiest pipe at 55 %, mostly spin loops)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    if (done) break;
    __nanosleep(200);
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Epilogue of both variants, 8 warps: warp w reads TMEM lanes 32 (w % 4).. (its quadrant) and the column half w / 4 of the tile.
// + bias, LeakyReLU; bf16 16-byte channel runs, fp32 planar for the first cout_real channels, or raw split-K partial sums.
// `cols` columns starting at column `c0` of the tile's accumulator (TMEM address `tacc`), lane quadrant q, voxel tile vt, split z
template <int NT, int COLS>
__device__ __forceinline__ void conv_epilogue_cols(const ConvParams& p, uint32_t tacc, int vt, int nt, int z, int q, int lane, int c0) {
  const long long v = (long long)vt * kConvRows + q * 32 + lane;
  const bool row_ok = v < p.n_vox;
  const uint32_t tq = tacc + ((uint32_t)(q * 32) << 16) + c0;
  const int cbase = nt * NT + c0;
  const float* bias = p.bias + cbase;
  const long long DHW = (long long)p.D * p.H * p.W;
  constexpr int kHalf = COLS;
#pragma unroll
  for (int c = 0; c < kHalf; c += 16) {
    float a[16];
    umma::ld16(tq + c, a);                                            // warp-collective: only the stores are predicated
    if (p.splits > 1) {                                               // raw partial sums; conv3d_reduce_kernel finishes the layer
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(p.ws + ((size_t)z * p.n_vox + v) * p.Cout + cbase + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      a[i] += __ldg(bias + c + i);
      if (p.act) a[i] = a[i] > 0.f ? a[i] : 0.01f * a[i];
    }
    if (!row_ok) continue;
    if (p.yf) {
      const long long bi = v / DHW, rem = v - bi * DHW;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (cbase + c + i < p.cout_real) p.yf[(bi * p.cout_real + cbase + c + i) * DHW + rem] = a[i];
    } else {
      if (p.res) {
        const uint4* rs = reinterpret_cast<const uint4*>(p.res + (size_t)v * p.Cout + cbase + c);
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {
          const uint4 q4 = __ldg(rs + hq);
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q4);
#pragma unroll
          for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); a[8 * hq + 2 * j] += f.x; a[8 * hq + 2 * j + 1] += f.y; }
        }
      }
      __nv_bfloat16* dst = p.y + (size_t)v * p.Cout + cbase + c;
      reinterpret_cast<uint4*>(dst)[0] = make_uint4(umma::pack2(a[0], a[1]), umma::pack2(a[2], a[3]), umma::pack2(a[4], a[5]), umma::pack2(a[6], a[7]));
      reinterpret_cast<uint4*>(dst)[1] = make_uint4(umma::pack2(a[8], a[9]), umma::pack2(a[10], a[11]), umma::pack2(a[12], a[13]), umma::pack2(a[14], a[15]));
    }
  }
}

template <int NT>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, uint32_t tb, int nt, int warp, int lane) {
  constexpr int kHalf = NT >= 32 ? NT / 2 : NT;
  const int q = warp & 3, h = warp >> 2;
  if (NT < 32 && h) return;
  conv_epilogue_cols<NT, kHalf>(p, tb, blockIdx.x, nt, blockIdx.z, q, lane, h * kHalf);
}

// The MMA warp's loop, shared by both variants: one thread waits for a full stage, issues TAPS x KC/16 MMAs, commits to `free`.
template <int NT, int S, int TAPS>
__device__ __forceinline__ void conv_mma_loop(uint32_t tb, const unsigned char* As, int a_bytes, int a_pitch16, const unsigned char* Bs,
                                              int b_bytes, int KC, int n_it, uint64_t* bar_full, uint64_t* bar_free, uint64_t* bar_done) {
  int s = 0;
  uint32_t ph = 0;
#pragma unroll 1
  for (int it = 0; it < n_it; ++it) {
    mbar_wait(&bar_full[s], ph);
    umma::fence_smem_to_async();          // cp.async wrote the stage through the generic proxy; the MMA reads through the async one
    umma::fence_after_sync();
#pragma unroll
    for (int kw = 0; kw < TAPS; ++kw)
      umma::gemm_issue(tb, As + s * a_bytes + kw * 16, a_pitch16, Bs + (s * TAPS + kw) * b_bytes, NT, NT, KC, it > 0 || kw > 0);
    umma::commit(&bar_free[s]);
    if (++s == S) { s = 0; ph ^= 1u; }
  }
  umma::commit(bar_done);
}

// Per-tap variant (any W).  Gather mapping: `kch` consecutive lanes copy one voxel's contiguous channel run (a warp request touches
// 32/kch lines), each thread serving rows row0 + i * row_step, i < kch/2.  Per row: bit (kd*3+kh) = that tap row lies inside the
// volume (zeros along depth / height otherwise), bits 9 / 10 = the voxel sits in the first / last column (wrap along width).
template <int NT, int S>
__global__ void __launch_bounds__(kConvThreads) conv3d_igemm_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_full[S], bar_free[S], bar_done;
  const int r = threadIdx.x, warp = r >> 5;
  const uint32_t a_pitch = kConvRows * 16 + kConvPad;       // bytes between k-chunk planes
  const int kch = p.KC >> 3;                                // 16-byte chunks per row and stage (2, 4 or 8)
  const int a_bytes = kch * a_pitch, b_bytes = p.KC * NT * 2;
  unsigned char* As = smem;
  unsigned char* Bs = smem + S * a_bytes;
  constexpr uint32_t kCols = NT < 32 ? 32 : NT;
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, kCols);
  if (r == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], kGatherThreads + 1); mbar_init(&bar_free[s], 1); }
    mbar_init(&bar_done, 1);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int nt = blockIdx.y;
  const int n_taps = p.tap_cnt / p.splits, tap0 = p.tap_lo + blockIdx.z * n_taps;       // this CTA's share of the taps

  if (warp == 8) {
    if (r == kGatherThreads)
      conv_mma_loop<NT, S, 1>(tb, As, a_bytes, (int)(a_pitch >> 4), Bs, b_bytes, p.KC, n_taps * p.n_cc, bar_full, bar_free, &bar_done);
  } else {
    const int lanes_shift = kch == 8 ? 3 : (kch == 4 ? 2 : 1);
    const int my_chunk = r & (kch - 1), row0 = r >> lanes_shift, row_step = kGatherThreads >> lanes_shift, n_rows = kch >> 1;
    const long long vbase = (long long)blockIdx.x * kConvRows + row0;
    const long long HW = (long long)p.H * p.W;
    uint32_t rowinfo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rowinfo[i] = 0;
      const long long vv = vbase + i * row_step;
      if (i < n_rows && vv < p.n_vox) {
        long long t = vv;
        const int x = (int)(t % p.W); t /= p.W;
        const int y = (int)(t % p.H); t /= p.H;
        const int d = (int)(t % p.D);
        uint32_t m = 0;
        for (int kd = 0; kd < 3; ++kd)
          for (int kh = 0; kh < 3; ++kh)
            if (d + kd - 1 >= 0 && d + kd - 1 < p.D && y + kh - 1 >= 0 && y + kh - 1 < p.H) m |= 1u << (kd * 3 + kh);
        rowinfo[i] = m | (x == 0 ? 512u : 0u) | (x == p.W - 1 ? 1024u : 0u);
      }
    }
    const uint32_t adst0 = smem_u32(As) + my_chunk * a_pitch + row0 * 16;
    const uint32_t dst_step = row_step * 16;
    const unsigned char* wt = p.wpk + ((size_t)nt * 27 + tap0) * p.n_cc * b_bytes;
    int s = 0;
    uint32_t ph = 1;                                                    // parity of the `free` phase that precedes the first use
    uint32_t adst = adst0;
#pragma unroll 1
    for (int tap = tap0; tap < tap0 + n_taps; ++tap) {
      const int krow = tap / 3, kw = tap - krow * 3 - 1;                // krow = kd*3 + kh
      const long long tap_off = (long long)(krow / 3 - 1) * HW + (long long)(krow % 3 - 1) * p.W + kw;
      const uint32_t wrap_bit = kw < 0 ? 512u : (kw > 0 ? 1024u : 0u);
      const long long wrap_off = kw < 0 ? p.W : -p.W;
      long long nv[4];
      uint32_t nb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = ((rowinfo[i] >> krow) & 1u) && (p.wrap || !(rowinfo[i] & wrap_bit));
        nb[i] = ok ? 16u : 0u;
        nv[i] = ok ? vbase + i * row_step + tap_off + ((rowinfo[i] & wrap_bit) ? wrap_off : 0) : 0;
      }
#pragma unroll 1
      for (int inp = 0; inp < 2; ++inp) {                               // the two inputs of the concatenation
        const int cs = inp ? p.Cb : p.Ca;
        if (cs == 0) continue;
        const __nv_bfloat16* rp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rp[i] = (inp ? p.xb : p.xa) + (size_t)nv[i] * cs + my_chunk * 8;
#pragma unroll 1
        for (int c = 0; c < cs; c += p.KC) {
          mbar_wait(&bar_free[s], ph);                                  // the MMAs that read this stage have completed
          if (r == 0) {
            mbar_expect_tx(&bar_full[s], b_bytes);
            bulk_g2s(Bs + s * b_bytes, wt, b_bytes, &bar_full[s]);
          }
          wt += b_bytes;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < n_rows) { cp_async16(adst + i * dst_step, rp[i], nb[i]); rp[i] += p.KC; }
          cp_async_arrive(&bar_full[s]);    // arrives when this thread's copies of the stage have landed; up to S stages in flight
          adst += a_bytes;
          if (++s == S) { s = 0; ph ^= 1u; adst = adst0; }
        }
      }
    }
    mbar_wait(&bar_done, 0);
    umma::fence_after_sync();
    conv_epilogue<NT>(p, tb, nt, warp, r & 31);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc(tb, kCols);
}

// Row variant (W % 128 == 0: every tile is a run of 128 voxels inside ONE image row).  A pipeline stage holds that run of one
// (kd, kh) neighbour row plus its two wrap-around halo voxels — 130 operand rows — and serves all three kw taps: in the un-swizzled
// k-chunk-major layout operand rows are 16 bytes apart, so tap kw is the same buffer with the descriptor start advanced by kw rows.
// One gather (and one third of the copy instructions / L1TEX traffic) feeds three taps of MMAs.
constexpr int kRowA = kConvRows + 2;
constexpr int kRowPitch = kRowA * 16 + 112;       // bytes between k-chunk planes: 16 (mod 128), see kConvPad

template <int NT, int S>
__global__ void __launch_bounds__(kConvThreads) conv3d_igemm_row_kernel(const ConvParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_full[S], bar_free[S], bar_done;
  const int r = threadIdx.x, warp = r >> 5;
  const int kch = p.KC >> 3;
  const int a_bytes = kch * kRowPitch, b_bytes = p.KC * NT * 2;
  unsigned char* As = smem;
  unsigned char* Bs = smem + S * a_bytes;
  constexpr uint32_t kCols = NT < 32 ? 32 : NT;
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, kCols);
  if (r == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], kGatherThreads + 1); mbar_init(&bar_free[s], 1); }
    mbar_init(&bar_done, 1);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int nt = blockIdx.y;
  const int n_krow = p.krow_cnt / p.splits, krow0 = p.krow_lo + blockIdx.z * n_krow;

  if (warp == 8) {
    if (r == kGatherThreads)
      conv_mma_loop<NT, S, 3>(tb, As, a_bytes, kRowPitch >> 4, Bs, b_bytes, p.KC, n_krow * p.n_cc, bar_full, bar_free, &bar_done);
  } else {
    const int lanes_shift = kch == 8 ? 3 : (kch == 4 ? 2 : 1);
    const int my_chunk = r & (kch - 1), row0 = r >> lanes_shift, row_step = kGatherThreads >> lanes_shift, n_rows = kch >> 1;
    // the tile: voxels [v0, v0 + 128) of image row (b, d, y), starting at column x0
    const long long v0 = (long long)blockIdx.x * kConvRows;
    const long long HW = (long long)p.H * p.W;
    long long t = v0;
    const int x0 = (int)(t % p.W); t /= p.W;
    const int y = (int)(t % p.H); t /= p.H;
    const int d = (int)(t % p.D);
    // halo voxels (wrap along width): threads [0, kch) copy the left one (operand row 0), [kch, 2 kch) the right one (row 129)
    const bool halo = r < 2 * kch;
    const long long halo_off = r < kch ? (x0 == 0 ? (long long)p.W - 1 : -1) : (x0 + kConvRows == p.W ? (long long)kConvRows - p.W : kConvRows);
    const int hc = r < kch ? r : r - kch;
    const bool halo_zero = !p.wrap && (r < kch ? x0 == 0 : x0 + kConvRows == p.W);      // zero padding: no neighbour across the seam
    const uint32_t hdst0 = smem_u32(As) + hc * kRowPitch + (r < kch ? 0 : kRowA - 1) * 16;
    const uint32_t adst0 = smem_u32(As) + my_chunk * kRowPitch + (row0 + 1) * 16;
    const uint32_t dst_step = row_step * 16;
    const size_t wtap = (size_t)p.n_cc * b_bytes;                        // distance between the weight blocks of two taps
    const unsigned char* wt = p.wpk + ((size_t)nt * 27 + krow0 * 3) * wtap;
    int s = 0;
    uint32_t ph = 1, adst = adst0, hdst = hdst0;
#pragma unroll 1
    for (int krow = krow0; krow < krow0 + n_krow; ++krow) {              // krow = kd*3 + kh
      const int dd = d + krow / 3 - 1, yy = y + krow % 3 - 1;
      const bool ok = dd >= 0 && dd < p.D && yy >= 0 && yy < p.H;        // zeros along depth / height (uniform over the tile)
      const uint32_t nbytes = ok ? 16u : 0u;
      const long long nv0 = ok ? v0 + (long long)(krow / 3 - 1) * HW + (long long)(krow % 3 - 1) * p.W : 0;
      const long long nvr = ok ? nv0 + row0 : 0, nvh = ok ? nv0 + halo_off : 0;
      const long long rstep = ok ? row_step : 0;
#pragma unroll 1
      for (int inp = 0; inp < 2; ++inp) {                                 // the two inputs of the concatenation
        const int cs = inp ? p.Cb : p.Ca;
        if (cs == 0) continue;
        const __nv_bfloat16* base = inp ? p.xb : p.xa;
        const __nv_bfloat16* rp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rp[i] = base + (size_t)(nvr + i * rstep) * cs + my_chunk * 8;
        const __nv_bfloat16* hp = base + (size_t)nvh * cs + hc * 8;
#pragma unroll 1
        for (int c = 0; c < cs; c += p.KC) {
          mbar_wait(&bar_free[s], ph);
          if (r == 0) {
            mbar_expect_tx(&bar_full[s], 3 * b_bytes);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) bulk_g2s(Bs + (s * 3 + kw) * b_bytes, wt + kw * wtap, b_bytes, &bar_full[s]);
          }
          wt += b_bytes;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < n_rows) { cp_async16(adst + i * dst_step, rp[i], nbytes); rp[i] += p.KC; }
          if (halo) { cp_async16(hdst, hp, halo_zero ? 0u : nbytes); hp += p.KC; }
          cp_async_arrive(&bar_full[s]);
          adst += a_bytes;
          hdst += a_bytes;
          if (++s == S) { s = 0; ph ^= 1u; adst = adst0; hdst = hdst0; }
        }
      }
      wt += 2 * wtap;                                                     // next tap row: skip the two taps already consumed
    }
    mbar_wait(&bar_done, 0);
    umma::fence_after_sync();
    conv_epilogue<NT>(p, tb, nt, warp, r & 31);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc(tb, kCols);
}

// ---- persistent variant ---------------------------------------------------------------------------------------------------------
// One CTA per SM slot walks work items (voxel tile, channel tile, split).  13 warps: 0-7 gather (the stage ring runs on across work
// items, so the copies of the next tile are in flight while the current one drains), 8 issues the MMAs into one of TWO accumulators
// in tensor memory, 9-12 are the epilogue of the other accumulator: the pipeline never drains between tiles.
constexpr int kPersistThreads = 416;

template <int NT, int S, bool ROW>
__global__ void __launch_bounds__(kPersistThreads) conv3d_persist_kernel(const ConvParams p, int n_vt) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_full[S], bar_free[S], acc_full[2], acc_empty[2];
  constexpr int TAPS = ROW ? 3 : 1;
  const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
  const uint32_t a_pitch = ROW ? kRowPitch : kConvRows * 16 + kConvPad;
  const int kch = p.KC >> 3;
  const int a_bytes = kch * a_pitch, b_bytes = p.KC * NT * 2;
  unsigned char* As = smem;
  unsigned char* Bs = smem + S * a_bytes;
  constexpr uint32_t kCols = 2 * NT < 32 ? 32 : 2 * NT;
  if (warp == 8) umma::tmem_alloc(&tmem_base_s, kCols);
  if (r == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], kGatherThreads + 1); mbar_init(&bar_free[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int n_nt = p.Cout / NT;
  const int total = n_vt * n_nt * p.splits;
  const int units = (ROW ? p.krow_cnt : p.tap_cnt) / p.splits;          // tap rows (ROW) or taps per work item
  const int n_st = units * p.n_cc;                                      // pipeline stages per work item

  if (warp < 8) {
    // ================= gather warps =================
    const int lanes_shift = kch == 8 ? 3 : (kch == 4 ? 2 : 1);
    const int my_chunk = r & (kch - 1), row0 = r >> lanes_shift, row_step = kGatherThreads >> lanes_shift, n_rows = kch >> 1;
    const long long HW = (long long)p.H * p.W;
    const uint32_t dst_step = row_step * 16;
    const uint32_t adst0 = smem_u32(As) + my_chunk * a_pitch + (row0 + (ROW ? 1 : 0)) * 16;
    const bool halo = ROW && r < 2 * kch;
    const int hc = r < kch ? r : r - kch;
    const uint32_t hdst0 = smem_u32(As) + hc * a_pitch + (r < kch ? 0 : kRowA - 1) * 16;
    const size_t wtap = (size_t)p.n_cc * b_bytes;                       // distance between the weight blocks of two taps
    int s = 0;
    uint32_t ph = 1, adst = adst0, hdst = hdst0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      const int vt = w % n_vt, rest = w / n_vt, nt = rest % n_nt, z = rest / n_nt;
      const long long v0 = (long long)vt * kConvRows;
      const unsigned char* wt = p.wpk + ((size_t)nt * 27 + (ROW ? (p.krow_lo + z * units) * 3 : p.tap_lo + z * units)) * wtap;
      // per-tile coordinates: ROW -> (x0, y, d) of the run; per-tap -> validity / wrap bits of this thread's rows
      int x0 = 0, ty = 0, td = 0;
      uint32_t rowinfo[4] = {0u, 0u, 0u, 0u};
      long long halo_off = 0;
      bool halo_zero = false;
      if (ROW) {
        long long t = v0;
        x0 = (int)(t % p.W); t /= p.W;
        ty = (int)(t % p.H); t /= p.H;
        td = (int)(t % p.D);
        halo_off = r < kch ? (x0 == 0 ? (long long)p.W - 1 : -1) : (x0 + kConvRows == p.W ? (long long)kConvRows - p.W : kConvRows);
        halo_zero = !p.wrap && (r < kch ? x0 == 0 : x0 + kConvRows == p.W);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long vv = v0 + row0 + i * row_step;
          if (i < n_rows && vv < p.n_vox) {
            long long t = vv;
            const int x = (int)(t % p.W); t /= p.W;
            const int y = (int)(t % p.H); t /= p.H;
            const int d = (int)(t % p.D);
            uint32_t m = 0;
            for (int kd = 0; kd < 3; ++kd)
              for (int kh = 0; kh < 3; ++kh)
                if (d + kd - 1 >= 0 && d + kd - 1 < p.D && y + kh - 1 >= 0 && y + kh - 1 < p.H) m |= 1u << (kd * 3 + kh);
            rowinfo[i] = m | (x == 0 ? 512u : 0u) | (x == p.W - 1 ? 1024u : 0u);
          }
        }
      }
#pragma unroll 1
      for (int u = 0; u < units; ++u) {
        // source rows of this tap row (ROW) / tap
        long long nv[4], nvh = 0;
        uint32_t nb[4];
        if (ROW) {
          const int krow = p.krow_lo + z * units + u;
          const int dd = td + krow / 3 - 1, yy = ty + krow % 3 - 1;
          const bool ok = dd >= 0 && dd < p.D && yy >= 0 && yy < p.H;
          const long long nv0 = ok ? v0 + (long long)(krow / 3 - 1) * HW + (long long)(krow % 3 - 1) * p.W : 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) { nb[i] = ok ? 16u : 0u; nv[i] = ok ? nv0 + row0 + i * row_step : 0; }
          nvh = ok ? nv0 + halo_off : 0;
        } else {
          const int tap = p.tap_lo + z * units + u;
          const int krow = tap / 3, kw = tap - krow * 3 - 1;
          const long long tap_off = (long long)(krow / 3 - 1) * HW + (long long)(krow % 3 - 1) * p.W + kw;
          const uint32_t wrap_bit = kw < 0 ? 512u : (kw > 0 ? 1024u : 0u);
          const long long wrap_off = kw < 0 ? p.W : -p.W;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = ((rowinfo[i] >> krow) & 1u) && (p.wrap || !(rowinfo[i] & wrap_bit));
            nb[i] = ok ? 16u : 0u;
            nv[i] = ok ? v0 + row0 + i * row_step + tap_off + ((rowinfo[i] & wrap_bit) ? wrap_off : 0) : 0;
          }
        }
#pragma unroll 1
        for (int inp = 0; inp < 2; ++inp) {                             // the two inputs of the concatenation
          const int cs = inp ? p.Cb : p.Ca;
          if (cs == 0) continue;
          const __nv_bfloat16* base = inp ? p.xb : p.xa;
          const __nv_bfloat16* rp[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) rp[i] = base + (size_t)nv[i] * cs + my_chunk * 8;
          const __nv_bfloat16* hp = base + (size_t)nvh * cs + hc * 8;
#pragma unroll 1
          for (int c = 0; c < cs; c += p.KC) {
            mbar_wait(&bar_free[s], ph);                                // the MMAs that read this stage have completed
            if (r == 0) {
              mbar_expect_tx(&bar_full[s], TAPS * b_bytes);
#pragma unroll
              for (int kw = 0; kw < TAPS; ++kw) bulk_g2s(Bs + (s * TAPS + kw) * b_bytes, wt + kw * wtap, b_bytes, &bar_full[s]);
            }
            wt += b_bytes;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < n_rows) { cp_async16(adst + i * dst_step, rp[i], nb[i]); rp[i] += p.KC; }
            if (halo) { cp_async16(hdst, hp, halo_zero ? 0u : nb[0]); hp += p.KC; }
            cp_async_arrive(&bar_full[s]);
            adst += a_bytes;
            hdst += a_bytes;
            if (++s == S) { s = 0; ph ^= 1u; adst = adst0; hdst = hdst0; }
          }
        }
        if (ROW) wt += 2 * wtap;                                        // next tap row: skip the two taps already consumed
      }
    }
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int s = 0, it = 0;
      uint32_t ph = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);                 // the epilogue has drained this accumulator
        umma::fence_after_sync();
        const uint32_t tacc = tb + ab * NT;
#pragma unroll 1
        for (int st = 0; st < n_st; ++st) {
          mbar_wait(&bar_full[s], ph);
          umma::fence_smem_to_async();
          umma::fence_after_sync();
#pragma unroll
          for (int kw = 0; kw < TAPS; ++kw)
            umma::gemm_issue(tacc, As + s * a_bytes + kw * 16, (int)(a_pitch >> 4), Bs + (s * TAPS + kw) * b_bytes, NT, NT, p.KC,
                             st > 0 || kw > 0);
          umma::commit(&bar_free[s]);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        umma::commit(&acc_full[ab]);
      }
    }
  } else {
    // ================= epilogue warps (TMEM lane quadrant = warp % 4) =================
    const int q = warp & 3;
    int it = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total; w += gridDim.x, ++it) {
      const int vt = w % n_vt, rest = w / n_vt, nt = rest % n_nt, z = rest / n_nt;
      const int ab = it & 1;
      mbar_wait_backoff(&acc_full[ab], (it >> 1) & 1);
      umma::fence_after_sync();
      conv_epilogue_cols<NT, NT>(p, tb + ab * NT, vt, nt, z, q, lane, 0);
      umma::fence_before_sync();
      mbar_arrive(&acc_empty[ab]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc(tb, kCols);
}

// split-K reduction: sum of the partial volumes + bias + LeakyReLU -> bf16 channels-last (or fp32 planar); thread = (voxel, 8 channels)
__global__ void __launch_bounds__(256) conv3d_reduce_kernel(const ConvParams p) {
  const int C8 = p.Cout / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_vox * C8) return;
  const long long v = i / C8;
  const int c = (int)(i - v * C8) * 8;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = __ldg(p.bias + c + k);
  for (int z = 0; z < p.splits; ++z) {
    const float4* src = reinterpret_cast<const float4*>(p.ws + ((size_t)z * p.n_vox + v) * p.Cout + c);
    const float4 q0 = __ldcs(src), q1 = __ldcs(src + 1);
    a[0] += q0.x; a[1] += q0.y; a[2] += q0.z; a[3] += q0.w; a[4] += q1.x; a[5] += q1.y; a[6] += q1.z; a[7] += q1.w;
  }
  if (p.act) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = a[k] > 0.f ? a[k] : 0.01f * a[k];
  }
  if (p.yf) {
    const long long DHW = (long long)p.D * p.H * p.W;
    const long long bi = v / DHW, rem = v - bi * DHW;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c + k < p.cout_real) p.yf[(bi * p.cout_real + c + k) * DHW + rem] = a[k];
  } else {
    if (p.res) {
      const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(p.res + (size_t)v * p.Cout + c));
      const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&q4);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(hh[j]); a[2 * j] += f.x; a[2 * j + 1] += f.y; }
    }
    *reinterpret_cast<uint4*>(p.y + (size_t)v * p.Cout + c) =
        make_uint4(umma::pack2(a[0], a[1]), umma::pack2(a[2], a[3]), umma::pack2(a[4], a[5]), umma::pack2(a[6], a[7]));
  }
}

// Single-output-channel 3x3x3 convolution over MANY input channels, second half.  The layer  out[v] = b + sum_tap sum_c w[tap][c]
// x[nbr(v, tap)][c]  is evaluated as  z[tap][u] = sum_c w[tap][c] x[u][c]  (a pointwise GEMM with the 27 taps as output channels:
// every voxel's channels are read ONCE instead of once per tap row) followed by this 27-tap stencil over the scalar planes:
// out[v] = act(b + sum_tap z[tap][nbr(v, tap)]), zeros along depth / height, wrap along width (WrapPadding3D).  z is fp32 planar
// (B, 27, D, H, W): a warp reads 128 contiguous bytes per tap.
__global__ void __launch_bounds__(256) conv3d_tapsum_kernel(const float* __restrict__ z, float bias, int B, int D, int H, int W, int act,
                                                            float* __restrict__ out) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W, DHW = HW * D;
  if (v >= (unsigned)B * (unsigned)DHW) return;
  unsigned t = v;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const long long b = t / D;
  const float* zb = z + b * 27 * DHW;
  const int xm = x == 0 ? W - 1 : x - 1, xp = x == W - 1 ? 0 : x + 1;
  float acc = bias;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int dd = d + kd - 1;
    if (dd < 0 || dd >= D) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      if (yy < 0 || yy >= H) continue;
      const float* row = zb + (long long)(kd * 9 + kh * 3) * DHW + (long long)dd * HW + (long long)yy * W;
      acc += __ldg(row + xm) + __ldg(row + DHW + x) + __ldg(row + 2 * DHW + xp);
    }
  }
  if (act) acc = acc > 0.f ? acc : 0.01f * acc;
  out[v] = acc;
}

// 1 -> 1 channel 3x3x3 convolution on an fp32 scalar volume (conv2 of the last decoder): 27-tap weighted stencil, thread = voxel
struct Taps27 { float w[27]; };
__global__ void __launch_bounds__(256) conv3d_scalar_kernel(const float* __restrict__ xin, const Taps27 taps, float bias, int B, int D, int H,
                                                            int W, int act, float* __restrict__ out) {
  const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
  const long long HW = (long long)H * W, DHW = HW * D;
  if (v >= (unsigned)B * (unsigned)DHW) return;
  unsigned t = v;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const float* xb = xin + (t / D) * DHW;
  const int xm = x == 0 ? W - 1 : x - 1, xp = x == W - 1 ? 0 : x + 1;
  float acc = bias;
#pragma unroll
  for (int kd = 0; kd < 3; ++kd) {
    const int dd = d + kd - 1;
    if (dd < 0 || dd >= D) continue;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int yy = y + kh - 1;
      if (yy < 0 || yy >= H) continue;
      const float* row = xb + (long long)dd * HW + (long long)yy * W;
      const float* w = taps.w + kd * 9 + kh * 3;
      acc = fmaf(w[0], __ldg(row + xm), acc);
      acc = fmaf(w[1], __ldg(row + x), acc);
      acc = fmaf(w[2], __ldg(row + xp), acc);
    }
  }
  if (act) acc = acc > 0.f ? acc : 0.01f * acc;
  out[v] = acc;
}

// single output channel (last decoder): fp32 SIMT.  in: bf16 channels-last (two concatenated inputs) or fp32 single channel
__global__ void __launch_bounds__(256) conv3d_cout1_kernel(const __nv_bfloat16* __restrict__ xa, int Ca, const __nv_bfloat16* __restrict__ xb,
                                                           int Cb, const float* __restrict__ xf, const float* __restrict__ w /* [27][Cin] */,
                                                           float bias, int B, int D, int H, int W, int act, float* __restrict__ out) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= (long long)B * D * H * W) return;
  long long t = v;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H); t /= H;
  const int d = (int)(t % D);
  const int b = (int)(t / D);
  const int Cin = xf ? 1 : Ca + Cb;
  float acc = bias;
  for (int tap = 0; tap < 27; ++tap) {
    const int dd = d + tap / 9 - 1, yy = y + (tap / 3) % 3 - 1;
    int xx = x + tap % 3 - 1;
    xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);
    if (dd < 0 || dd >= D || yy < 0 || yy >= H) continue;
    const size_t nv = (((size_t)b * D + dd) * H + yy) * W + xx;
    const float* wt = w + tap * Cin;
    if (xf) { acc = fmaf(__ldg(xf + nv), wt[0], acc); continue; }
    const uint4* pa = reinterpret_cast<const uint4*>(xa + nv * Ca);
    for (int j = 0; j < Ca / 8; ++j) {
      const uint4 q = __ldg(pa + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc = fmaf(f.x, wt[8 * j + 2 * i], acc); acc = fmaf(f.y, wt[8 * j + 2 * i + 1], acc); }
    }
    if (xb) {
      const uint4* pb = reinterpret_cast<const uint4*>(xb + nv * Cb);
      for (int j = 0; j < Cb / 8; ++j) {
        const uint4 q = __ldg(pb + j);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); acc = fmaf(f.x, wt[Ca + 8 * j + 2 * i], acc); acc = fmaf(f.y, wt[Ca + 8 * j + 2 * i + 1], acc); }
      }
    }
  }
  if (act) acc = acc > 0.f ? acc : 0.01f * acc;
  out[v] = acc;
}

// AvgPool3d(2) on bf16 channels-last: thread = (output voxel, 8-channel chunk)
__global__ void __launch_bounds__(256) avgpool3d2_kernel(const __nv_bfloat16* __restrict__ x, int B, int D, int H, int W, int C,
                                                         __nv_bfloat16* __restrict__ y) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)B * Do * Ho * Wo * C8) return;
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int dO = (int)(t % Do);
  const int b = (int)(t / Do);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dd = 2 * dO + (k >> 2), yy = 2 * yo + ((k >> 1) & 1), xx = 2 * xo + (k & 1);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + ((((size_t)b * D + dd) * H + yy) * W + xx) * C) + c8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
  }
  uint4 o;
  o.x = umma::pack2(acc[0] * 0.125f, acc[1] * 0.125f); o.y = umma::pack2(acc[2] * 0.125f, acc[3] * 0.125f);
  o.z = umma::pack2(acc[4] * 0.125f, acc[5] * 0.125f); o.w = umma::pack2(acc[6] * 0.125f, acc[7] * 0.125f);
  reinterpret_cast<uint4*>(y + (size_t)(i / C8) * C)[c8] = o;
}

// x2 trilinear upsampling, align_corners=False (ATen area_pixel_compute_source_index): thread = (output voxel, 8-channel chunk)
__device__ __forceinline__ void up_src(int o, int n, int& i0, int& i1, float& l) {
  float s = 0.5f * ((float)o + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  i1 = i0 + (i0 < n - 1 ? 1 : 0);
  l = s - (float)i0;
}
__global__ void __launch_bounds__(256) upsample3d2_kernel(const __nv_bfloat16* __restrict__ x, int B, int D, int H, int W, int C,
                                                          int scale_d, __nv_bfloat16* __restrict__ y) {
  const int Do = scale_d * D, Ho = 2 * H, Wo = 2 * W, C8 = C / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;          // 32-bit index math (the launcher checks the range): the
  if (i >= (unsigned)B * Do * Ho * Wo * C8) return;                 // 64-bit divisions were most of this kernel's instructions
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xo = (int)(t % Wo); t /= Wo;
  const int yo = (int)(t % Ho); t /= Ho;
  const int dO = (int)(t % Do);
  const int b = (int)(t / Do);
  int d0, d1, y0, y1, x0, x1;
  float ld, ly, lx;
  if (scale_d == 2) up_src(dO, D, d0, d1, ld); else { d0 = d1 = dO; ld = 0.f; }      // scale_d == 1: bilinear (2-D feature maps)
  up_src(yo, H, y0, y1, ly); up_src(xo, W, x0, x1, lx);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint4 q[8];                                                        // all eight corner loads in flight before the first use
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dd = (k & 4) ? d1 : d0, yy = (k & 2) ? y1 : y0, xx = (k & 1) ? x1 : x0;
    q[k] = __ldg(reinterpret_cast<const uint4*>(x + ((((size_t)b * D + dd) * H + yy) * W + xx) * C) + c8);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float wgt = ((k & 4) ? ld : 1.f - ld) * ((k & 2) ? ly : 1.f - ly) * ((k & 1) ? lx : 1.f - lx);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); acc[2 * j] = fmaf(wgt, f.x, acc[2 * j]); acc[2 * j + 1] = fmaf(wgt, f.y, acc[2 * j + 1]); }
  }
  uint4 o;
  o.x = umma::pack2(acc[0], acc[1]); o.y = umma::pack2(acc[2], acc[3]); o.z = umma::pack2(acc[4], acc[5]); o.w = umma::pack2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(y + (size_t)(i / C8) * C)[c8] = o;
}

// fp32 (B,C,D,H,W) with arbitrary element strides -> bf16 channels-last (B,D,H,W,Cpad), zero padded channels
__global__ void __launch_bounds__(256) to_bf16_cl_kernel(const float* __restrict__ x, long long sb, long long sc, long long sd, long long sh,
                                                         long long sw, int B, int C, int D, int H, int W, int Cpad,
                                                         __nv_bfloat16* __restrict__ y) {
  const int C8 = Cpad / 8;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)B * D * H * W * C8) return;
  unsigned t = i;
  const int c8 = (int)(t % C8); t /= C8;
  const int xx = (int)(t % W); t /= W;
  const int yy = (int)(t % H); t /= H;
  const int dd = (int)(t % D);
  const int b = (int)(t / D);
  const float* src = x + b * sb + dd * sd + yy * sh + xx * sw;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { const int c = 8 * c8 + k; v[k] = c < C ? __ldg(src + c * sc) : 0.f; }
  uint4 o;
  o.x = umma::pack2(v[0], v[1]); o.y = umma::pack2(v[2], v[3]); o.z = umma::pack2(v[4], v[5]); o.w = umma::pack2(v[6], v[7]);
  reinterpret_cast<uint4*>(y + (size_t)(i / C8) * Cpad)[c8] = o;
}

// decoders1 of the MVS head (models/test_models.py:147-158, pipeline3_model.py:866-879): 1x1 convolution over the depth axis of the
// regularised cost (C = D channels -> 1) followed by F.interpolate(scale_factor, 'bilinear', align_corners=False) and the depth
// rectification (1: clamp(min=0), 2: 1 / (clamp(min=0) + 1e-10)).  The 1x1 convolution commutes with the (linear) interpolation, so
// each output pixel blends its four source pixels' dot products; thread = output pixel.  fp32 throughout.
__global__ void __launch_bounds__(256) channel_dot_upsample_kernel(const float* __restrict__ x, long long sb, long long sc, long long sh,
                                                                   long long sw, int B, int C, int H, int W, const float* __restrict__ w,
                                                                   float bias, int scale, int rectify, float* __restrict__ out) {
  const int Ho = H * scale, Wo = W * scale;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * Ho * Wo) return;
  const int xo = (int)(i % Wo), yo = (int)((i / Wo) % Ho), b = (int)(i / ((long long)Wo * Ho));
  const float inv = 1.f / (float)scale;
  float sy = ((float)yo + 0.5f) * inv - 0.5f, sx = ((float)xo + 0.5f) * inv - 0.5f;      // ATen area_pixel_compute_source_index
  sy = sy < 0.f ? 0.f : sy; sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float* p00 = x + b * sb + y0 * sh + x0 * sw;
  const float* p01 = x + b * sb + y0 * sh + x1 * sw;
  const float* p10 = x + b * sb + y1 * sh + x0 * sw;
  const float* p11 = x + b * sb + y1 * sh + x1 * sw;
  float a00 = bias, a01 = bias, a10 = bias, a11 = bias;
  for (int c = 0; c < C; ++c) {
    const float wc = __ldg(w + c);
    a00 = fmaf(wc, __ldg(p00 + c * sc), a00); a01 = fmaf(wc, __ldg(p01 + c * sc), a01);
    a10 = fmaf(wc, __ldg(p10 + c * sc), a10); a11 = fmaf(wc, __ldg(p11 + c * sc), a11);
  }
  float v = (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
  if (rectify) v = fmaxf(v, 0.f);
  if (rectify == 2) v = 1.0f / (v + 1e-10f);
  out[i] = v;
}

}  // namespace pgrf

using namespace pgrf;

namespace pgrf {
int g_conv_stages = 0;   // pipeline depth of the conv3d kernels (pgrf_debug_set "conv_stages": 0 = automatic, else 3..6)
int g_conv_row = 1;      // use the row variant when W % 128 == 0 (pgrf_debug_set "conv_row")
int g_conv_kc = 64;       // channels per pipeline stage (pgrf_debug_set "conv_kc": 64, 32 or 16)
int g_conv_smem_kb = 56;  // shared-memory budget per CTA that sizes the pipeline (pgrf_debug_set "conv_smem_kb")
int g_conv_persist = 1;      // persistent CTAs with double-buffered accumulators (pgrf_debug_set "conv_persist")
int g_conv_persist_kb = 100; // their shared-memory budget per CTA (pgrf_debug_set "conv_persist_kb")
int g_conv_splits = 0;   // split-K factor (pgrf_debug_set "conv_splits": 0 = automatic, else 1, 3 or 9)
}

static inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
// the element-wise kernels decode their thread index with 32-bit arithmetic
#define PGRF_REQUIRE_32BIT(n, what) PGRF_REQUIRE((n) < 4294967040LL, what ": %lld work items exceed the 32-bit index range", (long long)(n))

extern "C" int pgrf_conv3d_to_bf16_cl(const float* x, long long sb, long long sc, long long sd, long long sh, long long sw, int B, int C,
                                      int D, int H, int W, int Cpad, void* y, void* stream) {
  PGRF_REQUIRE(x && y && B >= 1 && C >= 1 && Cpad >= C && Cpad % 8 == 0, "conv3d_to_bf16_cl: bad arguments");
  const long long n = (long long)B * D * H * W * (Cpad / 8);
  PGRF_REQUIRE_32BIT(n, "conv3d_to_bf16_cl");
  to_bf16_cl_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, sb, sc, sd, sh, sw, B, C, D, H, W, Cpad, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

struct ConvPlan {
  int KC, n_cc, NT, row, S, splits, krow_lo, krow_cnt;
  size_t smem;
  dim3 grid;
  long long ws_floats;
};

// tile / pipeline / split-K choice of one layer (shared by the launcher and the workspace query)
static int conv3d_plan(int Ca, int Cb, int Cout, int B, int D, int H, int W, bool pointwise, ConvPlan& pl) {
  PGRF_REQUIRE(Ca >= 16 && Ca % 16 == 0 && Cb >= 0 && Cb % 16 == 0 && Cout >= 16 && Cout % 16 == 0, "conv3d: channel counts must be "
               "multiples of 16 (Ca=%d Cb=%d Cout=%d)", Ca, Cb, Cout);
  PGRF_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 2, "conv3d: bad volume %dx%dx%dx%d", B, D, H, W);
  const int Cin = Ca + Cb;
  int KC = Cin < g_conv_kc ? Cin : g_conv_kc;
  while (Ca % KC || Cb % KC) KC >>= 1;           // a stage never straddles the two inputs of a concatenation
  PGRF_REQUIRE(KC >= 16, "conv3d: no common chunk size for Ca=%d Cb=%d", Ca, Cb);
  pl.KC = KC; pl.n_cc = Cin / KC;
  pl.NT = Cout >= 128 ? 128 : Cout;
  PGRF_REQUIRE(pl.NT == 16 || pl.NT == 32 || pl.NT == 64 || pl.NT == 128, "conv3d: Cout=%d unsupported", Cout);
  PGRF_REQUIRE(Cout % pl.NT == 0, "conv3d: Cout=%d not a multiple of the channel tile", Cout);
  const long long n_vox = (long long)B * D * H * W;
  pl.grid = dim3(blocks_for(n_vox, kConvRows), (unsigned)(Cout / pl.NT), 1);
  const long long n_cta = (long long)pl.grid.x * pl.grid.y;
  pl.row = g_conv_row && W % kConvRows == 0 && !pointwise;
  const size_t stage = pl.row ? (size_t)(KC / 8) * kRowPitch + 3 * (size_t)KC * pl.NT * 2            // [130-row operand | three taps of weights]
                              : (size_t)(KC / 8) * (kConvRows * 16 + kConvPad) + (size_t)KC * pl.NT * 2;
  // split-K over the nine (kd, kh) tap rows when the (voxel tile x channel tile) grid cannot fill the GPU (2 CTAs per SM assumed)
  pl.krow_lo = D == 1 ? 3 : 0;                   // D == 1: a 2-D convolution, only the kd == 1 tap rows can see data
  pl.krow_cnt = D == 1 ? 3 : 9;
  pl.splits = 1;
  if (g_conv_splits > 0) pl.splits = g_conv_splits;
  else if (n_cta < 444) pl.splits = n_cta * 3 >= 740 ? 3 : 9;
  PGRF_REQUIRE(pl.splits == 1 || pl.splits == 3 || pl.splits == 9, "conv3d: splits=%d (1, 3 or 9)", pl.splits);
  if (pl.splits > pl.krow_cnt) pl.splits = pl.krow_cnt;
  if (pointwise) pl.splits = 1;
  // pipeline depth: as many stages as keep two CTAs on an SM (one CTA's epilogue hides behind the other's main loop), at least 2 / 3
  const int s_min = pl.row ? 2 : 3, s_max = pl.row ? 4 : 6;
  int S = (int)(((size_t)g_conv_smem_kb * 1024) / stage);
  S = S < s_min ? s_min : (S > s_max ? s_max : S);
  if (!pl.row && S == 5) S = 4;
  if (g_conv_stages) S = g_conv_stages < s_min ? s_min : (g_conv_stages > s_max ? s_max : g_conv_stages);
  if (!pl.row && S == 5) S = 6;
  pl.S = S;
  pl.smem = S * stage;
  PGRF_REQUIRE(pl.smem <= 227 * 1024, "conv3d: %zu bytes of shared memory", pl.smem);
  pl.grid.z = pl.splits;
  pl.ws_floats = pl.splits > 1 ? (long long)pl.splits * n_vox * Cout : 0;
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_workspace(int Ca, int Cb, int Cout, int B, int D, int H, int W, long long* ws_floats) {
  PGRF_REQUIRE(ws_floats, "conv3d_workspace: null pointer argument");
  ConvPlan pl;
  const int rc = conv3d_plan(Ca, Cb, Cout, B, D, H, W, false, pl);
  if (rc != PGRF_OK) return rc;
  *ws_floats = pl.ws_floats;
  return PGRF_OK;
}

static int conv3d_launch(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, float* yf,
                         int cout_real, int Cout, int B, int D, int H, int W, int act, float* ws, long long ws_floats, bool pointwise,
                         const void* res, int wrap, void* stream) {
  PGRF_REQUIRE(xa && wpk && bias && ((y != nullptr) != (yf != nullptr)), "conv3d: null pointer argument (exactly one of y / yf)");
  PGRF_REQUIRE(Cb == 0 || xb, "conv3d: Cb=%d without a second input", Cb);
  PGRF_REQUIRE(!yf || (cout_real >= 1 && cout_real <= Cout), "conv3d: cout_real=%d outside [1, %d]", cout_real, Cout);
  ConvPlan pl;
  const int rc = conv3d_plan(Ca, Cb, Cout, B, D, H, W, pointwise, pl);
  if (rc != PGRF_OK) return rc;
  PGRF_REQUIRE(pl.ws_floats == 0 || (ws && ws_floats >= pl.ws_floats), "conv3d: workspace of %lld floats needed (pgrf_conv3d_workspace), "
               "%lld given", pl.ws_floats, ws ? ws_floats : 0LL);
  ConvParams p;
  p.xa = (const __nv_bfloat16*)xa; p.Ca = Ca; p.xb = (const __nv_bfloat16*)xb; p.Cb = Cb;
  p.wpk = (const unsigned char*)wpk; p.bias = bias; p.y = (__nv_bfloat16*)y; p.Cout = Cout; p.yf = yf; p.cout_real = cout_real;
  p.B = B; p.D = D; p.H = H; p.W = W; p.act = act; p.KC = pl.KC; p.n_cc = pl.n_cc;
  p.n_vox = (long long)B * D * H * W;
  PGRF_REQUIRE(!res || y, "conv3d: a residual needs the bf16 channels-last output");
  p.res = (const __nv_bfloat16*)res; p.wrap = wrap ? 1 : 0;
  p.splits = pl.splits; p.ws = ws; p.krow_lo = pl.krow_lo; p.krow_cnt = pl.krow_cnt;
  p.tap_lo = pointwise ? 13 : 3 * pl.krow_lo;          // 13 = (kd, kh, kw) = (1, 1, 1)
  p.tap_cnt = pointwise ? 1 : 3 * pl.krow_cnt;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid = pl.grid;
  const size_t smem = pl.smem;
  const int S = pl.S;
  if (g_conv_persist) {
    // persistent CTAs: pipeline depth from the shared-memory budget (two CTAs per SM by default), grid = the SM slots
    const size_t stage = pl.smem / pl.S;
    int Sp = (int)(((size_t)g_conv_persist_kb * 1024) / stage);
    Sp = Sp < 2 ? 2 : (Sp > 6 ? 6 : Sp);
    if (Sp == 5) Sp = 4;
    const size_t smem_p = Sp * stage;
    PGRF_REQUIRE(smem_p <= 227 * 1024, "conv3d: %zu bytes of shared memory", smem_p);
    int per_sm = (int)((227 * 1024) / (smem_p + 2048));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
    const int n_vt = (int)grid.x;
    const long long total = (long long)grid.x * grid.y * grid.z;
    const unsigned gp = (unsigned)(total < 148LL * per_sm ? total : 148LL * per_sm);
#define PGRF_CONVP(N, SS, R)                                                                                                 \
  {                                                                                                                          \
    PGRF_CUDA(cudaFuncSetAttribute(conv3d_persist_kernel<N, SS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p)); \
    conv3d_persist_kernel<N, SS, R><<<gp, kPersistThreads, smem_p, st>>>(p, n_vt);                                           \
  }
#define PGRF_CONVP_S(N, R)                                                                     \
  case N:                                                                                      \
    if (Sp == 2) PGRF_CONVP(N, 2, R) else if (Sp == 3) PGRF_CONVP(N, 3, R)                       \
    else if (Sp == 4) PGRF_CONVP(N, 4, R) else PGRF_CONVP(N, 6, R)                               \
    break;
    if (pl.row) { switch (pl.NT) { PGRF_CONVP_S(16, true) PGRF_CONVP_S(32, true) PGRF_CONVP_S(64, true) PGRF_CONVP_S(128, true) } }
    else { switch (pl.NT) { PGRF_CONVP_S(16, false) PGRF_CONVP_S(32, false) PGRF_CONVP_S(64, false) PGRF_CONVP_S(128, false) } }
#undef PGRF_CONVP_S
#undef PGRF_CONVP
  } else {
#define PGRF_CONV(K, N, SS)                                                                                \
  {                                                                                                        \
    PGRF_CUDA(cudaFuncSetAttribute(K<N, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
    K<N, SS><<<grid, kConvThreads, smem, st>>>(p);                                                         \
  }
  if (pl.row) {
#define PGRF_CONV_S(N)                                                                                     \
  case N:                                                                                                  \
    if (S == 2) PGRF_CONV(conv3d_igemm_row_kernel, N, 2) else if (S == 3) PGRF_CONV(conv3d_igemm_row_kernel, N, 3) \
    else PGRF_CONV(conv3d_igemm_row_kernel, N, 4)                                                          \
    break;
    switch (pl.NT) { PGRF_CONV_S(16) PGRF_CONV_S(32) PGRF_CONV_S(64) PGRF_CONV_S(128) }
#undef PGRF_CONV_S
  } else {
#define PGRF_CONV_S(N)                                                                                     \
  case N:                                                                                                  \
    if (S == 3) PGRF_CONV(conv3d_igemm_kernel, N, 3) else if (S == 4) PGRF_CONV(conv3d_igemm_kernel, N, 4) \
    else PGRF_CONV(conv3d_igemm_kernel, N, 6)                                                              \
    break;
    switch (pl.NT) { PGRF_CONV_S(16) PGRF_CONV_S(32) PGRF_CONV_S(64) PGRF_CONV_S(128) }
#undef PGRF_CONV_S
  }
#undef PGRF_CONV
  }
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  if (pl.splits > 1) {
    conv3d_reduce_kernel<<<blocks_for(p.n_vox * (Cout / 8), 256), 256, 0, st>>>(p);
    count_launch();
    PGRF_CUDA(cudaGetLastError());
  }
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, float* yf,
                               int cout_real, int Cout, int B, int D, int H, int W, int act, float* ws, long long ws_floats, void* stream) {
  return conv3d_launch(xa, Ca, xb, Cb, wpk, bias, y, yf, cout_real, Cout, B, D, H, W, act, ws, ws_floats, false, nullptr, 1, stream);
}

// same with an optional residual input and the choice of the width padding (wrap / zeros)
extern "C" int pgrf_conv3d_ex_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y, float* yf,
                                  int cout_real, int Cout, int B, int D, int H, int W, int act, float* ws, long long ws_floats,
                                  const void* res, int wrap, void* stream) {
  return conv3d_launch(xa, Ca, xb, Cb, wpk, bias, y, yf, cout_real, Cout, B, D, H, W, act, ws, ws_floats, false, res, wrap, stream);
}

// Pointwise (1x1x1) convolution through the same pipeline: only the centre tap of the packed weights is walked.
extern "C" int pgrf_conv3d_pointwise_fwd(const void* xa, int Ca, const void* xb, int Cb, const void* wpk, const float* bias, void* y,
                                         float* yf, int cout_real, int Cout, int B, int D, int H, int W, int act, void* stream) {
  return conv3d_launch(xa, Ca, xb, Cb, wpk, bias, y, yf, cout_real, Cout, B, D, H, W, act, nullptr, 0, true, nullptr, 1, stream);
}

extern "C" int pgrf_conv3d_tapsum_fwd(const float* z, float bias, int B, int D, int H, int W, int act, float* out, void* stream) {
  PGRF_REQUIRE(z && out && B >= 1 && D >= 1 && H >= 1 && W >= 2, "conv3d_tapsum: bad arguments");
  const long long n = (long long)B * D * H * W;
  PGRF_REQUIRE_32BIT(n, "conv3d_tapsum");
  conv3d_tapsum_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(z, bias, B, D, H, W, act, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_scalar_fwd(const float* x, const float* w27_host, float bias, int B, int D, int H, int W, int act, float* out,
                                      void* stream) {
  PGRF_REQUIRE(x && w27_host && out && B >= 1 && D >= 1 && H >= 1 && W >= 2, "conv3d_scalar: bad arguments");
  Taps27 taps;
  for (int i = 0; i < 27; ++i) taps.w[i] = w27_host[i];
  const long long n = (long long)B * D * H * W;
  PGRF_REQUIRE_32BIT(n, "conv3d_scalar");
  conv3d_scalar_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, taps, bias, B, D, H, W, act, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_conv3d_cout1_fwd(const void* xa, int Ca, const void* xb, int Cb, const float* xf, const float* w, float bias, int B,
                                     int D, int H, int W, int act, float* out, void* stream) {
  PGRF_REQUIRE(w && out && (xf || (xa && Ca % 8 == 0 && Cb % 8 == 0)), "conv3d_cout1: bad arguments");
  const long long n = (long long)B * D * H * W;
  conv3d_cout1_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)xa, Ca, (const __nv_bfloat16*)xb, Cb, xf, w,
                                                                            bias, B, D, H, W, act, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_avgpool3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "avgpool3d: sizes must be even, C %% 8 == 0");
  const long long n = (long long)B * (D / 2) * (H / 2) * (W / 2) * (C / 8);
  PGRF_REQUIRE_32BIT(n, "avgpool3d");
  avgpool3d2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, D, H, W, C, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_upsample2d2_fwd(const void* x, int B, int H, int W, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && C % 8 == 0, "upsample2d: C %% 8 == 0");
  const long long n = (long long)B * (2 * H) * (2 * W) * (C / 8);
  PGRF_REQUIRE_32BIT(n, "upsample2d");
  upsample3d2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, 1, H, W, C, 1, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_channel_dot_upsample_fwd(const float* x, long long sb, long long sc, long long sh, long long sw, int B, int C, int H,
                                             int W, const float* w, float bias, int scale, int rectify, float* out, void* stream) {
  PGRF_REQUIRE(x && w && out && B >= 1 && C >= 1 && H >= 1 && W >= 1 && scale >= 1, "channel_dot_upsample: bad arguments");
  PGRF_REQUIRE(rectify >= 0 && rectify <= 2, "channel_dot_upsample: rectify=%d (0 none, 1 depth, 2 disparity)", rectify);
  const long long n = (long long)B * H * scale * W * scale;
  channel_dot_upsample_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, sb, sc, sh, sw, B, C, H, W, w, bias, scale, rectify, out);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

extern "C" int pgrf_upsample3d2_fwd(const void* x, int B, int D, int H, int W, int C, void* y, void* stream) {
  PGRF_REQUIRE(x && y && C % 8 == 0, "upsample3d: C %% 8 == 0");
  const long long n = (long long)B * (2 * D) * (2 * H) * (2 * W) * (C / 8);
  PGRF_REQUIRE_32BIT(n, "upsample3d");
  upsample3d2_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, B, D, H, W, C, 2, (__nv_bfloat16*)y);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

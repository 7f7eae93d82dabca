// Self-test of the tcgen05 building blocks: D[128 x N] = A[128 x K] * W[N x K]^T (bf16 in, fp32 out)
// through shared-memory descriptors, TMEM and the thread-per-row epilogue.  Used by tests/test_umma_gpu.py.
#include "common.cuh"
#include "render_device.cuh"
#include "umma.cuh"

namespace pgrf {

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Wm,
                                                               float* __restrict__ out, int K, int N, int swap_lbo_sbo) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  __shared__ __align__(8) uint64_t bar;
  unsigned char* As = smem;                       // [K/8][128][8] bf16
  unsigned char* Ws = smem + (size_t)K * 128 * 2; // [K/8][N][8] bf16
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_base, 256);
  if (tid == 0) mbar_init(&bar, 1);
  // stage operands: thread = row
  for (int c = 0; c < K / 8; ++c) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * K + c * 8 + i];
    umma::store_chunk(As, 128, c, tid, v);
  }
  for (int r = tid; r < N; r += 128)
    for (int c = 0; c < K / 8; ++c) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = Wm[(size_t)r * K + c * 8 + i];
      umma::store_chunk(Ws, N, c, r, v);
    }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_base;
  if (swap_lbo_sbo >= 4) {
    // diagnostic: ONE M = 64 MMA on rows 0..63, accumulator at lane offset (variant - 4) * 16; dump every lane.
    // Measured on B200 (tools/umma_m64_probe.py): rows 16q..16q+15 land in lanes 32q + (offset & 16) + 0..15 — every warp's lane
    // quadrant holds 16 rows, so two M = 64 MMAs interleave INSIDE each warp and cannot serve as two independent pipelines.
    if (tid < 128) {   // clear the accumulator columns first (zeros via tcgen05.st)
      uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int c = 0; c < N; c += 8) umma::st8(tbase + ((uint32_t)(warp * 32) << 16) + c, z);
      umma::wait_st();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
      const uint32_t idesc = umma::instr_desc_bf16(64, N);
      for (int k = 0; k < K; k += 16) {
        const uint64_t ad = umma::smem_desc(umma::smem_addr(As) + (k >> 3) * 128 * 16, 128 * 16, 128);
        const uint64_t bd = umma::smem_desc(umma::smem_addr(Ws) + (k >> 3) * N * 16, N * 16, 128);
        umma::mma_bf16(tbase + ((uint32_t)((swap_lbo_sbo - 4) * 16) << 16), ad, bd, idesc, k > 0);
      }
      umma::commit(&bar);
    }
    mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const uint32_t lb = tbase + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < N; n0 += 16) {
      float v[16];
      umma::ld16(lb + n0, v);
      for (int i = 0; i < 16; ++i) out[(size_t)tid * N + n0 + i] = v[i];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 256);
    return;
  }
  if (swap_lbo_sbo == 2) {
    // variant 2: A operand in TMEM columns [128, 128 + K/2): every thread writes its own row with tcgen05.st
    const uint32_t a_lane = tbase + ((uint32_t)(warp * 32) << 16) + 128;
    for (int k = 0; k < K; k += 16) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) r[i] = umma::pack2(A[(size_t)tid * K + k + 2 * i], A[(size_t)tid * K + k + 2 * i + 1]);
      umma::st8(a_lane + (k >> 1), r);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
      const uint32_t idesc = umma::instr_desc_bf16(128, N);
      for (int k = 0; k < K; k += 16) {
        const uint64_t bd = umma::smem_desc(umma::smem_addr(Ws) + (k >> 3) * N * 16, N * 16, 128);
        umma::mma_bf16_ts(tbase, tbase + 128 + (k >> 1), bd, idesc, k > 0);
      }
      umma::commit(&bar);
    }
  } else
  if (tid == 0) {
    if (!swap_lbo_sbo) {
      umma::gemm_issue(tbase, As, 128, Ws, N, N, K);
    } else {
      const uint32_t idesc = umma::instr_desc_bf16(128, N);
      for (int k = 0; k < K; k += 16) {
        const uint64_t ad = umma::smem_desc(umma::smem_addr(As) + (k >> 3) * 128 * 16, 128, 128 * 16);
        const uint64_t bd = umma::smem_desc(umma::smem_addr(Ws) + (k >> 3) * N * 16, 128, N * 16);
        umma::mma_bf16(tbase, ad, bd, idesc, k > 0);
      }
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_sync();
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  for (int n0 = 0; n0 < N; n0 += 16) {
    float v[16];
    umma::ld16(lane_base + n0, v);
    for (int i = 0; i < 16; ++i) out[(size_t)tid * N + n0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

}  // namespace pgrf

using namespace pgrf;

extern "C" int pgrf_umma_selftest(const float* A, const float* W, float* out, int K, int N, int variant, void* stream) {
  PGRF_REQUIRE(A && W && out, "umma_selftest: null pointer");
  PGRF_REQUIRE(K % 16 == 0 && K >= 16 && K <= 256 && N % 16 == 0 && N >= 16 && N <= 256, "umma_selftest: bad K=%d N=%d", K, N);
  const size_t smem = (size_t)K * 128 * 2 + (size_t)K * N * 2;
  PGRF_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, out, K, N, variant);
  count_launch();
  PGRF_CUDA(cudaGetLastError());
  return PGRF_OK;
}

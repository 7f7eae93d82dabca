// tcgen05 / TMEM building blocks (inline PTX, sm_100a) for the bf16 render path.
//
// Operand convention used throughout ("k-chunk-major", no swizzle, K-major):
//   a [rows x K] bf16 operand is stored as  [K/8 chunks][rows][8 elements]  i.e. element (r,k) lives at
//   byte offset (k/8)*rows*16 + r*16 + (k%8)*2.  One 8x(16 byte) core matrix = 8 consecutive rows of
//   one chunk = 128 contiguous bytes.  Shared-memory descriptor: SBO (stride between 8-row groups) =
//   128 B, LBO (stride between the two K chunks of one K=16 MMA) = rows*16 B.
//   A thread that owns row r writes/reads its 8 k-values of chunk j with ONE conflict-free 128-bit
//   access at (j*rows + r)*16 — this is what makes the one-thread-per-row epilogues cheap.
// Accumulators: D[128 x N] fp32 in TMEM, lane = row, column = n; warp w of a warpgroup reads lanes
// 32*(w%4)..+31 with tcgen05.ld.32x32b (thread = row).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace pgrf {
namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one full warp executes these) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors ----
// K-major, no swizzle. `lbo_bytes` = rows*16 for the k-chunk-major layout above, `sbo_bytes` = 128.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
// kind::f16, A = B = bf16, D = fp32, both K-major, M = 128
__host__ __device__ constexpr uint32_t instr_desc_bf16(int M, int N) {
  return (1u << 4)                      // D format: f32
         | (1u << 7) | (1u << 10)       // A, B format: bf16
         | ((uint32_t)(N >> 3) << 17)   // N
         | ((uint32_t)(M >> 4) << 24);  // M
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K=16 step; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from TENSOR MEMORY: row r = TMEM lane r, two bf16 per 32-bit column, 8 columns per K=16 step
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM, thread = lane (row), 8 consecutive 32-bit columns (= 16 bf16 of one K=16 step)
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// Full GEMM: D[128 x N] = A[128 x K] * W[N x K]^T with both operands k-chunk-major.  One thread.
//   a_rows / w_rows = row pitch (rows per chunk) of the operand buffers.
static __device__ __noinline__ void gemm_issue(uint32_t d_tmem, const void* A, int a_rows, const void* Wt, int w_rows, int N, int K,
                                        bool accumulate_first = false) {
  const uint32_t idesc = instr_desc_bf16(128, N);
  const uint32_t a_lbo = a_rows * 16, w_lbo = w_rows * 16;
  uint64_t ad = smem_desc(smem_addr(A), a_lbo, 128);
  uint64_t bd = smem_desc(smem_addr(Wt), w_lbo, 128);
  const uint64_t a_step = (uint64_t)((2 * a_lbo) >> 4), w_step = (uint64_t)((2 * w_lbo) >> 4);   // two k-chunks per K=16 MMA
  for (int k = 0; k < K; k += 16) {
    mma_bf16(d_tmem, ad, bd, idesc, (k > 0 || accumulate_first) ? 1u : 0u);
    ad += a_step;     // start-address field (bits 0..13, 16-byte units); operands never cross the 256 KB field range
    bd += w_step;
  }
}

__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld2(uint32_t taddr, float& v0, float& v1) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v0 = __uint_as_float(r0); v1 = __uint_as_float(r1);
}

// ---- TMEM -> registers: thread = row (lane of its warp's 32-lane quadrant), 16 / 32 consecutive columns ----
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- bf16 packing helpers for the k-chunk-major layout ----
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// write 8 consecutive k-values of row `row` into chunk `chunk` of an operand buffer with `rows` rows per chunk
__device__ __forceinline__ void store_chunk(void* buf, int rows, int chunk, int row, const float* v8) {
  uint4 q;
  q.x = pack2(v8[0], v8[1]); q.y = pack2(v8[2], v8[3]); q.z = pack2(v8[4], v8[5]); q.w = pack2(v8[6], v8[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(buf) + ((size_t)chunk * rows + row) * 16) = q;
}
__device__ __forceinline__ void load_chunk(const void* buf, int rows, int chunk, int row, float* v8) {
  const uint4 q = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(buf) + ((size_t)chunk * rows + row) * 16);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v8[2 * i] = f.x; v8[2 * i + 1] = f.y;
  }
}

}  // namespace umma
}  // namespace pgrf

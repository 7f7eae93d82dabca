// Shared host/device helpers for the panogrf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/panogrf_b200.h"

namespace pgrf {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define PGRF_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      pgrf::set_error(__VA_ARGS__);             \
      return PGRF_EINVAL;                       \
    }                                           \
  } while (0)

#define PGRF_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      pgrf::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return PGRF_ECUDA;                                                                   \
    }                                                                                      \
  } while (0)

// ---- fp32 constants exactly as torch sees python doubles cast to float ---------------------
#define PGRF_PI_F 3.14159265358979323846f
#define PGRF_HALF_PI_F 1.57079632679489661923f
#define PGRF_TWO_PI_F 6.28318530717958647692f
#define PGRF_PI_D 3.14159265358979323846

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

// streaming (evict-first) 128-bit store: the volume is written once and never re-read here
__device__ __forceinline__ void stcs4(float4* p, const float4& v) { __stcs(p, v); }

// fmodf(t, 2*pi) for the argument range the ERP mapping produces; exact (Sterbenz) for 0<=t<4*pi
__device__ __forceinline__ float fmod_two_pi(float t) {
  if (t >= 0.f && t < 2.f * PGRF_TWO_PI_F) return t >= PGRF_TWO_PI_F ? t - PGRF_TWO_PI_F : t;
  return fmodf(t, PGRF_TWO_PI_F);
}

// packed fp32 pair arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100a): same FLOP rate as two scalar ops, ONE issue slot
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

}  // namespace pgrf

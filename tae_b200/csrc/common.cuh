// common.cuh — shared host/device helpers for libtae_b200 (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tae_b200.h"

namespace tae {

// ---- host-side error plumbing -------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define TAE_CHECK_SHAPE(cond, ...)          \
  do {                                      \
    if (!(cond)) {                          \
      ::tae::set_error(__VA_ARGS__);        \
      return TAE_ERR_SHAPE;                 \
    }                                       \
  } while (0)

#define TAE_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::tae::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                       __LINE__);                                                         \
      return TAE_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define TAE_CHECK_LAUNCH()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::tae::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                       __FILE__, __LINE__);                                               \
      return TAE_ERR_CUDA;                                                                \
    }                                                                                     \
    ::tae::count_launch();                                                                \
  } while (0)

int num_sms();  // cached per process (current device)
// Work counters for dynamically scheduled persistent kernels: returns a zeroed {next item, workers done} pair in device
// memory (a ring of slots, one per launch; the kernel re-arms its slot when its last worker finishes), or NULL when
// dynamic scheduling is off (the default: tae_set_dynamic_scheduling / TAE_GEMM_DYNAMIC=1 turn it on) or allocation
// failed; callers then use their static work lists.
int* sched_counter_slot();
int set_dynamic_scheduling(int enable);

// ---- device helpers -----------------------------------------------------------------------
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// round-to-nearest-even fp32 -> bf16 -> fp32 (the autocast rounding point of the reference)
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  bf162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  bf162 t = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(t);
}

// exact (erf) GELU and its derivative, as nn.GELU() with approximate='none' (tae.py:96,102)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Epilogue-rate exact GELU.  q(x) = erfc(|x|/sqrt2)/2 from the Abramowitz-Stegun 7.1.26 form (|abs err| <= 1.5e-7 on
// erf, far below bf16 resolution) with ONE rcp and ONE exp2 shared between Phi and phi:
//   t = 1/(1 + p|x|/sqrt2),  e = e^{-x^2/2},  q = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) e / 2
//   gelu(x)  = x Phi(x)          = max(x, 0) - |x| q
//   gelu'(x) = Phi(x) + x phi(x) = 1/2 + copysign(1/2 - q, x) + x e / sqrt(2 pi)
__device__ __forceinline__ void gelu_qe(float x, float& q, float& e) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.3275911f * 0.70710678118654752440f, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.72134752044448170368f) * x));
  float poly = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  poly = fmaf(t, poly, 0.5f * 1.421413741f);
  poly = fmaf(t, poly, 0.5f * -0.284496736f);
  poly = fmaf(t, poly, 0.5f * 0.254829592f);
  q = (poly * t) * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  float q, e;
  gelu_qe(x, q, e);
  return fmaf(-fabsf(x), q, fmaxf(x, 0.f));
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  float q, e;
  gelu_qe(x, q, e);
  return fmaf(x * 0.39894228040143267794f, e, 0.5f + copysignf(0.5f - q, x));
}
__device__ __forceinline__ void gelu_and_grad_fast(float x, float& g, float& gp) {
  float q, e;
  gelu_qe(x, q, e);
  // Phi(x) = 1 - q (x > 0) or q (x <= 0);  gelu = x Phi,  gelu' = Phi + x phi  (two instructions fewer than the
  // max/copysign forms above; identical rounding behaviour to ~1 ulp of fp32)
  const float cdf = x > 0.f ? 1.0f - q : q;
  g = x * cdf;
  gp = fmaf(x * 0.39894228040143267794f, e, cdf);
}
// bf16 rounding of a PAIR through one packed convert: returns the two rounded values as fp32
__device__ __forceinline__ float2 round_bf16x2(float a, float b) {
  const uint32_t u = pack_bf16x2(a, b);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two lanes of fp32 math per issued instruction) -------------
// A pair lives in one 64-bit register; ptxas keeps the halves in an aligned register pair, so pack/unpack are free.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_bcast(float c) { return f2_pack(c, c); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// gelu_and_grad_fast on a PAIR, starting from the raw accumulator and bias pairs: h = bf16(acc + bias), then the same
// formulas and roundings as the scalar version (bit-identical results), with every FMA-pipe instruction packed: 13.5
// issued instructions per element instead of 20.5.  (Measured: by itself this does not change the fc1 GEMM's time —
// 0.496 vs 0.498 ms — the two output streams bound that kernel; it is what lets 8 epilogue warps carry the row-layout
// GELU epilogue.)  Returns bf16x2 bit patterns of gelu(h) and gelu'(h).
__device__ __forceinline__ void gelu_and_grad_pair(f32x2 acc, f32x2 bias, uint32_t& g_bf16x2, uint32_t& gp_bf16x2) {
  float s0, s1;
  f2_unpack(f2_add(acc, bias), s0, s1);
  const uint32_t hb = pack_bf16x2(s0, s1);
  const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xffff0000u);
  const f32x2 h = f2_pack(h0, h1);
  float d0, d1, t0, t1, a0, a1, e0, e1;
  f2_unpack(f2_fma(f2_pack(fabsf(h0), fabsf(h1)), f2_bcast(0.3275911f * 0.70710678118654752440f), f2_bcast(1.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  f2_unpack(f2_mul(f2_mul(h, f2_bcast(-0.72134752044448170368f)), h), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const f32x2 t = f2_pack(t0, t1), e = f2_pack(e0, e1);
  f32x2 poly = f2_fma(t, f2_bcast(0.5f * 1.061405429f), f2_bcast(0.5f * -1.453152027f));
  poly = f2_fma(t, poly, f2_bcast(0.5f * 1.421413741f));
  poly = f2_fma(t, poly, f2_bcast(0.5f * -0.284496736f));
  poly = f2_fma(t, poly, f2_bcast(0.5f * 0.254829592f));
  const f32x2 q = f2_mul(f2_mul(poly, t), e);
  float q0, q1, o0, o1;
  f2_unpack(q, q0, q1);
  f2_unpack(f2_fma(q, f2_bcast(-1.0f), f2_bcast(1.0f)), o0, o1);
  const f32x2 cdf = f2_pack(h0 > 0.f ? o0 : q0, h1 > 0.f ? o1 : q1);
  float g0, g1, p0, p1;
  f2_unpack(f2_mul(h, cdf), g0, g1);
  f2_unpack(f2_fma(f2_mul(h, f2_bcast(0.39894228040143267794f)), e, cdf), p0, p1);
  g_bf16x2 = pack_bf16x2(g0, g1);
  gp_bf16x2 = pack_bf16x2(p0, p1);
}

// 128-bit streaming loads/stores (read-once data: do not allocate in L1)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_nc_f4(const void* p) {
  uint4 r = ld_nc_v4(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z),
                     __uint_as_float(r.w));
}

#endif  // __CUDACC__

}  // namespace tae

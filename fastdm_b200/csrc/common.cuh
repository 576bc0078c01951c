// Shared host/device helpers for libfastdm_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fastdm_b200.h"

namespace fdm {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int require_sm100();   // FDM_OK or FDM_ERR_ARCH for the current device
int num_sms();         // SM count of the current device (cached)

#define FDM_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::fdm::set_error(__VA_ARGS__);           \
      return FDM_ERR_ARG;                      \
    }                                          \
  } while (0)

#define FDM_CUDA(call)                                         \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::fdm::cuda_fail(e__, #call); \
  } while (0)

#define FDM_LAUNCH_CHECK(name)                                   \
  do {                                                           \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::fdm::cuda_fail(e__, name);  \
  } while (0)

// ---- 16-byte vector of raw bits ---------------------------------------------------------------
struct alignas(16) U128 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U128 ldg128_stream(const void* p) {
  U128 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ U128 ldg128(const void* p) {
  return *reinterpret_cast<const U128*>(p);
}
__device__ __forceinline__ void stg128(void* p, const U128& v) {
  *reinterpret_cast<U128*>(p) = v;
}
__device__ __forceinline__ void stg64(void* p, uint32_t a, uint32_t b) {
  *reinterpret_cast<uint2*>(p) = make_uint2(a, b);
}

// ---- dtype traits: load 8 elements as floats ---------------------------------------------------
template <typename T>
struct Elem;
template <>
struct Elem<__nv_bfloat16> {
  static constexpr int kId = FDM_BF16;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Elem<__half> {
  static constexpr int kId = FDM_F16;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
};
template <>
struct Elem<float> {
  static constexpr int kId = FDM_F32;
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
};

// bf16 / f16 pair <-> packed 32-bit
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float f16lo(uint32_t u) {
  return __half2float(__ushort_as_half((unsigned short)(u & 0xffff)));
}
__device__ __forceinline__ float f16hi(uint32_t u) {
  return __half2float(__ushort_as_half((unsigned short)(u >> 16)));
}

template <typename T>
__device__ __forceinline__ void unpack8(const U128& v, float (&f)[8]);
template <>
__device__ __forceinline__ void unpack8<__nv_bfloat16>(const U128& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
template <>
__device__ __forceinline__ void unpack8<__half>(const U128& v, float (&f)[8]) {
  f[0] = f16lo(v.x); f[1] = f16hi(v.x); f[2] = f16lo(v.y); f[3] = f16hi(v.y);
  f[4] = f16lo(v.z); f[5] = f16hi(v.z); f[6] = f16lo(v.w); f[7] = f16hi(v.w);
}
template <typename T>
__device__ __forceinline__ U128 pack8(const float (&f)[8]);
template <>
__device__ __forceinline__ U128 pack8<__nv_bfloat16>(const float (&f)[8]) {
  U128 r;
  r.x = pack_bf16(f[0], f[1]); r.y = pack_bf16(f[2], f[3]);
  r.z = pack_bf16(f[4], f[5]); r.w = pack_bf16(f[6], f[7]);
  return r;
}
template <>
__device__ __forceinline__ U128 pack8<__half>(const float (&f)[8]) {
  U128 r;
  r.x = pack_f16(f[0], f[1]); r.y = pack_f16(f[2], f[3]);
  r.z = pack_f16(f[4], f[5]); r.w = pack_f16(f[6], f[7]);
  return r;
}

// round-trip a float through T (the "rounded to the tensor dtype" step of the torch oracle)
template <typename T>
__device__ __forceinline__ float round_to(float v) {
  return Elem<T>::to_f(Elem<T>::from_f(v));
}

// ---- fp8 e4m3 conversion: two floats -> two e4m3 bytes (RNE, saturate-to-finite) ---------------
__device__ __forceinline__ uint16_t cvt_e4m3x2(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t cvt_e4m3x4(float a, float b, float c, float d) {
  return (uint32_t)cvt_e4m3x2(a, b) | ((uint32_t)cvt_e4m3x2(c, d) << 16);
}

// ---- reductions --------------------------------------------------------------------------------
template <int W = 32>
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int W = 32>
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int W = 32>
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- activations (fp32 in / fp32 out; callers round to the tensor dtype) ------------------------
// Both GELUs are evaluated with two MUFU ops (ex2 + rcp) and a handful of FMAs instead of libm's
// tanhf / erff (25-40 instructions with branches): in a GEMM epilogue the four epilogue warps have to
// push 32768 activations per 128x256 tile through this code while the tensor cores work on the next
// tile, and the libm versions made the ff1 / proj_mlp GEMMs 2.5x slower than the plain GEMM.
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// F.gelu(approximate="tanh") = 0.5 x (1 + tanh(u)) = x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3).
// Relative error ~3e-7 (ex2.approx + rcp.approx), far below a bf16 ulp.
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.79788456080286535588f * 2.0f * 1.44269504088896340736f;  // 2 sqrt(2/pi) log2(e)
  const float x2 = x * x;
  const float w = x * fmaf(x2, k * 0.044715f, k);  // 2u log2(e)
  return x * fast_rcp(1.0f + fast_ex2(-w));
}
// F.gelu (exact) = x * Phi(x), Phi(x) = 0.5 erfc(-x/sqrt2). erfc(z) for z >= 0 by Abramowitz-Stegun
// 7.1.26 (|error| <= 1.5e-7 absolute): erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) e^{-z^2},
// t = 1 / (1 + p z).
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float half_erfc = 0.5f * poly * t * fast_ex2(-z * z * 1.44269504088896340736f);  // 0.5 erfc(|x|/sqrt2)
  const float phi = x >= 0.f ? 1.0f - half_erfc : half_erfc;
  return x * phi;
}

// 32-byte global accesses (sm_100: LDG/STG.E.ENL2.256): two 16-byte vectors at a 32-byte aligned address. One
// instruction covers a whole 32-byte sector per lane -- half the LSU requests of two 16-byte accesses, and stores
// never write partial sectors.
__device__ __forceinline__ void ldg256(const void* p, U128& lo, U128& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const U128& lo, const U128& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// ---- packed bf16 arithmetic --------------------------------------------------------------------------------
// The reference's bf16 tensor ops are "fp32 op, round to bf16". For two bf16 operands that is exactly what the
// native packed instructions compute: a product of two 8-bit significands is exact in fp32, so RN_bf16(fp32
// product) = RN_bf16(exact product) = mul.rn.bf16x2; for a sum the fp32 add is exact whenever the operands'
// exponents are within 16 of each other, and beyond that the small operand is below a quarter ulp of the large
// one in bf16, so both orders of rounding return the large operand (or its neighbour by the same rule).
// Two elements per instruction and no separate rounding step: the element-wise bf16 chains drop from ~10 to
// ~3.5 instructions per element.
__device__ __forceinline__ uint32_t bmul2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t badd2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// interleaved rotation of 4 packed pairs (x[p] = {x_2p, x_2p+1}) by the cache entries c/s (4 bf16 each):
//   o_2p = T(T(x_2p c) - T(x_2p+1 s)),  o_2p+1 = T(T(x_2p+1 c) + T(x_2p s))      (torch/rotemb.py:40-48)
__device__ __forceinline__ void rope4_bf16(uint32_t (&x)[4], uint2 craw, uint2 sraw) {
  const uint32_t cw[2] = {craw.x, craw.y}, sw[2] = {sraw.x, sraw.y};
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t sel = (p & 1) ? 0x3232u : 0x1010u;                    // duplicate the high / low bf16
    const uint32_t cc = __byte_perm(cw[p >> 1], 0u, sel);                // { c,  c}
    const uint32_t ss = __byte_perm(sw[p >> 1], 0u, sel) ^ 0x00008000u;  // {-s, +s}
    const uint32_t xs = __byte_perm(x[p], 0u, 0x1032u);                  // {x_2p+1, x_2p}
    x[p] = badd2(bmul2(x[p], cc), bmul2(xs, ss));
  }
}
// T(T(f * rs) * w) for 8 values, packed result (rs is fp32: that product is an fp32 multiply + rounding)
__device__ __forceinline__ void norm_scale8_bf16(const float (&f)[8], float rs, const U128& w, uint32_t (&o)[4]) {
  const uint32_t wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int p = 0; p < 4; ++p) o[p] = bmul2(pack_bf16(f[2 * p] * rs, f[2 * p + 1] * rs), wv[p]);
}


}  // namespace fdm

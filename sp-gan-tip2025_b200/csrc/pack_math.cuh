// Packed fp32x2 arithmetic (FFMA2 / FMUL2 on sm_100a), pairwise 16-bit operand splits and the coordinate encoding shared by the
// operand producers (structure.cu) and the fused gather-GEMM (sphere_umma.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace {

constexpr int SPV_C = 256;   // main K columns per tap of the 256 + 3 channel spherical conv
constexpr int SPV_LD = 264;  // floats per (group, pixel) row of the repacked gather source: 259 flat-concat channels + padding

__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// (v0, v1) -> packed 16-bit hi pair and lo pair (lo = round(v - hi)), two values per conversion instruction.
template <bool kF16>
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  if (kF16) {
    v0 = fminf(fmaxf(v0, -65504.f), 65504.f);
    v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
    hi = hb;
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// not inlined: the trigonometric slow paths would sit nine times inside the hot loop, which takes them for one group in 32
__device__ __noinline__ float encode_coord(float v, int kind) {
  if (kind == 1) return tanhf(v);
  if (kind == 2) return cosf(v * 3.14159274101257324f);
  if (kind == 3) return sinf(v * 3.14159274101257324f);
  return v;
}


}  // namespace

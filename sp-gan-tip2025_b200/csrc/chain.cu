// Channels-last links of the fused texture-synthesiser chain (inference path).
//
// The tcgen05 GEMM consumes a channels-last 16-bit operand and its accumulator is channels-last by construction
// (TMEM lane = pixel, column = output channel), so between two convs of the texture synthesiser nothing needs the
// NCHW fp32 tensors the reference's modules exchange (models/spgan/spgan.py:924-978).  These kernels close the chain:
//   upblur_pack : polyphase planes of the transposed conv (NHWC fp32, written by the GEMM epilogue) -> interleave + 3x3
//                 FIR (Blur, models/ops.py:617-622) + noise + bias + leaky-ReLU (models/ops.py:784, fused_act.py:56-64)
//                 -> multiplied by the NEXT conv's style modulation -> that conv's packed hi/lo operand.  One read of the
//                 planes, one write of the operand; the fp32 activation and the separate pack pass disappear.
//   rgb_tail    : sums the ToRGB partial sums the GEMM epilogue produced (one slot per N tile and epilogue half) in a
//                 fixed order, adds the ToRGB bias and the upsampled skip (models/spgan_ops.py:1563-1586).
// Roofline: HBM.  upblur_pack moves 4 B/element in and 4 B/element out (2 x 16-bit planes).
#include "umma_common.cuh"

namespace {

constexpr int UP_THREADS = 256;
constexpr int UP_ROWS = 16;  // output rows per CTA band

struct UpPackParams {
  int C, C4, nsub;      // channels, float4 lanes per pixel, column pairs per CTA
  int Hq, Wq;           // polyphase plane size
  int zh, zw;           // interleaved image size (cropped transposed-conv output)
  int oh, ow;           // output size = (zh - 2, zw - 2)
  int Cp;               // leading dimension of the packed operand
  int64_t pk_rows;      // rows per 16-bit plane
  float alpha, scale;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x += w * v.x;
  a.y += w * v.y;
  a.z += w * v.z;
  a.w += w * v.w;
}

// Thread = 4 channels x 2 adjacent output columns, sliding down a band of UP_ROWS output rows with a 3-row window in
// registers: 4 float4 loads per 2 outputs (the L1 serves the column overlap between neighbouring threads), all accesses
// 512 B contiguous per warp.
template <bool kF16>
__global__ void __launch_bounds__(UP_THREADS) upblur_pack_kernel(uint16_t* __restrict__ out, const float* __restrict__ pp,
                                                                const float* __restrict__ kernel,
                                                                const float* __restrict__ noise,
                                                                const float* __restrict__ noise_w,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ next_mul, UpPackParams P) {
  const int c4 = threadIdx.x % P.C4;
  const int sub = threadIdx.x / P.C4;
  const int ox0 = (blockIdx.x * P.nsub + sub) * 2;
  const int oy0 = blockIdx.y * UP_ROWS;
  const int b = blockIdx.z;
  if (sub >= P.nsub || ox0 >= P.ow) return;
  float kf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) kf[i] = __ldg(kernel + (2 - i / 3) * 3 + (2 - i % 3));  // upfirdn2d flips the kernel
  const int c = 4 * c4;
  const float4 bv = bias ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mv = next_mul ? ld4(next_mul + (int64_t)b * P.C + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const int64_t Q = (int64_t)P.Hq * P.Wq;
  const float* base = pp + (int64_t)b * 4 * Q * P.C + c;
  // interleaved pixel (Y, X) lives in plane (Y & 1) * 2 + (X & 1) at (Y >> 1, X >> 1)
  auto load_row = [&](int Y, float4* r) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int X = ox0 + j;
      if (Y < P.zh && X < P.zw)
        r[j] = ld4(base + ((int64_t)((Y & 1) * 2 + (X & 1)) * Q + (int64_t)(Y >> 1) * P.Wq + (X >> 1)) * P.C);
      else
        r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float4 win[3][4];
  load_row(oy0, win[0]);
  load_row(oy0 + 1, win[1]);
  const int rows = min(UP_ROWS, P.oh - oy0);
  const bool two = ox0 + 1 < P.ow;
  uint16_t* out_lo = out + P.pk_rows * (int64_t)P.Cp;
  // one output row: (w0, w1, w2) are the window rows oy, oy + 1, oy + 2; the three call sites below rotate the roles so
  // that the window stays in registers (compile-time indices)
  auto step = [&](int r, float4* w0, float4* w1, float4* w2) {
    const int oy = oy0 + r;
    load_row(oy + 2, w2);
    float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        fma4(acc[j], kf[kx], w0[kx + j]);
        fma4(acc[j], kf[3 + kx], w1[kx + j]);
        fma4(acc[j], kf[6 + kx], w2[kx + j]);
      }
    }
    const int64_t prow = ((int64_t)b * P.oh + oy) * P.ow + ox0;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 1 && !two) break;
      const float nz = noise ? nw * __ldg(noise + prow + j) : 0.f;
      float v[4] = {acc[j].x + bv.x + nz, acc[j].y + bv.y + nz, acc[j].z + bv.z + nz, acc[j].w + bv.w + nz};
      const float m[4] = {mv.x, mv.y, mv.z, mv.w};
      uint16_t h[4], l[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = (v[k] > 0.f ? v[k] : v[k] * P.alpha) * P.scale;
        split16<kF16>(v[k] * m[k], h[k], l[k]);
      }
      const int64_t off = (prow + j) * P.Cp + c;
      *reinterpret_cast<uint2*>(out + off) = make_uint2(pack2x16(h[0], h[1]), pack2x16(h[2], h[3]));
      *reinterpret_cast<uint2*>(out_lo + off) = make_uint2(pack2x16(l[0], l[1]), pack2x16(l[2], l[3]));
    }
  };
  for (int r = 0; r < rows; r += 3) {
    step(r, win[0], win[1], win[2]);
    if (r + 1 < rows) step(r + 1, win[1], win[2], win[0]);
    if (r + 2 < rows) step(r + 2, win[2], win[0], win[1]);
  }
}

__global__ void __launch_bounds__(256) rgb_tail_kernel(float* __restrict__ out, const float* __restrict__ part,
                                                       const float* __restrict__ bias, const float* __restrict__ skip,
                                                       int slots, int64_t n, int rgb_n, int64_t plane) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < slots; ++s) a += __ldg(part + (int64_t)s * n + i);
    const int j = (int)((i / plane) % rgb_n);
    a += bias ? __ldg(bias + j) : 0.f;
    if (skip) a += __ldg(skip + i);
    out[i] = a;
  }
}

}  // namespace

extern "C" int spgan_upblur_pack(uint16_t* out, const float* pp, const float* kernel, const float* noise,
                                 const float* noise_w, const float* bias, const float* next_mul, int64_t batch,
                                 int channels, int zh, int zw, int Hq, int Wq, int Cp, int64_t out_rows, int fmt, float alpha,
                                 float scale, void* stream) {
  SPGAN_CHECK_ARG(batch >= 0 && channels >= 0 && zh >= 0 && zw >= 0, "spgan_upblur_pack: negative size");
  SPGAN_CHECK_ARG(Hq * 2 >= zh && Wq * 2 >= zw, "spgan_upblur_pack: polyphase planes %dx%d too small for %dx%d", Hq, Wq, zh, zw);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_upblur_pack: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  const int oh = zh - 2, ow = zw - 2;
  if (batch * channels == 0 || oh <= 0 || ow <= 0) return 0;
  SPGAN_CHECK_ARG(out && pp && kernel, "spgan_upblur_pack: null pointer");
  SPGAN_CHECK_ARG((noise == nullptr) == (noise_w == nullptr), "spgan_upblur_pack: noise and noise_w go together");
  SPGAN_CHECK_ARG(channels % 4 == 0 && UP_THREADS % (channels / 4) == 0,
                  "spgan_upblur_pack: channels=%d must be 4 * a divisor of %d", channels, UP_THREADS);
  SPGAN_CHECK_ARG(Cp == channels, "spgan_upblur_pack: the packed operand must have exactly %d columns, got %d", channels, Cp);
  SPGAN_CHECK_ARG(out_rows >= batch * oh * ow, "spgan_upblur_pack: packed operand has too few rows");
  SPGAN_CHECK_ARG(batch <= 65535, "spgan_upblur_pack: batch %lld > 65535", (long long)batch);
  SPGAN_CHECK_ARG(((((uintptr_t)out) | ((uintptr_t)pp)) & 15) == 0, "spgan_upblur_pack: pointers must be 16-byte aligned");
  UpPackParams P;
  P.C = channels;
  P.C4 = channels / 4;
  P.nsub = UP_THREADS / P.C4;
  P.Hq = Hq;
  P.Wq = Wq;
  P.zh = zh;
  P.zw = zw;
  P.oh = oh;
  P.ow = ow;
  P.Cp = Cp;
  P.pk_rows = out_rows;
  P.alpha = alpha;
  P.scale = scale;
  dim3 grid((ow + 2 * P.nsub - 1) / (2 * P.nsub), (oh + UP_ROWS - 1) / UP_ROWS, (unsigned)batch);
  if (fmt)
    upblur_pack_kernel<true><<<grid, UP_THREADS, 0, (cudaStream_t)stream>>>(out, pp, kernel, noise, noise_w, bias, next_mul, P);
  else
    upblur_pack_kernel<false><<<grid, UP_THREADS, 0, (cudaStream_t)stream>>>(out, pp, kernel, noise, noise_w, bias, next_mul, P);
  SPGAN_CHECK_LAUNCH("spgan_upblur_pack");
  return 0;
}

extern "C" int spgan_rgb_tail(float* out, const float* part, int slots, const float* bias, const float* skip,
                              int64_t batch, int rgb_n, int64_t plane, void* stream) {
  SPGAN_CHECK_ARG(batch >= 0 && rgb_n >= 1 && plane >= 0 && slots >= 1, "spgan_rgb_tail: bad size");
  const int64_t n = batch * rgb_n * plane;
  if (n == 0) return 0;
  SPGAN_CHECK_ARG(out && part, "spgan_rgb_tail: null pointer");
  rgb_tail_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, part, bias, skip, slots, n, rgb_n, plane);
  SPGAN_CHECK_LAUNCH("spgan_rgb_tail");
  return 0;
}

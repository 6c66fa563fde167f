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

// (v0, v1) -> packed 16-bit hi pair and lo pair (lo = round(v - hi)), two values per conversion instruction.
template <bool kF16>
__device__ __forceinline__ void split_pair16(float v0, float v1, uint32_t& hi, uint32_t& lo) {
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

__device__ __forceinline__ unsigned long long pk2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void fma2_acc(unsigned long long& acc, unsigned long long w, float a, float b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(pk2(a, b)), "l"(w));
}

// Thread = 4 channels x 2 adjacent output columns, sliding down a band of UP_ROWS output rows with a FOUR-row window in
// registers: the row needed by the NEXT output row is requested before the current one is computed (the three-row version
// had 32 KB in flight per SM at 16 resident warps and sat at ~45 % of the HBM roofline, ncu long-scoreboard 3.4 per issue);
// 4 float4 loads per 2 outputs (the L1 serves the column overlap between neighbouring threads), all accesses 512 B
// contiguous per warp; the 3x3 FIR runs on packed FFMA2.
template <bool kF16>
__global__ void __launch_bounds__(UP_THREADS, 2) upblur_pack_kernel(uint16_t* __restrict__ out, const float* __restrict__ pp,
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
  unsigned long long kf[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float k = __ldg(kernel + (2 - i / 3) * 3 + (2 - i % 3));  // upfirdn2d flips the kernel
    kf[i] = pk2(k, k);
  }
  const int c = 4 * c4;
  const float4 bv = bias ? ld4(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 mv = next_mul ? ld4(next_mul + (int64_t)b * P.C + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const int64_t Q = (int64_t)P.Hq * P.Wq;
  const float* base = pp + (int64_t)b * 4 * Q * P.C + c;
  // interleaved pixel (Y, X) lives in plane (Y & 1) * 2 + (X & 1) at (Y >> 1, X >> 1).  All addresses are running pointers:
  // the column part is fixed per thread, even rows walk planes 0/1 and odd rows planes 2/3 with one pointer each (the band
  // starts on an even row), outputs and noise advance by one image row per step.  (Recomputing the 64-bit addresses per row
  // cost ~60 of the ~280 instructions per step and made the kernel issue-bound at 60 % of the HBM roofline.)
  int col_off[4];
  bool col_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int X = ox0 + j;
    col_ok[j] = X < P.zw;
    col_off[j] = (int)(((int64_t)(X & 1) * Q + (X >> 1)) * P.C);  // < 2 * Q * C elements: the host checks it fits 31 bits
  }
  const int64_t rstep = (int64_t)P.Wq * P.C;
  const float* pe = base + (int64_t)(oy0 >> 1) * rstep;  // next even row to load
  const float* po = pe + 2 * Q * P.C;                    // next odd row to load
  int Yn = oy0;                                          // index of the next row to load
  auto load_next = [&](bool odd, float4* r) {
    const float* rp = odd ? po : pe;
    const bool row_ok = Yn < P.zh;
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] = (row_ok && col_ok[j]) ? ld4(rp + col_off[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (odd) po += rstep;
    else pe += rstep;
    ++Yn;
  };
  float4 win[4][4];
  load_next(false, win[0]);
  load_next(true, win[1]);
  load_next(false, win[2]);
  const int rows = min(UP_ROWS, P.oh - oy0);
  const bool two = ox0 + 1 < P.ow;
  const int64_t prow0 = ((int64_t)b * P.oh + oy0) * P.ow + ox0;
  const float* nzp = noise ? noise + prow0 : nullptr;
  uint16_t* op = out + prow0 * P.Cp + c;
  uint16_t* op_lo = op + P.pk_rows * (int64_t)P.Cp;
  const int64_t ostep = (int64_t)P.ow * P.Cp;
  // epilogue constants as packed pairs: bias, alpha, scale * next-layer modulation.  leaky-ReLU(v) = max(v, alpha v) for
  // alpha <= 1 (min otherwise): a multiply and a min/max instead of compare + select + multiply
  const unsigned long long b01 = pk2(bv.x, bv.y), b23 = pk2(bv.z, bv.w);
  const unsigned long long al2 = pk2(P.alpha, P.alpha);
  const unsigned long long sm01 = pk2(P.scale * mv.x, P.scale * mv.y), sm23 = pk2(P.scale * mv.z, P.scale * mv.w);
  const bool use_max = P.alpha <= 1.f;
  auto act_pair = [&](unsigned long long v2, unsigned long long bias2, unsigned long long nz2, unsigned long long sm2, uint32_t& h,
                      uint32_t& l) {
    unsigned long long t;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(v2), "l"(bias2));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(t), "l"(nz2));
    unsigned long long ta;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(ta) : "l"(t), "l"(al2));
    float x0, x1, y0, y1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(t));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(ta));
    x0 = use_max ? fmaxf(x0, y0) : fminf(x0, y0);
    x1 = use_max ? fmaxf(x1, y1) : fminf(x1, y1);
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk2(x0, x1)), "l"(sm2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(r));
    split_pair16<kF16>(x0, x1, h, l);
  };
  // one output row: (w0, w1, w2) are the window rows oy, oy + 1, oy + 2, wn receives row oy + 3 (odd for even r: the band
  // starts on an even row); the four call sites below rotate the roles so that the window stays in registers
  auto step = [&](int r, bool next_odd, float4* w0, float4* w1, float4* w2, float4* wn) {
    if (r + 1 < rows) load_next(next_odd, wn);
    unsigned long long acc[2][2] = {{0ull, 0ull}, {0ull, 0ull}};  // [column][channel pair], +0.0f bit patterns
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        fma2_acc(acc[j][0], kf[kx], w0[kx + j].x, w0[kx + j].y);
        fma2_acc(acc[j][1], kf[kx], w0[kx + j].z, w0[kx + j].w);
        fma2_acc(acc[j][0], kf[3 + kx], w1[kx + j].x, w1[kx + j].y);
        fma2_acc(acc[j][1], kf[3 + kx], w1[kx + j].z, w1[kx + j].w);
        fma2_acc(acc[j][0], kf[6 + kx], w2[kx + j].x, w2[kx + j].y);
        fma2_acc(acc[j][1], kf[6 + kx], w2[kx + j].z, w2[kx + j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 1 && !two) break;
      const float nz = nzp ? nw * __ldg(nzp + j) : 0.f;
      const unsigned long long nz2 = pk2(nz, nz);
      uint32_t h0, l0, h1, l1;
      act_pair(acc[j][0], b01, nz2, sm01, h0, l0);
      act_pair(acc[j][1], b23, nz2, sm23, h1, l1);
      *reinterpret_cast<uint2*>(op + (int64_t)j * P.Cp) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(op_lo + (int64_t)j * P.Cp) = make_uint2(l0, l1);
    }
    if (nzp) nzp += P.ow;
    op += ostep;
    op_lo += ostep;
  };
  for (int r = 0; r < rows; r += 4) {
    step(r, true, win[0], win[1], win[2], win[3]);
    if (r + 1 < rows) step(r + 1, false, win[1], win[2], win[3], win[0]);
    if (r + 2 < rows) step(r + 2, true, win[2], win[3], win[0], win[1]);
    if (r + 3 < rows) step(r + 3, false, win[3], win[0], win[1], win[2]);
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
  SPGAN_CHECK_ARG((int64_t)2 * Hq * Wq * channels < (1LL << 31), "spgan_upblur_pack: polyphase planes too large for 32-bit column offsets");
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

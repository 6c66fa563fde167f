// Epilogue of the tcgen05 implicit-GEMM kernels (conv_umma.cu, sphere_umma.cu): parameter blocks and the per-tile routine
// that turns one 128 x kBlockN fp32 accumulator in TMEM into the kernel's output sinks.
#pragma once
#include "umma_common.cuh"

namespace {

struct GemmParams {
  int32_t B;           // samples
  int32_t rows;        // M extent: B * Hl * Wl lattice points (tiled A loads) or B * My * Mx outputs (im2col A loads)
  int32_t Hl, Wl;      // row decode: points per sample = Hl * Wl, Wl per row (= My, Mx in im2col mode)
  int32_t My, Mx;      // valid lattice extent
  int32_t im2col;      // 1: the A tile is fetched with TMA im2col loads (no lattice waste), 0: flat row offsets
  int32_t img_lo;      // im2col: image offset of the lo plane (phases * B)
  int32_t tap_ox[SPGAN_MAX_TAPS], tap_oy[SPGAN_MAX_TAPS], tap_img[SPGAN_MAX_TAPS];  // im2col: offsets, phase * B
  int32_t Cout, out_H, out_W;
  int32_t out_stride, out_off_y, out_off_x;
  int64_t out_cstride;     // elements between output channels
  int32_t ntaps, kblocks;  // kblocks = ceil(kp / 64)
  int32_t last_ksteps;     // K = 16 steps in the last block of a tap (1..4)
  int32_t k2blocks;        // 64-wide blocks of the second K segment (A2 x W2, rows = the tile's M rows), 0 = none
  int32_t tap_off[SPGAN_MAX_TAPS];
  int32_t m_tiles, n_tiles;
  float out_scale;
  int32_t act;
  float act_alpha, act_gain;
  int32_t a_f16, b_f16;    // operand formats: 0 = bf16 planes, 1 = fp16 planes (instruction descriptor A / B format)
  // ---- output sinks (any combination; y may be null)
  int32_t y_nhwc;          // 0: y is NCHW fp32 (out_cstride between channels), 1: y is NHWC fp32 (channel fastest)
  int64_t y_bstride;       // NHWC: elements between samples
  int64_t pk_rows;         // packed sink: rows per 16-bit plane (the lo plane starts pk_rows * pk_cols elements later)
  int32_t pk_cols;         // packed sink: leading dimension (the next conv's Cp)
  int32_t pk_f16;          // packed sink: 0 = bf16 hi/lo, 1 = fp16 hi/lo
  int32_t rgb_n;           // ToRGB sink: number of RGB channels (3) or 0
  int64_t res_bstride;     // residual_nhwc: elements between samples
};

// Everything the epilogue may read or write besides the accumulator.
struct GemmSinks {
  float* y;                 // fp32 output (NCHW or NHWC), may be null
  const float* out_mul;     // (B, Cout) demodulation
  const float* noise;       // (B, out_H, out_W)
  const float* noise_w;     // (1)
  const float* bias;        // (Cout)
  const float* residual;    // same layout as an NCHW y
  const float* residual_nhwc;  // channels-last residual (b, Y, X, o) added after the activation (general sinks only)
  uint16_t* y_packed;       // the next conv's A operand [2][pk_rows][pk_cols], row (b*out_H + Y)*out_W + X; may be null
  const float* next_mul;    // (B, Cout) style modulation of the next conv, folded into y_packed; may be null
  const float* rgb_w;       // (B, rgb_n, Cout) per-sample modulated ToRGB weights; null = sink off
  float* rgb_part;          // (2 * n_tiles, B, rgb_n, out_H*out_W) partial sums, one slot per (N tile, epilogue half)
};


// One accumulator tile -> sinks.  Called by the eight epilogue warps (warp w may touch TMEM lanes 32*(w & 3)..+31; the two
// warps of a lane quarter take alternate blocks of 32 accumulator columns).  With one warp per scheduler the epilogue of
// the few-tap passes (parity passes of the transposed conv, 1x1 shortcuts) was issue-latency bound at ~56 instructions
// per column; here the per-column work is a multiply, a pointer bump and a store, the per-channel factors come in as
// 128-bit loads, and the rare terms (noise, bias, activation, residual) are behind one warp-uniform branch.
// m0: first M row of this CTA's 128-row accumulator; n_tile: index of the kBlockN-wide column tile.
// tfull_bar / aphase: the accumulator-complete barrier of this TMEM stage and its parity; tmem_acc: column base of the stage.
template <int kBlockN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& gp, const GemmSinks& sk, int m0, int n_tile, int warp,
                                                   int lane, uint32_t tfull_bar, uint32_t aphase, uint32_t tmem_acc) {
  float* __restrict__ y = sk.y;
  const float* __restrict__ out_mul = sk.out_mul;
  const float* __restrict__ noise = sk.noise;
  const float* __restrict__ noise_w = sk.noise_w;
  const float* __restrict__ bias = sk.bias;
  const float* __restrict__ residual = sk.residual;
  {
    const int quarter = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;  // 0: even 32-column blocks, 1: odd ones
    const int plane = gp.Hl * gp.Wl;
    const int64_t ostride_c = gp.out_cstride;
    const int64_t oplane = (int64_t)gp.out_H * gp.out_W;
    const bool has_noise = noise != nullptr && noise_w != nullptr;
    const float nw = has_noise ? __ldg(noise_w) : 0.f;
    const bool plain = !has_noise && bias == nullptr && residual == nullptr && !gp.act;
    const bool vec_ok = (gp.Cout & 3) == 0;  // per-channel rows start 16-byte aligned
    // legacy sink: one NCHW fp32 tensor.  Anything else (channels-last fp32, the next conv's packed operand, ToRGB
    // partial sums) goes through the general path below, which the host only selects when Cout is a multiple of 32.
    const bool nchw_only = y != nullptr && !gp.y_nhwc && sk.y_packed == nullptr && sk.rgb_w == nullptr && sk.residual_nhwc == nullptr;
    {
      const int n0 = n_tile * kBlockN;
      int n_eff = gp.Cout - n0;
      n_eff = n_eff > kBlockN ? kBlockN : ((n_eff + 15) & ~15);
      // decode this thread's lattice point
      const int p = m0 + quarter * 32 + lane;
      bool valid = p < gp.rows;
      int b = 0, Y = 0, X = 0;
      if (valid) {
        b = p / plane;
        const int r = p - b * plane;
        const int i = r / gp.Wl;
        const int j = r - i * gp.Wl;
        Y = i * gp.out_stride + gp.out_off_y;
        X = j * gp.out_stride + gp.out_off_x;
        valid = i < gp.My && j < gp.Mx && Y >= 0 && Y < gp.out_H && X >= 0 && X < gp.out_W;
      }
      const int64_t pix = (int64_t)Y * gp.out_W + X;
      const int64_t ybase = (int64_t)b * gp.Cout * ostride_c + pix;
      const float nz = (valid && has_noise) ? nw * __ldg(noise + (int64_t)b * oplane + pix) : 0.f;
      const float* om = out_mul ? out_mul + (int64_t)b * gp.Cout : nullptr;
      float rgb[3] = {0.f, 0.f, 0.f};

      mbar_wait(tfull_bar, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_acc + ((uint32_t)(quarter * 32) << 16);
      for (int c0 = half * 32; c0 < n_eff; c0 += 64) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (valid) {
          const int o0 = n0 + c0;
          const bool full = o0 + 32 <= gp.Cout;
          // per-channel factor out_scale * out_mul[b, o]
          float f[32];
          if (om && full && vec_ok) {
            const float4* omp = reinterpret_cast<const float4*>(om + o0);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 t4 = __ldg(omp + k);
              f[4 * k] = t4.x * gp.out_scale;
              f[4 * k + 1] = t4.y * gp.out_scale;
              f[4 * k + 2] = t4.z * gp.out_scale;
              f[4 * k + 3] = t4.w * gp.out_scale;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int oc = o0 + k < gp.Cout ? o0 + k : gp.Cout - 1;
              f[k] = (om ? __ldg(om + oc) : 1.f) * gp.out_scale;
            }
          }
          if (nchw_only) {
            float* yp = y + ybase + (int64_t)o0 * ostride_c;
            if (plain && full) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                *yp = v[k] * f[k];
                yp += ostride_c;
              }
            } else {
              // rare terms: all loads of a 16-column half are issued before the first dependent use
              const float* rp = residual ? residual + ybase + (int64_t)o0 * ostride_c : nullptr;
#pragma unroll
              for (int h0 = 0; h0 < 32; h0 += 16) {
                float bv[16], rv[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  const int o = o0 + h0 + k;
                  const int oc = o < gp.Cout ? o : gp.Cout - 1;
                  bv[k] = bias ? __ldg(bias + oc) : 0.f;
                  rv[k] = (rp && o < gp.Cout) ? __ldg(rp + (int64_t)(h0 + k) * ostride_c) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                  float r = v[h0 + k] * f[h0 + k] + nz + bv[k];
                  if (gp.act) r = (r > 0.f ? r : r * gp.act_alpha) * gp.act_gain;
                  if (o0 + h0 + k < gp.Cout) yp[(int64_t)(h0 + k) * ostride_c] = r + rv[k];
                }
              }
            }
          } else {
            // ---- general sinks (host guarantees Cout % 32 == 0, so every block is full and 16-byte aligned)
            if (bias) {
              const float4* bp = reinterpret_cast<const float4*>(bias + o0);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 t4 = __ldg(bp + k);
                v[4 * k] = v[4 * k] * f[4 * k] + nz + t4.x;  // same association as the NCHW path: bit-identical values
                v[4 * k + 1] = v[4 * k + 1] * f[4 * k + 1] + nz + t4.y;
                v[4 * k + 2] = v[4 * k + 2] * f[4 * k + 2] + nz + t4.z;
                v[4 * k + 3] = v[4 * k + 3] * f[4 * k + 3] + nz + t4.w;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = v[k] * f[k] + nz + 0.f;
            }
            if (gp.act) {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = (v[k] > 0.f ? v[k] : v[k] * gp.act_alpha) * gp.act_gain;
            }
            if (sk.residual_nhwc != nullptr) {
              const float4* rp4 = reinterpret_cast<const float4*>(sk.residual_nhwc + (int64_t)b * gp.res_bstride + pix * gp.Cout + o0);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float4 t4 = __ldg(rp4 + k);
                v[4 * k] += t4.x;
                v[4 * k + 1] += t4.y;
                v[4 * k + 2] += t4.z;
                v[4 * k + 3] += t4.w;
              }
            }
            if (y != nullptr) {
              if (gp.y_nhwc) {
                float4* yp4 = reinterpret_cast<float4*>(y + (int64_t)b * gp.y_bstride + pix * gp.Cout + o0);
#pragma unroll
                for (int k = 0; k < 8; ++k) yp4[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
              } else {
                float* yp = y + ybase + (int64_t)o0 * ostride_c;
#pragma unroll
                for (int k = 0; k < 32; ++k) yp[(int64_t)k * ostride_c] = v[k];
              }
            }
            if (sk.rgb_w != nullptr) {
              // ToRGB (1x1 modulated conv without demodulation, models/spgan_ops.py:1563-1586) folded into the producer
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                if (j < gp.rgb_n) {
                  const float4* wp = reinterpret_cast<const float4*>(sk.rgb_w + ((int64_t)b * gp.rgb_n + j) * gp.Cout + o0);
                  float a = 0.f;
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    const float4 t4 = __ldg(wp + k);
                    a += v[4 * k] * t4.x + v[4 * k + 1] * t4.y + v[4 * k + 2] * t4.z + v[4 * k + 3] * t4.w;
                  }
                  rgb[j] += a;
                }
              }
            }
            if (sk.y_packed != nullptr) {
              // the next conv's A operand: style modulation of THAT conv, then the 16-bit hi/lo split (what spgan_pack_act
              // would compute from the fp32 tensor, bit for bit)
              if (sk.next_mul) {
                const float4* mp = reinterpret_cast<const float4*>(sk.next_mul + (int64_t)b * gp.Cout + o0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float4 t4 = __ldg(mp + k);
                  v[4 * k] *= t4.x;
                  v[4 * k + 1] *= t4.y;
                  v[4 * k + 2] *= t4.z;
                  v[4 * k + 3] *= t4.w;
                }
              }
              const int64_t prow = (int64_t)b * oplane + pix;
              uint16_t* ph = sk.y_packed + prow * gp.pk_cols + o0;
              uint16_t* pl = ph + gp.pk_rows * (int64_t)gp.pk_cols;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  uint16_t h0, l0, h1, l1;
                  if (gp.pk_f16) {
                    split16<true>(v[8 * q + 2 * u], h0, l0);
                    split16<true>(v[8 * q + 2 * u + 1], h1, l1);
                  } else {
                    split16<false>(v[8 * q + 2 * u], h0, l0);
                    split16<false>(v[8 * q + 2 * u + 1], h1, l1);
                  }
                  hw[u] = pack2x16(h0, h1);
                  lw[u] = pack2x16(l0, l1);
                }
                reinterpret_cast<uint4*>(ph)[q] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                reinterpret_cast<uint4*>(pl)[q] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              }
            }
          }
        }
      }
      if (sk.rgb_w != nullptr && valid) {
        // one slot per (N tile, epilogue half): every (slot, b, channel, pixel) is written exactly once, and
        // spgan_rgb_tail sums the slots in a fixed order (deterministic, unlike atomics)
        const int slot = n_tile * 2 + half;
        float* rp = sk.rgb_part + (((int64_t)slot * gp.B + b) * gp.rgb_n) * oplane + pix;
#pragma unroll
        for (int j = 0; j < 3; ++j)
          if (j < gp.rgb_n) rp[(int64_t)j * oplane] = rgb[j];
      }
    }
  }
}

}  // namespace

// Exact-fp32 SIMT implementation of the conv pass (see SpganConvPass in include/spgan_b200.h), its weight gradient,
// the demodulation coefficients and per-plane dot products.
//
// Role on the path: (1) the layers that are NOT GEMM-shaped — ToRGB (Cout = 3, reads a 512-channel activation once:
// HBM-bound) and the 3->3 spherical RGB convs; (2) every gradient pass of this round; (3) the on-device fp32
// comparator for the tcgen05 implicit-GEMM kernel (conv_umma.cu), which takes the dense layers when
// SpganConvPass.precision != 0.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct TapTable {
  int dy[SPGAN_MAX_TAPS], dx[SPGAN_MAX_TAPS], w[SPGAN_MAX_TAPS];
};

__global__ void __launch_bounds__(256) conv_pass_fp32(SpganConvPass p, float* __restrict__ y,
                                                     const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ in_mul, const float* __restrict__ out_mul,
                                                     const float* __restrict__ noise, const float* __restrict__ noise_w,
                                                     const float* __restrict__ bias, const float* __restrict__ residual) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  __shared__ TapTable taps;
  const int tid = threadIdx.x;
  for (int t = tid; t < p.ntaps; t += 256) {
    taps.dy[t] = p.tap_dy[t];
    taps.dx[t] = p.tap_dx[t];
    taps.w[t] = p.tap_w[t];
  }
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int Mtot = p.My * p.Mx;
  const int Ktot = p.Cin * p.ntaps;
  // A loader: pixel lm = tid & 63 (coalesced along x), k rows (tid >> 6) + 4u
  const int lm = tid & 63, lk = tid >> 6;
  const int m_ld = m0 + lm;
  const bool m_ok = m_ld < Mtot;
  const int li = m_ok ? m_ld / p.Mx : 0;
  const int lj = m_ok ? m_ld - li * p.Mx : 0;
  const int iy0 = li * p.in_stride, ix0 = lj * p.in_stride;
  // B loader: k column bk = tid & 15 (contiguous in the native (Cout, Cin, taps) layout), out channel (tid >> 4) + 16u
  const int bk = tid & 15, bn = tid >> 4;
  const float* xb = x + (int64_t)b * p.Cin * p.H * p.W;
  const float* imul = in_mul ? in_mul + (int64_t)b * p.Cin : nullptr;
  // compute mapping: pixels tx + 16u (tx = tid & 15), out channels ty*4 + v (ty = tid >> 4)
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  __syncthreads();

  for (int k0 = 0; k0 < Ktot; k0 += BK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + lk + 4 * u;
      float v = 0.f;
      if (m_ok && kk < Ktot) {
        const int c = kk / p.ntaps, t = kk - c * p.ntaps;
        const int yy = iy0 + taps.dy[t], xx = ix0 + taps.dx[t];
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
          v = __ldg(xb + ((int64_t)c * p.H + yy) * p.W + xx);
          if (imul) v *= __ldg(imul + c);
        }
      }
      As[lk + 4 * u][lm] = v;
    }
    {
      const int kk = k0 + bk;
      const bool k_ok = kk < Ktot;
      const int c = k_ok ? kk / p.ntaps : 0, t = k_ok ? kk - c * p.ntaps : 0;
      const int64_t woff = (int64_t)c * p.ws_c + taps.w[t];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = n0 + bn + 16 * u;
        Bs[bk][bn + 16 * u] = (k_ok && o < p.Cout) ? __ldg(w + (int64_t)o * p.ws_o + woff) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = As[k][tx + 16 * u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bb[v] = Bs[k][ty * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * bb[v];
    }
    __syncthreads();
  }

  const float nw = (noise && noise_w) ? __ldg(noise_w) : 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int m = m0 + tx + 16 * u;
    if (m >= Mtot) continue;
    const int i = m / p.Mx, j = m - i * p.Mx;
    const int Y = i * p.out_stride + p.out_off_y, X = j * p.out_stride + p.out_off_x;
    if (Y < 0 || Y >= p.out_H || X < 0 || X >= p.out_W) continue;
    const float nz = (noise && noise_w) ? nw * __ldg(noise + ((int64_t)b * p.out_H + Y) * p.out_W + X) : 0.f;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int o = n0 + ty * 4 + v;
      if (o >= p.Cout) continue;
      float r = acc[u][v] * p.out_scale;
      if (out_mul) r *= __ldg(out_mul + (int64_t)b * p.Cout + o);
      r += nz;
      if (bias) r += __ldg(bias + o);
      if (p.act) r = (r > 0.f ? r : r * p.act_alpha) * p.act_gain;
      const int64_t cst = p.out_cstride ? p.out_cstride : (int64_t)p.out_H * p.out_W;
      const int64_t idx = ((int64_t)b * p.Cout + o) * cst + (int64_t)Y * p.out_W + X;
      if (residual) r += __ldg(residual + idx);
      y[idx] = r;
    }
  }
}

// Few output channels (ToRGB: 512 -> 3, the 3 -> 3 spherical RGB convs): HBM-bound, not GEMM-shaped.  One thread per
// lattice point accumulates all (<= 4) outputs; the per-sample modulated weights w * in_mul live in shared memory
// as [c][tap][4], so each input element is read exactly once, coalesced along x, and costs one LDS.128 + 4 FMAs.
constexpr int SMALL_COUT = 4;

__global__ void __launch_bounds__(256) conv_small_cout(SpganConvPass p, float* __restrict__ y, const float* __restrict__ x,
                                                      const float* __restrict__ w, const float* __restrict__ in_mul,
                                                      const float* __restrict__ out_mul, const float* __restrict__ noise,
                                                      const float* __restrict__ noise_w, const float* __restrict__ bias,
                                                      const float* __restrict__ residual) {
  extern __shared__ float4 wsm[];  // [Cin][ntaps]
  const int b = blockIdx.y;
  const int nt = p.ntaps;
  for (int i = threadIdx.x; i < p.Cin * nt; i += blockDim.x) {
    const int c = i / nt, t = i - c * nt;
    const float m = in_mul ? __ldg(in_mul + (int64_t)b * p.Cin + c) : 1.f;
    float v[SMALL_COUT];
#pragma unroll
    for (int o = 0; o < SMALL_COUT; ++o)
      v[o] = o < p.Cout ? m * __ldg(w + (int64_t)o * p.ws_o + (int64_t)c * p.ws_c + p.tap_w[t]) : 0.f;
    wsm[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const int Mtot = p.My * p.Mx;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mtot) return;
  const int i = m / p.Mx, j = m - i * p.Mx;
  const int Y = i * p.out_stride + p.out_off_y, X = j * p.out_stride + p.out_off_x;
  if (Y < 0 || Y >= p.out_H || X < 0 || X >= p.out_W) return;
  const float* xb = x + (int64_t)b * p.Cin * p.H * p.W;
  const int64_t plane = (int64_t)p.H * p.W;
  float acc[SMALL_COUT] = {0.f, 0.f, 0.f, 0.f};
  if (nt == 1) {
    const int yy = i * p.in_stride + p.tap_dy[0], xx = j * p.in_stride + p.tap_dx[0];
    if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
      const float* xp = xb + (int64_t)yy * p.W + xx;
      int c = 0;
      for (; c + 8 <= p.Cin; c += 8) {
        float xv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) xv[u] = __ldcs(xp + (int64_t)(c + u) * plane);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 wv = wsm[c + u];
          acc[0] += xv[u] * wv.x;
          acc[1] += xv[u] * wv.y;
          acc[2] += xv[u] * wv.z;
          acc[3] += xv[u] * wv.w;
        }
      }
      for (; c < p.Cin; ++c) {
        const float xv = __ldcs(xp + (int64_t)c * plane);
        const float4 wv = wsm[c];
        acc[0] += xv * wv.x;
        acc[1] += xv * wv.y;
        acc[2] += xv * wv.z;
        acc[3] += xv * wv.w;
      }
    }
  } else {
    for (int c = 0; c < p.Cin; ++c) {
      const float* xc = xb + (int64_t)c * plane;
      for (int t = 0; t < nt; ++t) {
        const int yy = i * p.in_stride + p.tap_dy[t], xx = j * p.in_stride + p.tap_dx[t];
        if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) continue;
        const float xv = __ldg(xc + (int64_t)yy * p.W + xx);
        const float4 wv = wsm[c * nt + t];
        acc[0] += xv * wv.x;
        acc[1] += xv * wv.y;
        acc[2] += xv * wv.z;
        acc[3] += xv * wv.w;
      }
    }
  }
  const int64_t opix = (int64_t)Y * p.out_W + X;
  const int64_t oplane = p.out_cstride ? p.out_cstride : (int64_t)p.out_H * p.out_W;
  const float nz = (noise && noise_w) ? __ldg(noise_w) * __ldg(noise + (int64_t)b * p.out_H * p.out_W + opix) : 0.f;
#pragma unroll
  for (int o = 0; o < SMALL_COUT; ++o) {
    if (o >= p.Cout) break;
    float r = acc[o] * p.out_scale;
    if (out_mul) r *= __ldg(out_mul + (int64_t)b * p.Cout + o);
    r += nz;
    if (bias) r += __ldg(bias + o);
    if (p.act) r = (r > 0.f ? r : r * p.act_alpha) * p.act_gain;
    const int64_t idx = ((int64_t)b * p.Cout + o) * oplane + opix;
    if (residual) r += __ldg(residual + idx);
    y[idx] = r;
  }
}

// dw tile: 64 output channels x 64 (c, tap) columns; reduction over the pixels of sample blockIdx.z; atomicAdd.
__global__ void __launch_bounds__(256) conv_wgrad_fp32(SpganConvPass p, float* __restrict__ dw,
                                                      const float* __restrict__ g, const float* __restrict__ x,
                                                      const float* __restrict__ in_mul,
                                                      const float* __restrict__ out_mul, int nsplit) {
  __shared__ float Gs[BK][BM + 1];  // [pixel][out channel]
  __shared__ float Xs[BK][BN + 1];  // [pixel][(c, tap)]
  __shared__ TapTable taps;
  const int tid = threadIdx.x;
  for (int t = tid; t < p.ntaps; t += 256) {
    taps.dy[t] = p.tap_dy[t];
    taps.dx[t] = p.tap_dx[t];
    taps.w[t] = p.tap_w[t];
  }
  // blockIdx.z = sample * nsplit + pixel chunk: the few-channel layers (ToRGB 512 -> 3, the discriminator's 3 -> 256 stem)
  // have only a handful of (out, in) tiles, so the pixel range is split to fill the SMs (partial sums meet in atomicAdd)
  const int b = blockIdx.z / nsplit;
  const int split = blockIdx.z - b * nsplit;
  const int o0 = blockIdx.x * BM, q0 = blockIdx.y * BN;
  const int Mtot = p.My * p.Mx;
  const int Mchunk = ((Mtot + nsplit - 1) / nsplit + BK - 1) / BK * BK;
  const int Mbeg = split * Mchunk, Mend = min(Mtot, Mbeg + Mchunk);
  const int Qtot = p.Cin * p.ntaps;
  const int lp = tid & 15, lc = tid >> 4;  // loader: pixel lp (contiguous), column lc + 16u
  const int tx = tid & 15, ty = tid >> 4;  // compute: out channels tx + 16u, columns ty*4 + v
  __syncthreads();
  // per-thread loader constants for the 4 columns
  int lo[4], lq_c[4], lq_dy[4], lq_dx[4];
  float gmul[4], xmul[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    lo[u] = o0 + lc + 16 * u;
    gmul[u] = (lo[u] < p.Cout) ? p.out_scale * (out_mul ? __ldg(out_mul + (int64_t)b * p.Cout + lo[u]) : 1.f) : 0.f;
    const int q = q0 + lc + 16 * u;
    if (q < Qtot) {
      const int c = q / p.ntaps, t = q - c * p.ntaps;
      lq_c[u] = c;
      lq_dy[u] = taps.dy[t];
      lq_dx[u] = taps.dx[t];
      xmul[u] = in_mul ? __ldg(in_mul + (int64_t)b * p.Cin + c) : 1.f;
    } else {
      lq_c[u] = -1;
      lq_dy[u] = lq_dx[u] = 0;
      xmul[u] = 0.f;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;

  for (int m0 = Mbeg; m0 < Mend; m0 += BK) {
    const int m = m0 + lp;
    const bool m_ok = m < Mend;
    const int i = m_ok ? m / p.Mx : 0, j = m_ok ? m - i * p.Mx : 0;
    const int Y = i * p.out_stride + p.out_off_y, X = j * p.out_stride + p.out_off_x;
    const bool out_ok = m_ok && Y >= 0 && Y < p.out_H && X >= 0 && X < p.out_W;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float gv = 0.f;
      if (out_ok && lo[u] < p.Cout) gv = gmul[u] * __ldg(g + (((int64_t)b * p.Cout + lo[u]) * p.out_H + Y) * p.out_W + X);
      Gs[lp][lc + 16 * u] = gv;
      float xv = 0.f;
      if (out_ok && lq_c[u] >= 0) {
        const int yy = i * p.in_stride + lq_dy[u], xx = j * p.in_stride + lq_dx[u];
        if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W)
          xv = xmul[u] * __ldg(x + (((int64_t)b * p.Cin + lq_c[u]) * p.H + yy) * p.W + xx);
      }
      Xs[lp][lc + 16 * u] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = Gs[k][tx + 16 * u];
#pragma unroll
      for (int v = 0; v < 4; ++v) bb[v] = Xs[k][ty * 4 + v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * bb[v];
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int o = o0 + tx + 16 * u;
    if (o >= p.Cout) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int q = q0 + ty * 4 + v;
      if (q >= Qtot) continue;
      const int c = q / p.ntaps, t = q - c * p.ntaps;
      atomicAdd(dw + (int64_t)o * p.ws_o + (int64_t)c * p.ws_c + taps.w[t], acc[u][v]);
    }
  }
}

__global__ void __launch_bounds__(256) demod_kernel(float* __restrict__ d, const float* __restrict__ s,
                                                   const float* __restrict__ w, int B, int Cin, int Cout, int taps,
                                                   float scale2, float eps) {
  extern __shared__ float wsq[];
  __shared__ float red[8];
  const int o = blockIdx.x;
  const float* wo = w + (int64_t)o * Cin * taps;
  for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
    float a = 0.f;
    for (int t = 0; t < taps; ++t) {
      const float v = __ldg(wo + (int64_t)c * taps + t);
      a += v * v;
    }
    wsq[c] = a;
  }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    float a = 0.f;
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
      const float sv = __ldg(s + (int64_t)b * Cin + c);
      a += sv * sv * wsq[c];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      d[(int64_t)b * Cout + o] = rsqrtf(t * scale2 + eps);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) plane_dot_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                       const float* __restrict__ b, int64_t planes, int64_t inner) {
  __shared__ float red[8];
  for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
    const float* ap = a + pl * inner;
    const float* bp = b + pl * inner;
    float s = 0.f;
    for (int64_t k = threadIdx.x; k < inner; k += blockDim.x) s += __ldcs(ap + k) * __ldcs(bp + k);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      out[pl] = t;
    }
    __syncthreads();
  }
}

int check_pass(const SpganConvPass* p, const char* who) {
  SPGAN_CHECK_ARG(p != nullptr, "%s: null pass descriptor", who);
  SPGAN_CHECK_ARG(p->B >= 0 && p->Cin >= 0 && p->Cout >= 0 && p->H >= 0 && p->W >= 0 && p->My >= 0 && p->Mx >= 0,
                  "%s: negative size", who);
  SPGAN_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= SPGAN_MAX_TAPS, "%s: %d taps unsupported (1..%d)", who, p->ntaps,
                  SPGAN_MAX_TAPS);
  SPGAN_CHECK_ARG(p->in_stride >= 1 && p->out_stride >= 1, "%s: strides must be >= 1", who);
  SPGAN_CHECK_ARG(p->B <= 65535, "%s: batch %d > 65535", who, p->B);
  return 0;
}

}  // namespace

extern "C" int spgan_conv_pass(const SpganConvPass* p, float* y, const float* x, const float* w, const float* in_mul,
                               const float* out_mul, const float* noise, const float* noise_w, const float* bias,
                               const float* residual, void* stream) {
  if (int e = check_pass(p, "spgan_conv_pass")) return e;
  if (p->B == 0 || p->Cout == 0 || p->My == 0 || p->Mx == 0) return 0;
  SPGAN_CHECK_ARG(y && x && w, "spgan_conv_pass: null pointer");
  SPGAN_CHECK_ARG(p->precision == 0,
                  "spgan_conv_pass: precision %d runs on the tcgen05 path: pack the operands (spgan_pack_act / "
                  "spgan_pack_weight) and call spgan_conv_gemm",
                  p->precision);
  const size_t small_smem = sizeof(float4) * (size_t)p->Cin * p->ntaps;
  if (p->Cout <= SMALL_COUT && small_smem <= 48 * 1024) {
    dim3 grid((p->My * p->Mx + 255) / 256, p->B);
    conv_small_cout<<<grid, 256, small_smem, (cudaStream_t)stream>>>(*p, y, x, w, in_mul, out_mul, noise, noise_w, bias,
                                                                     residual);
    SPGAN_CHECK_LAUNCH("spgan_conv_pass");
    return 0;
  }
  dim3 grid((p->My * p->Mx + BM - 1) / BM, (p->Cout + BN - 1) / BN, p->B);
  conv_pass_fp32<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, y, x, w, in_mul, out_mul, noise, noise_w, bias, residual);
  SPGAN_CHECK_LAUNCH("spgan_conv_pass");
  return 0;
}

extern "C" int spgan_conv_wgrad(const SpganConvPass* p, float* dw, const float* g, const float* x, const float* in_mul,
                                const float* out_mul, int accumulate, void* stream) {
  if (int e = check_pass(p, "spgan_conv_wgrad")) return e;
  SPGAN_CHECK_ARG(dw && g && x, "spgan_conv_wgrad: null pointer");
  if (p->B == 0 || p->Cout == 0 || p->Cin == 0 || p->My == 0 || p->Mx == 0) return 0;
  const int gx = (p->Cout + BM - 1) / BM, gy = (p->Cin * p->ntaps + BN - 1) / BN;
  int nsplit = (2 * SPGAN_NUM_SMS + gx * gy * p->B - 1) / (gx * gy * p->B);
  const int max_split = (p->My * p->Mx + 8 * BK - 1) / (8 * BK);
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  SPGAN_CHECK_ARG((int64_t)p->B * nsplit <= 65535, "spgan_conv_wgrad: batch %d too large", p->B);
  dim3 grid(gx, gy, p->B * nsplit);
  (void)accumulate;  // the caller zeroes dw when it does not accumulate (the tensor extent is only known to the host)
  conv_wgrad_fp32<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, dw, g, x, in_mul, out_mul, nsplit);
  SPGAN_CHECK_LAUNCH("spgan_conv_wgrad");
  return 0;
}

extern "C" int spgan_demod(float* d, const float* s, const float* w, int B, int Cin, int Cout, int taps, float scale,
                           float eps, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && Cin >= 0 && Cout >= 0 && taps >= 1, "spgan_demod: bad size");
  if (B == 0 || Cout == 0) return 0;
  SPGAN_CHECK_ARG(d && s && w, "spgan_demod: null pointer");
  SPGAN_CHECK_ARG(Cin <= 12000, "spgan_demod: Cin %d too large", Cin);
  demod_kernel<<<Cout, 256, sizeof(float) * Cin, (cudaStream_t)stream>>>(d, s, w, B, Cin, Cout, taps, scale * scale, eps);
  SPGAN_CHECK_LAUNCH("spgan_demod");
  return 0;
}

extern "C" int spgan_plane_dot(float* out, const float* a, const float* b, int64_t planes, int64_t inner,
                               void* stream) {
  if (planes <= 0) return 0;
  SPGAN_CHECK_ARG(out && a && b, "spgan_plane_dot: null pointer");
  plane_dot_kernel<<<grid_for(planes, 1, 8, 4), 256, 0, (cudaStream_t)stream>>>(out, a, b, planes, inner);
  SPGAN_CHECK_LAUNCH("spgan_plane_dot");
  return 0;
}

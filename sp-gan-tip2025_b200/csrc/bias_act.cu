// K1: fused bias + activation (forward, gradient, second gradient) and its fused first-order backward.
// HBM-bound: 8 B/element forward, 12 B/element backward.  128-bit accesses whenever a float4 cannot
// straddle a bias channel; otherwise 4 independent scalar accesses in flight per thread.
#include "common.cuh"

namespace {

__device__ __forceinline__ float act_fn(float v, float r, int code, float alpha) {
  switch (code) {
    case 30: return v > 0.f ? v : v * alpha;
    case 31: return r > 0.f ? v : v * alpha;
    case 12:
    case 32: return 0.f;
    default: return v;  // 10, 11 and the reference kernel's `default:` label
  }
}

// 128-bit path for any channel extent >= 4: a float4 may straddle one channel boundary (odd planes such as 101 x 101), so
// each element picks its own bias; the channel of the first element comes from two exact 32-bit fast divisions (the
// hardware-emulated 64-bit division of the scalar path costs more than the memory traffic).  Two independent 128-bit
// loads per thread and iteration are in flight before the first store.
__global__ void __launch_bounds__(256) bias_act_vec4(float4* __restrict__ out, const float4* __restrict__ x,
                                                    const float* __restrict__ bias, const float4* __restrict__ ref,
                                                    uint32_t n4, uint32_t step_b, uint32_t size_b, FastDiv dstep,
                                                    FastDiv dsize, int code, float alpha, float scale) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    float4 v[2], r[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t i = i0 + u * stride;
      ok[u] = i < n4;
      v[u] = ok[u] ? __ldcs(x + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      r[u] = (ok[u] && ref) ? __ldcs(ref + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const uint32_t i = i0 + u * stride;
      float b0 = 0.f, b1 = 0.f;
      uint32_t left = 4;  // elements of this float4 that still belong to the first channel
      if (bias) {
        const uint32_t e = 4u * i;
        const uint32_t c = fdiv(e, dstep);
        left = (c + 1) * step_b - e;
        const uint32_t cm = c - fdiv(c, dsize) * size_b;
        b0 = __ldg(bias + cm);
        b1 = __ldg(bias + (cm + 1 == size_b ? 0 : cm + 1));
      }
      float4 o;
      o.x = act_fn(v[u].x + b0, r[u].x, code, alpha) * scale;
      o.y = act_fn(v[u].y + (left > 1 ? b0 : b1), r[u].y, code, alpha) * scale;
      o.z = act_fn(v[u].z + (left > 2 ? b0 : b1), r[u].z, code, alpha) * scale;
      o.w = act_fn(v[u].w + (left > 3 ? b0 : b1), r[u].w, code, alpha) * scale;
      __stcs(out + i, o);
    }
  }
}

__global__ void __launch_bounds__(256) bias_act_scalar(float* __restrict__ out, const float* __restrict__ x,
                                                      const float* __restrict__ bias, const float* __restrict__ ref,
                                                      int64_t n, int64_t step_b, int64_t size_b, int code, float alpha,
                                                      float scale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    float v[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = __ldcs(x + i + u * stride);
      r[u] = ref ? __ldcs(ref + i + u * stride) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int64_t j = i + u * stride;
      float b = bias ? __ldg(bias + (j / step_b) % size_b) : 0.f;
      __stcs(out + j, act_fn(v[u] + b, r[u], code, alpha) * scale);
    }
  }
  for (; i < n; i += stride) {
    float b = bias ? __ldg(bias + (i / step_b) % size_b) : 0.f;
    float r = ref ? ref[i] : 0.f;
    out[i] = act_fn(x[i] + b, r, code, alpha) * scale;
  }
}

// Backward: blockIdx.x = channel, blockIdx.y = slice of that channel's batch*inner elements.
// Writes grad_in and reduces grad_bias with one atomicAdd per CTA.
__global__ void __launch_bounds__(256) bias_act_bwd_kernel(float* __restrict__ gi, float* __restrict__ gb,
                                                          const float* __restrict__ go, const float* __restrict__ ref,
                                                          int64_t batch, int64_t channels, int64_t inner, FastDiv dinner,
                                                          float alpha, float scale) {
  const int64_t c = blockIdx.x;
  const int64_t total = batch * inner;
  const int64_t per = (total + gridDim.y - 1) / gridDim.y;
  const int64_t begin = (int64_t)blockIdx.y * per;
  const int64_t end = begin + per < total ? begin + per : total;
  float acc = 0.f;
  // batches of 4 independent (grad, ref) load pairs per thread; the sample index of an element comes from one exact
  // 32-bit fast division (total = batch * inner < 2^31 is checked by the launcher)
  for (int64_t e0 = begin + threadIdx.x; e0 < end; e0 += 4 * (int64_t)blockDim.x) {
    float g[4], r[4];
    int64_t idx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t e = e0 + u * (int64_t)blockDim.x;
      idx[u] = -1;
      g[u] = r[u] = 0.f;
      if (e < end) {
        const uint32_t b = fdiv((uint32_t)e, dinner);
        const uint32_t k = (uint32_t)e - b * (uint32_t)inner;
        idx[u] = ((int64_t)b * channels + c) * inner + k;
        g[u] = __ldcs(go + idx[u]);
        r[u] = __ldcs(ref + idx[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (idx[u] < 0) continue;
      const float v = (r[u] > 0.f ? g[u] : g[u] * alpha) * scale;
      __stcs(gi + idx[u], v);
      acc += v;
    }
  }
  __shared__ float warp_sums[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(gb + c, s);
  }
}

// 128-bit variant of the backward over the FLAT tensor: a CTA walks contiguous chunks of 512 float4s (2048 elements, a handful
// of (sample, channel) planes), two 128-bit (grad, ref) load pairs per thread in flight, one 128-bit store each; the bias
// gradient is reduced per plane in shared-memory bins (one shuffle reduction + one shared atomic per warp when the warp sits
// inside one plane, the common case) and flushed with one global atomicAdd per touched channel and chunk.  The per-channel
// kernel above reads 4 bytes per load and decodes an index per element: 0.65 of the HBM roofline.
constexpr int BWD_CHUNK4 = 512;
constexpr int BWD_BINS = 40;  // planes one chunk can touch: 2048 / inner + 2 with inner >= 64

__global__ void __launch_bounds__(256) bias_act_bwd_vec4(float4* __restrict__ gi, float* __restrict__ gb,
                                                        const float4* __restrict__ go, const float4* __restrict__ ref,
                                                        uint32_t n4, uint32_t channels, uint32_t inner, FastDiv dinner,
                                                        FastDiv dchan, float alpha, float scale) {
  __shared__ float bins[BWD_BINS];
  const uint32_t nchunks = (n4 + BWD_CHUNK4 - 1) / BWD_CHUNK4;
  for (uint32_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const uint32_t base4 = chunk * BWD_CHUNK4;
    const uint32_t first_plane = fdiv(4u * base4, dinner);
    if (threadIdx.x < BWD_BINS) bins[threadIdx.x] = 0.f;
    __syncthreads();
    float4 g[2], r[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t i = base4 + threadIdx.x + 256u * u;
      ok[u] = i < n4;
      g[u] = ok[u] ? __ldcs(go + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      r[u] = ok[u] ? __ldcs(ref + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t i = base4 + threadIdx.x + 256u * u;
      const uint32_t e = 4u * i;
      const uint32_t plane = fdiv(e, dinner);
      const uint32_t left = (plane + 1) * inner - e;  // elements of this float4 inside `plane`
      float4 v;
      v.x = (r[u].x > 0.f ? g[u].x : g[u].x * alpha) * scale;
      v.y = (r[u].y > 0.f ? g[u].y : g[u].y * alpha) * scale;
      v.z = (r[u].z > 0.f ? g[u].z : g[u].z * alpha) * scale;
      v.w = (r[u].w > 0.f ? g[u].w : g[u].w * alpha) * scale;
      if (ok[u]) __stcs(gi + i, v);
      float s0 = v.x + (left > 1 ? v.y : 0.f) + (left > 2 ? v.z : 0.f) + (left > 3 ? v.w : 0.f);
      float s1 = (left > 1 ? 0.f : v.y) + (left > 2 ? 0.f : v.z) + (left > 3 ? 0.f : v.w);
      if (!ok[u]) s0 = s1 = 0.f;
      const uint32_t bin = plane - first_plane;
      const uint32_t bin_lo = __shfl_sync(0xffffffffu, bin, 0);
      if (__all_sync(0xffffffffu, bin == bin_lo && left >= 4)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&bins[bin_lo], s0);
      } else if (ok[u]) {
        atomicAdd(&bins[bin], s0);
        if (left < 4) atomicAdd(&bins[bin + 1], s1);
      }
    }
    __syncthreads();
    if (threadIdx.x < BWD_BINS) {
      const float t = bins[threadIdx.x];
      if (t != 0.f) {
        const uint32_t plane = first_plane + threadIdx.x;
        atomicAdd(gb + (plane - fdiv(plane, dchan) * channels), t);
      }
    }
    __syncthreads();
  }
}

// NoiseInjection + bias + leaky-ReLU: plane0 + blockIdx.y = (b, c) plane, threads stride over the plane's pixels.
__global__ void __launch_bounds__(256) noise_bias_act_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                            const float* __restrict__ noise,
                                                            const float* __restrict__ noise_w,
                                                            const float* __restrict__ bias, int64_t plane0,
                                                            int64_t channels, int64_t inner, float alpha, float scale) {
  const int64_t plane = plane0 + blockIdx.y;
  const int64_t b = plane / channels, c = plane - b * channels;
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const float bv = bias ? __ldg(bias + c) : 0.f;
  const float* xp = x + plane * inner;
  const float* np = noise ? noise + b * inner : nullptr;
  float* op = out + plane * inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < inner; i0 += 4 * stride) {
    float v[4], z[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // all loads of the batch before the first store
      const int64_t i = i0 + u * stride;
      v[u] = i < inner ? __ldcs(xp + i) : 0.f;
      z[u] = (np && i < inner) ? __ldg(np + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= inner) continue;
      float t = v[u] + bv;
      if (np) t += nw * z[u];
      __stcs(op + i, (t > 0.f ? t : t * alpha) * scale);
    }
  }
}

// 128-bit variant of the kernel above over the FLAT tensor (as bias_act_vec4): a float4 may straddle one plane boundary
// (odd planes such as 101 x 101), so every element picks its own (sample, channel, pixel); the plane of the first element
// comes from two exact 32-bit fast divisions.  The scalar plane-per-CTA kernel above reads 4 bytes per load and sat at 0.63 of
// the HBM roofline against 0.96 for the float4 bias-act kernel.
__global__ void __launch_bounds__(256) noise_bias_act_vec4(float4* __restrict__ out, const float4* __restrict__ x,
                                                          const float* __restrict__ noise, const float* __restrict__ noise_w,
                                                          const float* __restrict__ bias, uint32_t n4, uint32_t channels,
                                                          uint32_t inner, FastDiv dinner, FastDiv dchan, float alpha,
                                                          float scale) {
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
    float4 v[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t i = i0 + u * stride;
      ok[u] = i < n4;
      v[u] = ok[u] ? __ldcs(x + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const uint32_t i = i0 + u * stride;
      const uint32_t e = 4u * i;
      const uint32_t plane = fdiv(e, dinner);
      const uint32_t p = e - plane * inner;
      const uint32_t b = fdiv(plane, dchan);
      const uint32_t c = plane - b * channels;
      const uint32_t left = inner - p;  // elements of this float4 that still belong to the first plane
      // the next plane: channel c + 1 of the same sample, or channel 0 of the next sample
      const uint32_t c1 = c + 1 == channels ? 0u : c + 1;
      const uint32_t b1 = c + 1 == channels ? b + 1 : b;
      const float bv0 = bias ? __ldg(bias + c) : 0.f, bv1 = bias ? __ldg(bias + c1) : 0.f;
      float in[4] = {v[u].x, v[u].y, v[u].z, v[u].w}, o[4];
#pragma unroll
      for (uint32_t j = 0; j < 4; ++j) {
        const bool first = j < left;
        float t = in[j] + (first ? bv0 : bv1);
        if (noise) t += nw * __ldg(noise + (first ? (uint64_t)b * inner + p + j : (uint64_t)b1 * inner + (j - left)));
        o[j] = (t > 0.f ? t : t * alpha) * scale;
      }
      __stcs(out + i, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

}  // namespace

extern "C" int spgan_noise_bias_act(float* out, const float* x, const float* noise, const float* noise_w,
                                    const float* bias, int64_t batch, int64_t channels, int64_t inner, float alpha,
                                    float scale, void* stream) {
  SPGAN_CHECK_ARG(batch >= 0 && channels >= 0 && inner >= 0, "spgan_noise_bias_act: negative size");
  if (batch * channels * inner == 0) return 0;
  SPGAN_CHECK_ARG(out && x, "spgan_noise_bias_act: null pointer");
  SPGAN_CHECK_ARG((noise == nullptr) == (noise_w == nullptr), "spgan_noise_bias_act: noise and noise_w go together");
  const int64_t n = batch * channels * inner;
  if (n % 4 == 0 && n < (1LL << 31) && inner >= 4 && channels < (1LL << 31) && ((((uintptr_t)out) | ((uintptr_t)x)) & 15) == 0) {
    const int64_t n4 = n / 4;
    noise_bias_act_vec4<<<grid_for(n4, 512, 8), 256, 0, (cudaStream_t)stream>>>(
        (float4*)out, (const float4*)x, noise, noise_w, bias, (uint32_t)n4, (uint32_t)channels, (uint32_t)inner,
        make_fastdiv((uint32_t)inner), make_fastdiv((uint32_t)channels), alpha, scale);
    SPGAN_CHECK_LAUNCH("spgan_noise_bias_act");
    return 0;
  }
  int64_t gx = ceil_div64(inner, 1024);
  if (gx > 64) gx = 64;
  const int64_t planes = batch * channels;
  for (int64_t p0 = 0; p0 < planes; p0 += 65535) {  // gridDim.y <= 65535
    const int64_t np = planes - p0 < 65535 ? planes - p0 : 65535;
    dim3 grid((unsigned)gx, (unsigned)np);
    noise_bias_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, noise, noise_w, bias, p0, channels, inner,
                                                                  alpha, scale);
  }
  SPGAN_CHECK_LAUNCH("spgan_noise_bias_act");
  return 0;
}

extern "C" int spgan_bias_act(float* out, const float* x, const float* bias, const float* ref, int64_t n,
                              int64_t step_b, int64_t size_b, int act, int grad, float alpha, float scale,
                              void* stream) {
  SPGAN_CHECK_ARG(n >= 0, "spgan_bias_act: negative size");
  if (n == 0) return 0;
  SPGAN_CHECK_ARG(out && x, "spgan_bias_act: input must be a CUDA tensor (null pointer)");
  SPGAN_CHECK_ARG(!bias || (step_b > 0 && size_b > 0), "spgan_bias_act: bad bias geometry");
  const int code = act * 10 + grad;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned = ((((uintptr_t)out) | ((uintptr_t)x) | ((uintptr_t)ref)) & 15) == 0;
  if (aligned && n % 4 == 0 && n < (1LL << 31) && (!bias || (step_b >= 4 && step_b < (1LL << 31) && size_b < (1LL << 31)))) {
    const int64_t n4 = n / 4;
    const uint32_t sb = bias ? (uint32_t)step_b : 4u, zb = bias ? (uint32_t)size_b : 1u;
    bias_act_vec4<<<grid_for(n4, 512, 8), 256, 0, st>>>((float4*)out, (const float4*)x, bias, (const float4*)ref, (uint32_t)n4,
                                                        sb, zb, make_fastdiv(sb), make_fastdiv(zb), code, alpha, scale);
  } else {
    bias_act_scalar<<<grid_for(n, 1024, 8), 256, 0, st>>>(out, x, bias, ref, n, bias ? step_b : 1, bias ? size_b : 1,
                                                          code, alpha, scale);
  }
  SPGAN_CHECK_LAUNCH("spgan_bias_act");
  return 0;
}

extern "C" int spgan_bias_act_bwd(float* grad_in, float* grad_bias, const float* grad_out, const float* out_ref,
                                  int64_t batch, int64_t channels, int64_t inner, float alpha, float scale,
                                  void* stream) {
  SPGAN_CHECK_ARG(batch >= 0 && channels >= 0 && inner >= 0, "spgan_bias_act_bwd: negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (channels > 0) SPGAN_CUDA(cudaMemsetAsync(grad_bias, 0, sizeof(float) * channels, st), "spgan_bias_act_bwd memset");
  if (batch * channels * inner == 0) return 0;
  SPGAN_CHECK_ARG(grad_in && grad_bias && grad_out && out_ref, "spgan_bias_act_bwd: null pointer");
  SPGAN_CHECK_ARG(channels <= 2147483647, "spgan_bias_act_bwd: too many channels");
  SPGAN_CHECK_ARG(batch * inner < (1LL << 31), "spgan_bias_act_bwd: batch * inner = %lld exceeds 2^31", (long long)(batch * inner));
  const int64_t n = batch * channels * inner;
  if (n % 4 == 0 && n < (1LL << 31) && inner >= 64 &&
      ((((uintptr_t)grad_in) | ((uintptr_t)grad_out) | ((uintptr_t)out_ref)) & 15) == 0) {
    const int64_t n4 = n / 4;
    bias_act_bwd_vec4<<<grid_for(ceil_div64(n4, BWD_CHUNK4), 1, 8, 8), 256, 0, st>>>(
        (float4*)grad_in, grad_bias, (const float4*)grad_out, (const float4*)out_ref, (uint32_t)n4, (uint32_t)channels,
        (uint32_t)inner, make_fastdiv((uint32_t)inner), make_fastdiv((uint32_t)channels), alpha, scale);
    SPGAN_CHECK_LAUNCH("spgan_bias_act_bwd");
    return 0;
  }
  // enough slices that channels*slices covers >= 4 waves of 148 SMs x 8 CTAs, each slice >= 2048 elements
  int64_t slices = ceil_div64((int64_t)SPGAN_NUM_SMS * 8 * 4, channels);
  const int64_t max_slices = ceil_div64(batch * inner, 2048);
  if (slices > max_slices) slices = max_slices;
  if (slices > 65535) slices = 65535;
  if (slices < 1) slices = 1;
  dim3 grid((unsigned)channels, (unsigned)slices);
  bias_act_bwd_kernel<<<grid, 256, 0, st>>>(grad_in, grad_bias, grad_out, out_ref, batch, channels, inner,
                                            make_fastdiv((uint32_t)inner), alpha, scale);
  SPGAN_CHECK_LAUNCH("spgan_bias_act_bwd");
  return 0;
}

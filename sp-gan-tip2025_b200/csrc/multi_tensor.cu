// Multi-tensor / tail kernels of the training loop.
//   ema_multi         : EMA `accumulate` (utils.py:86-94: p_ema = p_ema * decay + p * (1 - decay) for every parameter) as ONE
//                       launch over a device table of (dst, src, count) chunks instead of 2 x 115 tiny launches.
//   minibatch_stddev  : the discriminator's minibatch-stddev feature (models/stylegan2discriminator.py:205-212): per group of
//                       `group` samples the standard deviation over the group at every (channel, pixel), averaged to one scalar
//                       per sub-batch, appended as channel C of the output — fused with the concatenation copy.
// Roofline: HBM (12 B per parameter for the EMA; read + write of the feature map for the stddev).
#include "common.cuh"

namespace {

struct EmaChunk {
  float* dst;
  const float* src;
  int64_t n;
};

constexpr int EMA_CHUNK = 16384;  // elements per CTA

__global__ void __launch_bounds__(256) ema_multi_kernel(const EmaChunk* __restrict__ table, float decay, float alpha) {
  const EmaChunk c = table[blockIdx.x];
  // same arithmetic as torch's mul_(decay) followed by add_(src, alpha = 1 - decay): fma(alpha, src, dst * decay)
  if (((((uintptr_t)c.dst) | ((uintptr_t)c.src)) & 15) == 0) {
    const int n4 = (int)(c.n >> 2);
    float4* d4 = reinterpret_cast<float4*>(c.dst);
    const float4* s4 = reinterpret_cast<const float4*>(c.src);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 d = d4[i];
      const float4 s = __ldg(s4 + i);
      d.x = __fmaf_rn(alpha, s.x, __fmul_rn(d.x, decay));
      d.y = __fmaf_rn(alpha, s.y, __fmul_rn(d.y, decay));
      d.z = __fmaf_rn(alpha, s.z, __fmul_rn(d.z, decay));
      d.w = __fmaf_rn(alpha, s.w, __fmul_rn(d.w, decay));
      d4[i] = d;
    }
    for (int64_t i = ((int64_t)n4 << 2) + threadIdx.x; i < c.n; i += blockDim.x)
      c.dst[i] = __fmaf_rn(alpha, __ldg(c.src + i), __fmul_rn(c.dst[i], decay));
  } else {
    for (int64_t i = threadIdx.x; i < c.n; i += blockDim.x)
      c.dst[i] = __fmaf_rn(alpha, __ldg(c.src + i), __fmul_rn(c.dst[i], decay));
  }
}

// Phase 1: grid (M sub-batches, nblk): per-position standard deviation over the group, block-reduced to partial[m][blk].
__global__ void __launch_bounds__(256) stddev_partial_kernel(float* __restrict__ partial, const float* __restrict__ h, int M,
                                                            int group, int64_t chw, float eps) {
  __shared__ float red[8];
  const int m = blockIdx.x;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < chw; i += (int64_t)gridDim.y * blockDim.x) {
    // sample n of sub-batch m is batch row n * M + m (the reference's view(group, -1, ...) puts the group index first)
    float mean = 0.f;
    for (int n = 0; n < group; ++n) mean += __ldg(h + ((int64_t)n * M + m) * chw + i);
    mean /= (float)group;
    float var = 0.f;
    for (int n = 0; n < group; ++n) {
      const float d = __ldg(h + ((int64_t)n * M + m) * chw + i) - mean;
      var += d * d;
    }
    acc += sqrtf(var / (float)group + eps);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[m * gridDim.y + blockIdx.y] = s;
  }
}

// Phase 2: out (B, C + 1, HW) = [h, stddev of the sample's sub-batch broadcast over the pixels].
__global__ void __launch_bounds__(256) stddev_concat_kernel(float* __restrict__ out, const float* __restrict__ h,
                                                           const float* __restrict__ partial, int B, int M, int nblk, int C,
                                                           int HW) {
  const int64_t chw = (int64_t)C * HW;
  const int64_t per = chw + HW;
  const int64_t total = (int64_t)B * per;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const int64_t r = idx - (int64_t)b * per;
    if (r < chw) {
      out[idx] = __ldg(h + (int64_t)b * chw + r);
    } else {
      const int m = b % M;
      float s = 0.f;
      for (int k = 0; k < nblk; ++k) s += __ldg(partial + m * nblk + k);  // fixed order: deterministic
      out[idx] = s / (float)chw;
    }
  }
}

}  // namespace

extern "C" int spgan_ema_chunk_elems(void) { return EMA_CHUNK; }

extern "C" int spgan_ema_multi(const void* table, int nchunks, float decay, float alpha, void* stream) {
  SPGAN_CHECK_ARG(nchunks >= 0, "spgan_ema_multi: negative chunk count");
  if (nchunks == 0) return 0;
  SPGAN_CHECK_ARG(table != nullptr && (((uintptr_t)table) & 7) == 0, "spgan_ema_multi: chunk table must be an 8-byte aligned device pointer");
  ema_multi_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const EmaChunk*>(table), decay, alpha);
  SPGAN_CHECK_LAUNCH("spgan_ema_multi");
  return 0;
}

extern "C" int spgan_minibatch_stddev(float* out, float* partial, const float* h, int B, int C, int HW, int group, float eps,
                                      void* stream) {
  SPGAN_CHECK_ARG(B >= 1 && C >= 1 && HW >= 1 && group >= 1 && B % group == 0, "spgan_minibatch_stddev: batch %d must be a multiple of the group %d", B, group);
  SPGAN_CHECK_ARG(out && partial && h, "spgan_minibatch_stddev: null pointer");
  const int M = B / group;
  const int64_t chw = (int64_t)C * HW;
  const int nblk = 8;  // partial (M, 8)
  stddev_partial_kernel<<<dim3(M, nblk), 256, 0, (cudaStream_t)stream>>>(partial, h, M, group, chw, eps);
  SPGAN_CHECK_LAUNCH("spgan_minibatch_stddev");
  stddev_concat_kernel<<<grid_for((int64_t)B * (chw + HW), 256, 4), 256, 0, (cudaStream_t)stream>>>(out, h, partial, B, M, nblk, C, HW);
  SPGAN_CHECK_LAUNCH("spgan_minibatch_stddev");
  return 0;
}

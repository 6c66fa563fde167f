// L7: EqualLinear — y = x (W * w_scale)^T + bias * b_scale, optional fused leaky-relu * gain.
// The layers on this path are tiny (M = batch <= 64 rows, K, N <= 4608): latency-bound, not GEMM-peak-bound, so one
// warp owns one output column, keeps its weight row in registers and streams the activations (L1/L2 resident).
#include "common.cuh"

namespace {

constexpr int LIN_MAX_KREG = 16;  // weight-row elements per lane kept in registers per K chunk (512 columns)

__global__ void __launch_bounds__(128) linear_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                    const float* __restrict__ w, const float* __restrict__ bias, int M,
                                                    int N, int K, float w_scale, float b_scale, int act, float alpha,
                                                    float gain) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int m0 = blockIdx.y * 32;
  const int m1 = min(m0 + 32, M);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const float* wr = w + (int64_t)n * K;
  for (int k0 = 0; k0 < K; k0 += 32 * LIN_MAX_KREG) {
    float wreg[LIN_MAX_KREG];
#pragma unroll
    for (int j = 0; j < LIN_MAX_KREG; ++j) {
      const int k = k0 + j * 32 + lane;
      wreg[j] = k < K ? __ldg(wr + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int m = m0 + i;
      if (m < m1) {
        const float* xr = x + (int64_t)m * K;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < LIN_MAX_KREG; ++j) {
          const int k = k0 + j * 32 + lane;
          if (k < K) s += wreg[j] * __ldg(xr + k);
        }
        acc[i] += s;
      }
    }
  }
  const float b = bias ? bias[n] * b_scale : 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float s = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int m = m0 + i;
    if (lane == 0 && m < m1) {
      float v = s * w_scale + b;
      if (act) v = (v > 0.f ? v : v * alpha) * gain;
      y[(int64_t)m * N + n] = v;
    }
  }
}

// Small-batch variant (M <= 64, K a multiple of 128, 16-byte aligned operands): the kernel above gives one warp per output
// column and 4 warps per CTA — N / 4 CTAs of 128 threads, ~3.5 warps per SM for the 512-wide layers of the mapping network and
// the modulations — so a launch is one long dependent chain of L1 loads (~19 us for 2 MFLOP; the training iteration issues
// 560 of them = 9 % of its kernel time).  Here a CTA of 8 warps owns 4 output columns and a tile of 16 rows: the two warps of
// a column split K, the activation tile is staged in shared memory per 512-wide K chunk (128-bit reads, no global latency in
// the inner loop), weights come in as 128-bit coalesced loads.  7 warps per SM at N = 512, M = 8.
constexpr int LS_ROWS = 16;
constexpr int LS_CHUNK = 512;

__global__ void __launch_bounds__(256) linear_small_kernel(float* __restrict__ y, const float* __restrict__ x,
                                                          const float* __restrict__ w, const float* __restrict__ bias, int M,
                                                          int N, int K, float w_scale, float b_scale, int act, float alpha,
                                                          float gain) {
  __shared__ __align__(16) float xs[LS_ROWS][LS_CHUNK];
  __shared__ float part[4][LS_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = warp & 3, khalf = warp >> 2;
  const int n = blockIdx.x * 4 + col;
  const int m0 = blockIdx.y * LS_ROWS;
  const int rows = min(LS_ROWS, M - m0);
  float acc[LS_ROWS];
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r) acc[r] = 0.f;
  const bool live = n < N;
  for (int k0 = 0; k0 < K; k0 += LS_CHUNK) {
    const int kc = min(LS_CHUNK, K - k0);  // multiple of 128
    __syncthreads();
    for (int idx = threadIdx.x; idx < LS_ROWS * (kc >> 2); idx += blockDim.x) {
      const int r = idx / (kc >> 2), q = idx - r * (kc >> 2);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = __ldg(reinterpret_cast<const float4*>(x + (int64_t)(m0 + r) * K + k0) + q);
      *reinterpret_cast<float4*>(&xs[r][4 * q]) = v;
    }
    __syncthreads();
    if (live) {
      // this warp's half of the chunk: float4 index lane + 32 j inside [khalf * kc/2, (khalf + 1) * kc/2)
      const int half4 = kc >> 3;  // float4s per half
      for (int j = lane; j < half4; j += 32) {
        const int q = khalf * half4 + j;
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)n * K + k0) + q);
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r) {
          const float4 xv = *reinterpret_cast<const float4*>(&xs[r][4 * q]);
          acc[r] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
  if (khalf == 1 && lane < LS_ROWS) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r)
      if (r == lane) v = acc[r];
    part[col][lane] = v;
  }
  __syncthreads();
  if (khalf == 0 && lane < rows && live) {
    float v = 0.f;
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r)
      if (r == lane) v = acc[r];
    v = (v + part[col][lane]) * w_scale + (bias ? __ldg(bias + n) * b_scale : 0.f);
    if (act) v = (v > 0.f ? v : v * alpha) * gain;
    y[(int64_t)(m0 + lane) * N + n] = v;
  }
}

// Weight gradient of the linear layer: dw[n][k] = scale * sum_m g[m][n] * x[m][k] with M = batch (<= 64): an outer-product
// accumulation, HBM-bound on writing dw (the general kernel above would run a K = batch contraction with 8 of 32 lanes).
__global__ void __launch_bounds__(256) linear_wgrad_kernel(float* __restrict__ dw, const float* __restrict__ g,
                                                          const float* __restrict__ x, int M, int N, int K, float scale) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = blockIdx.y * 8;
  if (k >= K) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int m = 0; m < M; ++m) {
    const float xv = __ldg(x + (int64_t)m * K + k);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + i;
      acc[i] += (n < N ? __ldg(g + (int64_t)m * N + n) : 0.f) * xv;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (n0 + i < N) dw[(int64_t)(n0 + i) * K + k] = acc[i] * scale;
}

}  // namespace

extern "C" int spgan_linear_wgrad(float* dw, const float* g, const float* x, int M, int N, int K, float scale,
                                  void* stream) {
  SPGAN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "spgan_linear_wgrad: negative size");
  if (N == 0 || K == 0) return 0;
  SPGAN_CHECK_ARG(dw && (M == 0 || (g && x)), "spgan_linear_wgrad: null pointer");
  SPGAN_CHECK_ARG((N + 7) / 8 <= 65535, "spgan_linear_wgrad: N=%d too large", N);
  dim3 grid((K + 255) / 256, (N + 7) / 8);
  linear_wgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dw, g, x, M, N, K, scale);
  SPGAN_CHECK_LAUNCH("spgan_linear_wgrad");
  return 0;
}

extern "C" int spgan_linear(float* y, const float* x, const float* w, const float* bias, int M, int N, int K,
                            float w_scale, float b_scale, int act, float alpha, float gain, void* stream) {
  SPGAN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "spgan_linear: negative size");
  if (M == 0 || N == 0) return 0;
  SPGAN_CHECK_ARG(y && x && w, "spgan_linear: null pointer");
  if (M <= 64 && K % 128 == 0 && ((((uintptr_t)x) | ((uintptr_t)w)) & 15) == 0) {
    dim3 grid((N + 3) / 4, (M + LS_ROWS - 1) / LS_ROWS);
    linear_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(y, x, w, bias, M, N, K, w_scale, b_scale, act, alpha, gain);
    SPGAN_CHECK_LAUNCH("spgan_linear");
    return 0;
  }
  dim3 grid((N + 3) / 4, (M + 31) / 32);
  linear_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(y, x, w, bias, M, N, K, w_scale, b_scale, act, alpha, gain);
  SPGAN_CHECK_LAUNCH("spgan_linear");
  return 0;
}

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
  dim3 grid((N + 3) / 4, (M + 31) / 32);
  linear_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(y, x, w, bias, M, N, K, w_scale, b_scale, act, alpha, gain);
  SPGAN_CHECK_LAUNCH("spgan_linear");
  return 0;
}

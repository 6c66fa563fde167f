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

}  // namespace

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

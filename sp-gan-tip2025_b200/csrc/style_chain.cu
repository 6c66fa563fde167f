// Mapping network and per-layer modulation / demodulation of the generator as two launches.
//
// The reference runs, per generator call, PixelNorm + 8 x (EqualLinear 512 -> 512 + fused leaky-ReLU) (models/spgan/spgan.py:
// 405-412, models/ops.py:190-222) and, for each of the 20 modulated convs, one EqualLinear 512 -> Cin plus the demodulation
// rsqrt(sum (scale W s)^2 + 1e-8) (models/ops.py:598-604): ~50 tiny launches that depend only on the latent.
//   mapping_chain    : ONE kernel.  A thread-block cluster of 8 CTAs owns a tile of 8 latent rows; CTA r computes output features
//                      [64 r, 64 r + 64) of every layer and writes them into the activation buffer of ALL 8 CTAs through
//                      distributed shared memory (st.shared::cluster), one cluster barrier per layer; the 1 MB weight of a
//                      layer is read once per cluster, 128 KB per CTA, coalesced.
//   modulation_batch : ONE kernel over a device table of layer descriptors: CTA (layer, 8-row tile) computes the modulation
//                      s = style (W_m scale)^T + b lr_mul into shared memory and, from it, the demodulation
//                      d = rsqrt(scale_c^2 sum_c s^2 wsq[o][c] + eps) with wsq[o][c] = sum_t W[o][c][t]^2 (cached per weight
//                      version by the host).
// Roofline: latency / L2 (8 MB of mapping weights, ~30 MB of modulation tables per call); they replace launch overhead, not
// bandwidth.
#include "common.cuh"

namespace {

constexpr int SC_ROWS = 8;     // latent rows per cluster / CTA
constexpr int SC_DIM = 512;    // latent width
constexpr int SC_CLUSTER = 8;  // CTAs per cluster: 64 output features each
constexpr int SC_MAX_LAYERS = 16;

struct MappingParams {
  const float* w[SC_MAX_LAYERS];  // (512, 512) each
  const float* b[SC_MAX_LAYERS];  // (512) each or null
  int n_layers;
  float w_scale, b_scale, alpha, gain, eps;
};

__device__ __forceinline__ uint32_t sc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __cluster_dims__(SC_CLUSTER, 1, 1) __launch_bounds__(256)
mapping_chain_kernel(float* __restrict__ out, const float* __restrict__ z, int64_t z_stride, int B, const MappingParams P) {
  __shared__ __align__(16) float act[2][SC_ROWS][SC_DIM];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int row0 = (blockIdx.x / SC_CLUSTER) * SC_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- PixelNorm (models/ops.py:13-20): every CTA of the cluster normalises the whole row tile into its own buffer
  {
    const int r = warp;  // 8 warps = 8 rows
    float v[SC_DIM / 32];
    float ss = 0.f;
    const bool live = row0 + r < B;
#pragma unroll
    for (int i = 0; i < SC_DIM / 32; ++i) {
      v[i] = live ? __ldg(z + (int64_t)(row0 + r) * z_stride + lane + 32 * i) : 0.f;
      ss += v[i] * v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(ss / (float)SC_DIM + P.eps);
#pragma unroll
    for (int i = 0; i < SC_DIM / 32; ++i) act[0][r][lane + 32 * i] = v[i] * inv;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  int cur = 0;
  for (int l = 0; l < P.n_layers; ++l) {
    const float* __restrict__ W = P.w[l];
    const float* __restrict__ bias = P.b[l];
    // warp `warp` computes output features o = 64 rank + 8 warp + j, j < 8, for the 8 rows of the tile
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const int o = (int)rank * 64 + warp * 8 + j;
      float4 wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)o * SC_DIM + 128 * i) + lane);
      float acc[SC_ROWS];
#pragma unroll
      for (int r = 0; r < SC_ROWS; ++r) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(&act[cur][r][128 * i + 4 * lane]);
          s += wv[i].x * x.x + wv[i].y * x.y + wv[i].z * x.z + wv[i].w * x.w;
        }
        acc[r] = s;
      }
#pragma unroll
      for (int r = 0; r < SC_ROWS; ++r)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
      // lane (8 t + r) delivers row r to CTA t of the cluster: every CTA receives the whole next activation
      const int r = lane & 7, target = lane >> 3;
      float y = 0.f;
#pragma unroll
      for (int q = 0; q < SC_ROWS; ++q)
        if (q == r) y = acc[q];
      y = y * P.w_scale + (bias ? __ldg(bias + o) * P.b_scale : 0.f);
      y = (y > 0.f ? y : y * P.alpha) * P.gain;
      const uint32_t local = sc_smem_u32(&act[cur ^ 1][r][o]);
#pragma unroll
      for (int h = 0; h < SC_CLUSTER / 4; ++h) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((uint32_t)(target + 4 * h)));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(y) : "memory");
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    cur ^= 1;
  }
  // ---- every CTA holds the final activations: CTA r writes its 64-feature slice
  for (int idx = threadIdx.x; idx < SC_ROWS * 64; idx += blockDim.x) {
    const int r = idx / 64, c = (int)rank * 64 + idx % 64;
    if (row0 + r < B) out[(int64_t)(row0 + r) * SC_DIM + c] = act[cur][r][c];
  }
}

// One record per modulated conv.  All pointers are device pointers; `style_sel` picks the latent source (0: styles
// (B, n_latent, 512) column style_idx; 1: the raw global latent (B, 2, 512) column 0, structure synthesiser).
struct ModLayer {
  const float* wm;    // (Cin, 512) modulation weight
  const float* bm;    // (Cin) modulation bias or null
  const float* wsq;   // (Cout, Cin) sum over taps of W^2, or null (no demodulation)
  int64_t s_off;      // offset of s (B, Cin) in the output buffer
  int64_t d_off;      // offset of d (B, Cout) in the output buffer
  int32_t Cin, Cout;
  int32_t style_sel, style_idx;
  float m_scale, m_lr_mul, c_scale, eps;
};

__global__ void __launch_bounds__(256) modulation_batch_kernel(float* __restrict__ out, const ModLayer* __restrict__ layers,
                                                              const float* __restrict__ styles, int64_t styles_bstride,
                                                              const float* __restrict__ gl, int64_t gl_bstride, int B) {
  __shared__ __align__(16) float sty[SC_ROWS][SC_DIM];
  __shared__ float s_sm[SC_ROWS][SC_DIM + 8];  // Cin <= 520
  const ModLayer L = layers[blockIdx.x];
  const int row0 = blockIdx.y * SC_ROWS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int idx = threadIdx.x; idx < SC_ROWS * SC_DIM; idx += blockDim.x) {
    const int r = idx / SC_DIM, k = idx - r * SC_DIM;
    const int b = row0 + r;
    float v = 0.f;
    if (b < B) v = L.style_sel ? __ldg(gl + (int64_t)b * gl_bstride + k) : __ldg(styles + (int64_t)b * styles_bstride + (int64_t)L.style_idx * SC_DIM + k);
    sty[r][k] = v;
  }
  __syncthreads();
  // ---- modulation: s[r][c] = sum_k style[r][k] wm[c][k] * m_scale + bm[c] * m_lr_mul
  for (int c = warp; c < L.Cin; c += 8) {
    float4 wv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(L.wm + (int64_t)c * SC_DIM + 128 * i) + lane);
    float acc[SC_ROWS];
#pragma unroll
    for (int r = 0; r < SC_ROWS; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 x = *reinterpret_cast<const float4*>(&sty[r][128 * i + 4 * lane]);
        s += wv[i].x * x.x + wv[i].y * x.y + wv[i].z * x.z + wv[i].w * x.w;
      }
      acc[r] = s;
    }
#pragma unroll
    for (int r = 0; r < SC_ROWS; ++r)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
    if (lane < SC_ROWS) {
      float y = 0.f;
#pragma unroll
      for (int q = 0; q < SC_ROWS; ++q)
        if (q == lane) y = acc[q];
      y = y * L.m_scale + (L.bm ? __ldg(L.bm + c) * L.m_lr_mul : 0.f);
      s_sm[lane][c] = y;
      if (row0 + lane < B) out[L.s_off + (int64_t)(row0 + lane) * L.Cin + c] = y;
    }
  }
  if (L.wsq == nullptr) return;
  __syncthreads();
  // ---- demodulation: d[r][o] = rsqrt(c_scale^2 * sum_c s[r][c]^2 wsq[o][c] + eps)
  for (int o = warp; o < L.Cout; o += 8) {
    float acc[SC_ROWS];
#pragma unroll
    for (int r = 0; r < SC_ROWS; ++r) acc[r] = 0.f;
    for (int c = lane; c < L.Cin; c += 32) {
      const float q = __ldg(L.wsq + (int64_t)o * L.Cin + c);
#pragma unroll
      for (int r = 0; r < SC_ROWS; ++r) acc[r] += s_sm[r][c] * s_sm[r][c] * q;
    }
#pragma unroll
    for (int r = 0; r < SC_ROWS; ++r)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], off);
    if (lane < SC_ROWS && row0 + lane < B) {
      float y = 0.f;
#pragma unroll
      for (int q = 0; q < SC_ROWS; ++q)
        if (q == lane) y = acc[q];
      out[L.d_off + (int64_t)(row0 + lane) * L.Cout + o] = rsqrtf(L.c_scale * L.c_scale * y + L.eps);
    }
  }
}

}  // namespace

extern "C" int spgan_mapping_chain(float* out, const float* z, int64_t z_stride, int B, const float* const* weights,
                                   const float* const* biases, int n_layers, float w_scale, float b_scale, float alpha,
                                   float gain, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && n_layers >= 0 && n_layers <= SC_MAX_LAYERS, "spgan_mapping_chain: bad sizes (B=%d, %d layers)", B, n_layers);
  if (B == 0) return 0;
  SPGAN_CHECK_ARG(out && z && weights && biases && z_stride >= SC_DIM, "spgan_mapping_chain: null pointer / row stride < 512");
  MappingParams P;
  for (int l = 0; l < SC_MAX_LAYERS; ++l) {
    P.w[l] = l < n_layers ? weights[l] : nullptr;
    P.b[l] = l < n_layers ? biases[l] : nullptr;
    SPGAN_CHECK_ARG(l >= n_layers || (P.w[l] != nullptr && (((uintptr_t)P.w[l]) & 15) == 0), "spgan_mapping_chain: weight %d null or misaligned", l);
  }
  P.n_layers = n_layers;
  P.w_scale = w_scale;
  P.b_scale = b_scale;
  P.alpha = alpha;
  P.gain = gain;
  P.eps = 1e-8f;
  const int tiles = (B + SC_ROWS - 1) / SC_ROWS;
  mapping_chain_kernel<<<tiles * SC_CLUSTER, 256, 0, (cudaStream_t)stream>>>(out, z, z_stride, B, P);
  SPGAN_CHECK_LAUNCH("spgan_mapping_chain");
  return 0;
}

extern "C" int spgan_modulation_layer_bytes(void) { return (int)sizeof(ModLayer); }

extern "C" int spgan_modulation_batch(float* out, const void* layers, int n_layers, const float* styles, int64_t styles_bstride,
                                      const float* gl, int64_t gl_bstride, int B, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && n_layers >= 0, "spgan_modulation_batch: negative size");
  if (B == 0 || n_layers == 0) return 0;
  SPGAN_CHECK_ARG(out && layers && (styles || gl), "spgan_modulation_batch: null pointer");
  SPGAN_CHECK_ARG((((uintptr_t)layers) & 7) == 0, "spgan_modulation_batch: descriptor table must be 8-byte aligned");
  dim3 grid(n_layers, (B + SC_ROWS - 1) / SC_ROWS);
  modulation_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, reinterpret_cast<const ModLayer*>(layers), styles,
                                                                 styles_bstride, gl, gl_bstride, B);
  SPGAN_CHECK_LAUNCH("spgan_modulation_batch");
  return 0;
}

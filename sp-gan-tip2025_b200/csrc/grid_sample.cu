// General 2-D grid samplers with their TRUE input gradients (the samplers of the reference that the spgan.yaml generator
// does not instantiate but whose signatures north_star pins):
//   mode 0  F.grid_sample(bilinear, border, align_corners=True)          GridSamplerNew          grid_generator.py:588-592
//   mode 1  grid_sample_github: bilinear weights from the UNCLIPPED coordinate, corner indices clamped (so a grid that
//           overshoots [-1, 1] extrapolates)                              GridSamplerNewTexture   grid_sample_ops.py:5-55
//   mode 2  F.grid_sample(nearest, zeros, align_corners=True)             GridSampler             grid_sample_grad_fix.py:29-48
// forward: out[b, c, oy, ox] = sum_corner w * z[b, c, y_corner, x_corner]; backward: the transposed scatter
// (aten::grid_sampler_2d_backward's grad_input for modes 0 / 2, autograd of torch.gather for mode 1); the op is linear in z,
// so its double backward is the forward again.  Index arithmetic as in sphere_gather.cu: ATen's fp32 sequence with
// round-to-nearest intrinsics.  HBM-bound: 4*B*C*(IH*IW + OH*OW) bytes + the grid.
#include "common.cuh"

namespace {

struct Taps {
  int o[4];    // pixel offsets y*W + x of the corners (or -1: contributes nothing)
  float w[4];
};

__device__ __forceinline__ float unnorm(float g, int size) { return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1)); }

template <int kMode>
__device__ __forceinline__ Taps make_taps(float gx, float gy, int H, int W) {
  Taps t;
  float ix = unnorm(gx, W), iy = unnorm(gy, H);
  if (kMode == 2) {
    // nearest, zeros padding: nearbyint (round half to even), nothing outside the image
    const float rx = nearbyintf(ix), ry = nearbyintf(iy);
    const bool inside = rx >= 0.f && rx <= (float)(W - 1) && ry >= 0.f && ry <= (float)(H - 1);
    t.o[0] = inside ? (int)ry * W + (int)rx : -1;
    t.w[0] = inside ? 1.f : 0.f;
    t.o[1] = t.o[2] = t.o[3] = -1;
    t.w[1] = t.w[2] = t.w[3] = 0.f;
    return t;
  }
  if (kMode == 0) {
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  }
  const float fx = floorf(ix), fy = floorf(iy);
  // weights: (x_se - ix) * (y_se - iy) etc. with x_se = fx + 1 (both references compute them from the unclamped corners)
  const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
  int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  if (kMode == 0) {
    x1 = min(x1, W - 1);
    y1 = min(y1, H - 1);
  } else {
    x0 = min(max(x0, 0), W - 1);
    x1 = min(max(x1, 0), W - 1);
    y0 = min(max(y0, 0), H - 1);
    y1 = min(max(y1, 0), H - 1);
  }
  t.o[0] = y0 * W + x0;
  t.o[1] = y0 * W + x1;
  t.o[2] = y1 * W + x0;
  t.o[3] = y1 * W + x1;
  t.w[0] = ex * ey;
  t.w[1] = wx * ey;
  t.w[2] = ex * wy;
  t.w[3] = wx * wy;
  return t;
}

constexpr int GS_CCHUNK = 16;

// One thread per (sample, output pixel, chunk of 16 channels): the taps are computed once and reused over the chunk;
// consecutive threads read consecutive grid entries and write consecutive output pixels.
template <int kMode, bool kBackward>
__global__ void __launch_bounds__(256) grid_sample_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                         const float* __restrict__ grid, int B, int C, int IH, int IW,
                                                         int OH, int OW, int grid_batch) {
  const int64_t opix = (int64_t)OH * OW, ipix = (int64_t)IH * IW;
  const int64_t total = (int64_t)B * opix;
  const int c_begin = blockIdx.y * GS_CCHUNK;
  const int c_end = min(c_begin + GS_CCHUNK, C);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / opix);
    const int64_t pix = idx - (int64_t)b * opix;
    const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + (int64_t)(grid_batch == 1 ? 0 : b) * opix + pix);
    const Taps t = make_taps<kMode>(g.x, g.y, IH, IW);
    if (!kBackward) {
      const float* zp = src + ((int64_t)b * C + c_begin) * ipix;
      float* op = dst + ((int64_t)b * C + c_begin) * opix + pix;
      for (int c = c_begin; c < c_end; ++c) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kMode != 2 || k == 0) v += (t.o[k] >= 0 ? __ldg(zp + t.o[k]) : 0.f) * t.w[k];
        *op = v;
        zp += ipix;
        op += opix;
      }
    } else {
      // dst = grad wrt z (zero-initialised by the caller), src = grad wrt the sampled output
      float* zp = dst + ((int64_t)b * C + c_begin) * ipix;
      const float* op = src + ((int64_t)b * C + c_begin) * opix + pix;
      for (int c = c_begin; c < c_end; ++c) {
        const float gv = __ldg(op);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if ((kMode != 2 || k == 0) && t.o[k] >= 0 && t.w[k] != 0.f) atomicAdd(zp + t.o[k], gv * t.w[k]);
        zp += ipix;
        op += opix;
      }
    }
  }
}

template <bool kBackward>
int launch(float* dst, const float* src, const float* grid, int B, int C, int IH, int IW, int OH, int OW, int grid_batch,
           int mode, cudaStream_t st) {
  const int64_t total = (int64_t)B * OH * OW;
  dim3 g(grid_for(total, 256, 8, 2), (C + GS_CCHUNK - 1) / GS_CCHUNK);
  if (mode == 0) grid_sample_kernel<0, kBackward><<<g, 256, 0, st>>>(dst, src, grid, B, C, IH, IW, OH, OW, grid_batch);
  else if (mode == 1) grid_sample_kernel<1, kBackward><<<g, 256, 0, st>>>(dst, src, grid, B, C, IH, IW, OH, OW, grid_batch);
  else grid_sample_kernel<2, kBackward><<<g, 256, 0, st>>>(dst, src, grid, B, C, IH, IW, OH, OW, grid_batch);
  return 0;
}

int check(const void* a, const void* b, const void* grid, int B, int C, int IH, int IW, int OH, int OW, int grid_batch, int mode,
          const char* who) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && IH >= 0 && IW >= 0 && OH >= 0 && OW >= 0, "%s: negative size", who);
  SPGAN_CHECK_ARG(mode >= 0 && mode <= 2, "%s: mode must be 0 (bilinear/border), 1 (texture) or 2 (nearest/zeros), got %d", who, mode);
  SPGAN_CHECK_ARG(grid_batch == 1 || grid_batch == B, "%s: grid batch %d must be 1 or %d", who, grid_batch, B);
  SPGAN_CHECK_ARG((int64_t)IH * IW < (1LL << 31), "%s: image too large", who);
  if (B == 0 || C == 0 || OH == 0 || OW == 0) return -1;
  SPGAN_CHECK_ARG(IH > 0 && IW > 0, "%s: empty input image", who);
  SPGAN_CHECK_ARG(a && b && grid, "%s: null pointer", who);
  SPGAN_CHECK_ARG((((uintptr_t)grid) & 7) == 0, "%s: grid must be 8-byte aligned", who);
  return 0;
}

}  // namespace

extern "C" int spgan_grid_sample(float* out, const float* z, const float* grid, int B, int C, int IH, int IW, int OH, int OW,
                                 int grid_batch, int mode, void* stream) {
  const int rc = check(out, z, grid, B, C, IH, IW, OH, OW, grid_batch, mode, "spgan_grid_sample");
  if (rc) return rc < 0 ? 0 : rc;
  launch<false>(out, z, grid, B, C, IH, IW, OH, OW, grid_batch, mode, (cudaStream_t)stream);
  SPGAN_CHECK_LAUNCH("spgan_grid_sample");
  return 0;
}

extern "C" int spgan_grid_sample_bwd(float* grad_z, const float* grad_out, const float* grid, int B, int C, int IH, int IW,
                                     int OH, int OW, int grid_batch, int mode, void* stream) {
  const int rc = check(grad_z, grad_out, grid, B, C, IH, IW, OH, OW, grid_batch, mode, "spgan_grid_sample_bwd");
  if (rc) return rc < 0 ? 0 : rc;
  SPGAN_CUDA(cudaMemsetAsync(grad_z, 0, (size_t)B * C * IH * IW * sizeof(float), (cudaStream_t)stream), "spgan_grid_sample_bwd");
  launch<true>(grad_z, grad_out, grid, B, C, IH, IW, OH, OW, grid_batch, mode, (cudaStream_t)stream);
  SPGAN_CHECK_LAUNCH("spgan_grid_sample_bwd");
  return 0;
}

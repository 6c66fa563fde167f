// Corner indices and bilinear weights of one spherical tap, shared by the operand packer (conv_umma.cu) and the fused
// gather-GEMM (sphere_umma.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// ATen's fp32 index sequence (GridSampler.h:27-36, 58-60) with explicit round-to-nearest ops (no FMA contraction).
__device__ __forceinline__ float unnorm_clip(float g, int size) {
  float v = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1));
  return fminf((float)(size - 1), fmaxf(v, 0.f));
}

struct TapCorners {
  int o_nw, o_ne, o_sw, o_se;  // pixel offsets y*W + x of the four corners
  float w_nw, w_ne, w_sw, w_se;
};

__device__ __forceinline__ TapCorners tap_corners(const float* __restrict__ grid, int bg, int H, int W, int py, int px,
                                                  int ty, int tx) {
  const float2 gxy =
      __ldg(reinterpret_cast<const float2*>(grid) + ((int64_t)bg * 3 * H + (3 * py + ty)) * (3 * W) + 3 * px + tx);
  const float ix = unnorm_clip(gxy.x, W), iy = unnorm_clip(gxy.y, H);
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
  TapCorners c;
  c.o_nw = y0 * W + x0;
  c.o_ne = y0 * W + x1;
  c.o_sw = y1 * W + x0;
  c.o_se = y1 * W + x1;
  c.w_nw = ex * ey;
  c.w_ne = wx * ey;
  c.w_sw = ex * wy;
  c.w_se = wx * wy;
  return c;
}

}  // namespace

// L1: SphereNet-style distortion-aware gather — bilinear sampling of a (B, C, H, W) feature at a 3x3 tap grid
// (B|1, 3H, 3W, 2), border padding, align_corners=True — its index-only variant, and the reference's surrogate
// backward (3x3 block mean * 0.1).  Index arithmetic is ATen's fp32 sequence, evaluated with explicit
// round-to-nearest intrinsics so that no FMA contraction can change a floor():
//     ix = ((gx + 1) / 2) * (W - 1);  ix = min(W - 1, max(ix, 0));  x0 = floor(ix)
// HBM roofline: writes 9x the input (4*B*C*9*H*W bytes) + reads input once (the 36 corner reads per input pixel
// hit L1/L2) + the grid.
#include "common.cuh"

namespace {

struct Corner {
  int x0, y0, x1, y1;
  float nw, ne, sw, se;
};

__device__ __forceinline__ float unnormalize_clip(float g, int size) {
  float v = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), (float)(size - 1));
  return fminf((float)(size - 1), fmaxf(v, 0.f));
}

__device__ __forceinline__ Corner corners(float gx, float gy, int H, int W) {
  Corner c;
  const float ix = unnormalize_clip(gx, W), iy = unnormalize_clip(gy, H);
  const float fx = floorf(ix), fy = floorf(iy);
  c.x0 = (int)fx;
  c.y0 = (int)fy;
  c.x1 = min(c.x0 + 1, W - 1);
  c.y1 = min(c.y0 + 1, H - 1);
  const float ex = __fsub_rn(__fadd_rn(fx, 1.f), ix), ey = __fsub_rn(__fadd_rn(fy, 1.f), iy);  // weight of the low corner
  const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
  c.nw = ex * ey;
  c.ne = wx * ey;
  c.sw = ex * wy;
  c.se = wx * wy;
  return c;
}

constexpr int GATHER_CCHUNK = 16;

__global__ void __launch_bounds__(256) sphere_gather_kernel(float* __restrict__ out, const float* __restrict__ z,
                                                           const float* __restrict__ grid, int B, int C, int H, int W,
                                                           int grid_batch, int64_t out_bstride, int64_t out_coff,
                                                           int encode) {
  const int OH = 3 * H, OW = 3 * W;
  const int64_t opix = (int64_t)OH * OW;
  const int64_t total = (int64_t)B * opix;
  const int c_begin = blockIdx.y * GATHER_CCHUNK;
  const int c_end = min(c_begin + GATHER_CCHUNK, C);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / opix);
    const int64_t pix = idx - (int64_t)b * opix;
    const int bg = grid_batch == 1 ? 0 : b;
    const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + (int64_t)bg * opix + pix);
    const Corner k = corners(g.x, g.y, H, W);
    const int o_nw = k.y0 * W + k.x0, o_ne = k.y0 * W + k.x1, o_sw = k.y1 * W + k.x0, o_se = k.y1 * W + k.x1;
    const float* zp = z + ((int64_t)b * C + c_begin) * H * W;
    float* op = out + ((int64_t)b * out_bstride + out_coff + c_begin) * opix + pix;
#pragma unroll 4
    for (int c = c_begin; c < c_end; ++c) {
      float v = __ldg(zp + o_nw) * k.nw + __ldg(zp + o_ne) * k.ne + __ldg(zp + o_sw) * k.sw + __ldg(zp + o_se) * k.se;
      if (encode) {
        if (c == 0) v = tanhf(v);
        else if (c == 1) v = cosf(v * 3.14159274101257324f);
        else if (c == 2) v = sinf(v * 3.14159274101257324f);
      }
      __stcs(op, v);
      zp += (int64_t)H * W;
      op += opix;
    }
  }
}

__global__ void __launch_bounds__(256) sphere_indices_kernel(int32_t* __restrict__ x0, int32_t* __restrict__ y0,
                                                            float* __restrict__ wx, float* __restrict__ wy,
                                                            const float* __restrict__ grid, int64_t n, int H, int W) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 g = reinterpret_cast<const float2*>(grid)[i];
    const float ix = unnormalize_clip(g.x, W), iy = unnormalize_clip(g.y, H);
    const float fx = floorf(ix), fy = floorf(iy);
    x0[i] = (int)fx;
    y0[i] = (int)fy;
    wx[i] = __fsub_rn(ix, fx);
    wy[i] = __fsub_rn(iy, fy);
  }
}

__global__ void __launch_bounds__(256) sphere_gather_bwd_kernel(float* __restrict__ gi, const float* __restrict__ go,
                                                               int64_t planes, int H, int W) {
  const int64_t total = planes * H * W;
  const int OW = 3 * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = idx / ((int64_t)H * W);
    const int rem = (int)(idx - plane * H * W);
    const int y = rem / W, x = rem - y * W;
    const float* g = go + plane * 9 * H * W + (int64_t)(3 * y) * OW + 3 * x;
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) s += __ldcs(g + r * OW) + __ldcs(g + r * OW + 1) + __ldcs(g + r * OW + 2);
    gi[idx] = __fmul_rn(__fdiv_rn(s, 9.f), 0.1f);
  }
}

// Assemble per-sample training grids on the device from the two factors the grid separates into:
// the latitude part and the tangent-plane longitude offsets depend only on the rows of the window (p_x_st / p_x_ed),
// the normalised column base only on its columns (p_y_st / p_y_ed / circular flag).  The host keeps the float64 factor
// tables it computed with numpy (bit-exact libm), this kernel redoes the reference's remaining float64 +, /, * sequence
// (models/spherenet/grid_generator.py:270-283, models/spgan_ops_gs.py:416-423) with explicit round-to-nearest
// intrinsics, so the result equals the host-built grid bit for bit without a host round trip per sample.
//   out[b, 3y+ky, 3x+kx, 0] = float(((((lon[ix[b]][y][k] + nlon[iy[b]][x]) / 2 + 0.5) * y_total) / y_total) * 2 - 1)
//   out[b, 3y+ky, 3x+kx, 1] = lat_n[ix[b]][y][k]
__global__ void __launch_bounds__(256) grid_assemble_kernel(float* __restrict__ out, const float* __restrict__ lat_n,
                                                           const double* __restrict__ lon, const double* __restrict__ nlon,
                                                           const int* __restrict__ ix, const int* __restrict__ iy, int B,
                                                           int H, int W, double y_total) {
  const int OW = 3 * W;
  const int64_t per = (int64_t)9 * H * W;
  const int64_t total = (int64_t)B * per;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per);
    const int r = (int)(idx - (int64_t)b * per);
    const int oy = r / OW, ox = r - oy * OW;
    const int y = oy / 3, ky = oy - 3 * y, x = ox / 3, kx = ox - 3 * x;
    const int sx = __ldg(ix + b), sy = __ldg(iy + b);
    const int64_t t = ((int64_t)sx * H + y) * 9 + ky * 3 + kx;
    const double a = __dadd_rn(__ldg(lon + t), __ldg(nlon + (int64_t)sy * W + x));
    const double g = __dmul_rn(__dadd_rn(__ddiv_rn(a, 2.0), 0.5), y_total);
    const double n = __dsub_rn(__dmul_rn(__ddiv_rn(g, y_total), 2.0), 1.0);
    float2 v;
    v.x = __double2float_rn(n);
    v.y = __ldg(lat_n + t);
    reinterpret_cast<float2*>(out)[idx] = v;
  }
}

}  // namespace

extern "C" int spgan_sphere_grid_assemble(float* out, const float* lat_n, const double* lon, const double* nlon,
                                          const int32_t* ix, const int32_t* iy, int B, int H, int W, double y_total,
                                          void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && H >= 0 && W >= 0, "spgan_sphere_grid_assemble: negative size");
  if (B == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && lat_n && lon && nlon && ix && iy, "spgan_sphere_grid_assemble: null pointer");
  SPGAN_CHECK_ARG((((uintptr_t)out) & 7) == 0, "spgan_sphere_grid_assemble: out must be 8-byte aligned");
  const int64_t total = (int64_t)B * 9 * H * W;
  grid_assemble_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, lat_n, lon, nlon, ix, iy, B, H, W, y_total);
  SPGAN_CHECK_LAUNCH("spgan_sphere_grid_assemble");
  return 0;
}

namespace {
}  // namespace

extern "C" int spgan_sphere_gather(float* out, const float* z, const float* grid, int B, int C, int H, int W,
                                   int grid_batch, int64_t out_bstride, int64_t out_coff, int encode, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "spgan_sphere_gather: negative size");
  if (B == 0 || C == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && z && grid, "spgan_sphere_gather: null pointer");
  SPGAN_CHECK_ARG(grid_batch == 1 || grid_batch == B, "spgan_sphere_gather: grid batch %d must be 1 or %d", grid_batch, B);
  SPGAN_CHECK_ARG(!encode || C == 3, "spgan_sphere_gather: coordinate encoding expects 3 channels, got %d", C);
  SPGAN_CHECK_ARG((((uintptr_t)grid) & 7) == 0, "spgan_sphere_gather: grid must be 8-byte aligned");
  const int64_t total = (int64_t)B * 9 * H * W;
  dim3 g(grid_for(total, 256, 8, 2), (C + GATHER_CCHUNK - 1) / GATHER_CCHUNK);
  sphere_gather_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(out, z, grid, B, C, H, W, grid_batch, out_bstride, out_coff,
                                                           encode);
  SPGAN_CHECK_LAUNCH("spgan_sphere_gather");
  return 0;
}

extern "C" int spgan_sphere_gather_indices(int32_t* x0, int32_t* y0, float* wx, float* wy, const float* grid, int64_t n,
                                           int H, int W, void* stream) {
  if (n <= 0) return 0;
  SPGAN_CHECK_ARG(x0 && y0 && wx && wy && grid, "spgan_sphere_gather_indices: null pointer");
  sphere_indices_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(x0, y0, wx, wy, grid, n, H, W);
  SPGAN_CHECK_LAUNCH("spgan_sphere_gather_indices");
  return 0;
}

extern "C" int spgan_sphere_gather_bwd(float* grad_in, const float* grad_out, int64_t planes, int H, int W,
                                       void* stream) {
  if (planes <= 0 || H <= 0 || W <= 0) return 0;
  SPGAN_CHECK_ARG(grad_in && grad_out, "spgan_sphere_gather_bwd: null pointer");
  sphere_gather_bwd_kernel<<<grid_for(planes * H * W, 256, 8), 256, 0, (cudaStream_t)stream>>>(grad_in, grad_out, planes,
                                                                                            H, W);
  SPGAN_CHECK_LAUNCH("spgan_sphere_gather_bwd");
  return 0;
}

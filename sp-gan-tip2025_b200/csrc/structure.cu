// Operand producers of the structure synthesiser's channels-last inference chain (models/spgan/spgan.py:79-169, 172-254).
//
// Both convs of a structure block see 256 feature channels + 3 coordinate channels (259).  Padding 259 to whole 64-wide
// K blocks per tap (320) spends 19 % of the tensor-core work and of the operand bytes on zeros, for 9 taps (spherical conv)
// and 49 taps (7x7 conv).  Here the 256 features keep kp = 256 per tap and the three trailing channels of ALL taps go into
// one dense second K segment (spgan_conv_gemm_ex: a2_packed / w2_packed):
//   sphere_pack_seg : the spherical gather producer (bilinear border gather at the 3x3 tangent taps, coordinate encoding,
//                     the reference's flat (1,B*C)++(1,B*3) concat table, style modulation, 16-bit split) writing the main
//                     operand [2][B*H*W][9*Cm] and the tail operand [2][B*H*W][kp2], k2 = tap*Cx + j.  One sampling grid per
//                     GROUP of samples (grid_group): several lattice positions of a panorama run as one batch.
//   coord_taps_pack : the tail operand of the 7x7 conv: tanh / cos pi / sin pi of the raw coordinate planes (coord_handler.py:
//                     696-711) times the style modulation, im2col'ed over the kh*kw taps, [2][B*My*Mx][kp2], k2 = tap*nc + j.
// Roofline: HBM (sphere_pack_seg writes 4*9*Cm + 4*kp2 bytes per pixel and gathers 4x that from L2).
#include "sphere_taps.cuh"
#include "umma_common.cuh"

namespace {

template <int KITER, bool kF16>
__global__ void __launch_bounds__(256, 2) sphere_pack_seg_kernel(uint16_t* __restrict__ out, uint16_t* __restrict__ out2,
                                                                const float* __restrict__ xh,
                                                                const float* __restrict__ coords,
                                                                const float* __restrict__ grid,
                                                                const float* __restrict__ in_mul,
                                                                const uint32_t* __restrict__ chan_map, int B, int C, int nc,
                                                                int H, int W, int grid_group, int cmap_ld, int kp2) {
  constexpr int Cm = 64 * KITER;
  const int Ct = C + nc;
  const int Cx = Ct - Cm;  // trailing channels of every group that go to the second segment (0..31)
  const int HW = H * W;
  const int64_t plane_elems = (int64_t)B * HW * 9 * Cm;
  const int64_t plane2_elems = (int64_t)B * HW * kp2;
  const int64_t warps_total = (int64_t)B * HW;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warp_stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wid = warp0; wid < warps_total; wid += warp_stride) {
    const int p = (int)(wid % HW);
    const int g = (int)(wid / HW);
    const int py = p / W, px = p - py * W;
    TapCorners mine;
    mine.o_nw = mine.o_ne = mine.o_sw = mine.o_se = 0;
    mine.w_nw = mine.w_ne = mine.w_sw = mine.w_se = 0.f;
    if (lane < 9) mine = tap_corners(grid, g / grid_group, H, W, py, px, lane / 3, lane % 3);
    const uint32_t* mrow = chan_map + (int64_t)g * cmap_ld;
    const float* mulrow = in_mul ? in_mul + (int64_t)g * Ct : nullptr;
    uint32_t mw[KITER][2];
    uint32_t soff[KITER][2];
    float mv[KITER][2];
#pragma unroll
    for (int j = 0; j < KITER; ++j) {
      const uint2 mm = __ldg(reinterpret_cast<const uint2*>(mrow + 2 * lane + 64 * j));
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t m = u ? mm.y : mm.x;
        const bool valid = m != 0xFFFFFFFFu;
        const uint32_t bs = (m >> 15) & 0xFFFFu, cs = m & 0x7FFFu;
        mw[j][u] = m;
        soff[j][u] = !valid ? 0u : ((m >> 31) ? (bs * (uint32_t)nc + cs) * (uint32_t)HW : bs * (uint32_t)HW * (uint32_t)C + cs);
        mv[j][u] = (valid && mulrow) ? __ldg(mulrow + 2 * lane + 64 * j + u) : 1.f;
      }
    }
    // this lane's tail channel (lanes 0..Cx-1)
    uint32_t mwx = 0xFFFFFFFFu, soffx = 0u;
    float mvx = 1.f;
    if (lane < Cx) {
      mwx = __ldg(mrow + Cm + lane);
      const bool valid = mwx != 0xFFFFFFFFu;
      const uint32_t bs = (mwx >> 15) & 0xFFFFu, cs = mwx & 0x7FFFu;
      soffx = !valid ? 0u : ((mwx >> 31) ? (bs * (uint32_t)nc + cs) * (uint32_t)HW : bs * (uint32_t)HW * (uint32_t)C + cs);
      mvx = (valid && mulrow) ? __ldg(mulrow + Cm + lane) : 1.f;
    }
    uint16_t* obase = out + wid * 9 * Cm;
    uint16_t* obase2 = out2 ? out2 + wid * kp2 : nullptr;
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {
      TapCorners cn;
      cn.o_nw = __shfl_sync(0xffffffffu, mine.o_nw, t);
      cn.o_ne = __shfl_sync(0xffffffffu, mine.o_ne, t);
      cn.o_sw = __shfl_sync(0xffffffffu, mine.o_sw, t);
      cn.o_se = __shfl_sync(0xffffffffu, mine.o_se, t);
      cn.w_nw = __shfl_sync(0xffffffffu, mine.w_nw, t);
      cn.w_ne = __shfl_sync(0xffffffffu, mine.w_ne, t);
      cn.w_sw = __shfl_sync(0xffffffffu, mine.w_sw, t);
      cn.w_se = __shfl_sync(0xffffffffu, mine.w_se, t);
      float cv[KITER][2][4];
#pragma unroll
      for (int j = 0; j < KITER; ++j)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint32_t m = mw[j][u];
          const bool valid = m != 0xFFFFFFFFu;
          const bool is_coord = (m >> 31) != 0;
          const float* sp = (is_coord ? coords : xh) + soff[j][u];
          const int st = is_coord ? 1 : C;
          cv[j][u][0] = valid ? __ldg(sp + cn.o_nw * st) : 0.f;
          cv[j][u][1] = valid ? __ldg(sp + cn.o_ne * st) : 0.f;
          cv[j][u][2] = valid ? __ldg(sp + cn.o_sw * st) : 0.f;
          cv[j][u][3] = valid ? __ldg(sp + cn.o_se * st) : 0.f;
        }
      float xv[4] = {0.f, 0.f, 0.f, 0.f};
      const bool xvalid = mwx != 0xFFFFFFFFu;
      const bool xcoord = (mwx >> 31) != 0;
      if (xvalid) {
        const float* sp = (xcoord ? coords : xh) + soffx;
        const int st = xcoord ? 1 : C;
        xv[0] = __ldg(sp + cn.o_nw * st);
        xv[1] = __ldg(sp + cn.o_ne * st);
        xv[2] = __ldg(sp + cn.o_sw * st);
        xv[3] = __ldg(sp + cn.o_se * st);
      }
      uint16_t* orow = obase + t * Cm;
#pragma unroll
      for (int j = 0; j < KITER; ++j) {
        float v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          float val = cv[j][u][0] * cn.w_nw + cv[j][u][1] * cn.w_ne + cv[j][u][2] * cn.w_sw + cv[j][u][3] * cn.w_se;
          const uint32_t m = mw[j][u];
          if (m != 0xFFFFFFFFu && (m >> 31) != 0) {
            const uint32_t cs = m & 0x7FFFu;
            if (cs == 0) val = tanhf(val);
            else if (cs == 1) val = cosf(val * 3.14159274101257324f);
            else if (cs == 2) val = sinf(val * 3.14159274101257324f);
          }
          v[u] = val * mv[j][u];
        }
        uint16_t h0, l0, h1, l1;
        split16<kF16>(v[0], h0, l0);
        split16<kF16>(v[1], h1, l1);
        const int k0 = 2 * lane + 64 * j;
        *reinterpret_cast<uint32_t*>(orow + k0) = pack2x16(h0, h1);
        *reinterpret_cast<uint32_t*>(orow + plane_elems + k0) = pack2x16(l0, l1);
      }
      if (lane < Cx) {
        float val = xv[0] * cn.w_nw + xv[1] * cn.w_ne + xv[2] * cn.w_sw + xv[3] * cn.w_se;
        if (xvalid && xcoord) {
          const uint32_t cs = mwx & 0x7FFFu;
          if (cs == 0) val = tanhf(val);
          else if (cs == 1) val = cosf(val * 3.14159274101257324f);
          else if (cs == 2) val = sinf(val * 3.14159274101257324f);
        }
        val *= mvx;
        uint16_t h, l;
        split16<kF16>(val, h, l);
        obase2[t * Cx + lane] = h;
        obase2[plane2_elems + t * Cx + lane] = l;
      }
    }
    if (obase2)
      for (int c = 9 * Cx + lane; c < kp2; c += 32) {
        obase2[c] = 0;
        obase2[plane2_elems + c] = 0;
      }
  }
}

// One thread per (output row, column of the tail operand).
template <bool kF16>
__global__ void __launch_bounds__(256) coord_taps_pack_kernel(uint16_t* __restrict__ out2, const float* __restrict__ coords,
                                                             const float* __restrict__ in_mul, int64_t rows, int nc, int H,
                                                             int W, int kw, int ntaps, int My, int Mx, int mul_ld, int mul_off,
                                                             int kp2) {
  const int64_t total = rows * kp2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % kp2);
    const int64_t row = idx / kp2;
    uint16_t h = 0, l = 0;
    if (col < ntaps * nc) {
      const int t = col / nc, j = col - t * nc;
      const int ty = t / kw, tx = t - ty * kw;
      const int x = (int)(row % Mx);
      const int64_t r = row / Mx;
      const int y = (int)(r % My);
      const int b = (int)(r / My);
      float v = __ldg(coords + (((int64_t)b * nc + j) * H + (y + ty)) * W + (x + tx));
      if (j == 0) v = tanhf(v);
      else if (j == 1) v = cosf(v * 3.14159274101257324f);
      else if (j == 2) v = sinf(v * 3.14159274101257324f);
      if (in_mul) v *= __ldg(in_mul + (int64_t)b * mul_ld + mul_off + j);
      split16<kF16>(v, h, l);
    }
    out2[idx] = h;
    out2[total + idx] = l;
  }
}

}  // namespace

extern "C" int spgan_sphere_pack_seg(uint16_t* out, uint16_t* out2, const float* x_nhwc, const float* coords,
                                     const float* grid, const float* in_mul, const uint32_t* chan_map, int B, int C, int H,
                                     int W, int grid_group, int Cm, int cmap_ld, int kp2, int fmt, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "spgan_sphere_pack_seg: negative size");
  const int nc = coords ? 3 : 0;
  const int Cx = C + nc - Cm;
  SPGAN_CHECK_ARG(Cm > 0 && Cm % 64 == 0 && Cm <= 320 && Cx >= 0 && Cx < 32,
                  "spgan_sphere_pack_seg: Cm=%d must be a multiple of 64 (<= 320) with 0 <= C + nc - Cm < 32, got C + nc = %d", Cm, C + nc);
  SPGAN_CHECK_ARG(kp2 % 64 == 0 && kp2 >= 9 * Cx && (Cx == 0) == (kp2 == 0),
                  "spgan_sphere_pack_seg: kp2=%d must be a multiple of 64, >= 9 * %d tail channels (0 iff there are none)", kp2, Cx);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_sphere_pack_seg: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  if (B == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && x_nhwc && grid && chan_map && (out2 || kp2 == 0), "spgan_sphere_pack_seg: null pointer");
  SPGAN_CHECK_ARG(B <= 65535 && C <= 32767, "spgan_sphere_pack_seg: B=%d / C=%d exceed the channel-map encoding", B, C);
  SPGAN_CHECK_ARG(grid_group >= 1 && B % grid_group == 0, "spgan_sphere_pack_seg: grid_group=%d must divide the batch %d", grid_group, B);
  SPGAN_CHECK_ARG(cmap_ld >= C + nc && cmap_ld % 2 == 0, "spgan_sphere_pack_seg: chan_map row stride %d too small / odd", cmap_ld);
  SPGAN_CHECK_ARG((int64_t)B * H * W * (C > 3 ? C : 3) < (1LL << 31), "spgan_sphere_pack_seg: input too large for 32-bit plane offsets");
  SPGAN_CHECK_ARG((((uintptr_t)grid) & 7) == 0 && (((uintptr_t)chan_map) & 7) == 0,
                  "spgan_sphere_pack_seg: grid and chan_map must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = grid_for((int64_t)B * H * W, 8, 2, 16);
#define SPGAN_SEG(KI)                                                                                                          \
  do {                                                                                                                         \
    if (fmt) sphere_pack_seg_kernel<KI, true><<<nblk, 256, 0, st>>>(out, out2, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W, grid_group, cmap_ld, kp2); \
    else sphere_pack_seg_kernel<KI, false><<<nblk, 256, 0, st>>>(out, out2, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W, grid_group, cmap_ld, kp2);    \
  } while (0)
  switch (Cm / 64) {
    case 1: SPGAN_SEG(1); break;
    case 2: SPGAN_SEG(2); break;
    case 3: SPGAN_SEG(3); break;
    case 4: SPGAN_SEG(4); break;
    default: SPGAN_SEG(5); break;
  }
#undef SPGAN_SEG
  SPGAN_CHECK_LAUNCH("spgan_sphere_pack_seg");
  return 0;
}

extern "C" int spgan_coord_taps_pack(uint16_t* out2, const float* coords, const float* in_mul, int B, int nc, int H, int W,
                                     int kh, int kw, int mul_ld, int mul_off, int kp2, int fmt, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && nc >= 1 && nc <= 3 && H >= kh && W >= kw && kh >= 1 && kw >= 1,
                  "spgan_coord_taps_pack: bad geometry (B=%d nc=%d %dx%d, %dx%d taps)", B, nc, H, W, kh, kw);
  SPGAN_CHECK_ARG(kp2 % 64 == 0 && kp2 >= kh * kw * nc, "spgan_coord_taps_pack: kp2=%d must be a multiple of 64 and >= %d", kp2, kh * kw * nc);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_coord_taps_pack: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  const int My = H - kh + 1, Mx = W - kw + 1;
  const int64_t rows = (int64_t)B * My * Mx;
  if (rows == 0) return 0;
  SPGAN_CHECK_ARG(out2 && coords, "spgan_coord_taps_pack: null pointer");
  SPGAN_CHECK_ARG(in_mul == nullptr || (mul_ld >= mul_off + nc && mul_off >= 0), "spgan_coord_taps_pack: modulation slice out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = grid_for(rows * kp2, 256, 4);
  if (fmt)
    coord_taps_pack_kernel<true><<<nblk, 256, 0, st>>>(out2, coords, in_mul, rows, nc, H, W, kw, kh * kw, My, Mx, mul_ld, mul_off, kp2);
  else
    coord_taps_pack_kernel<false><<<nblk, 256, 0, st>>>(out2, coords, in_mul, rows, nc, H, W, kw, kh * kw, My, Mx, mul_ld, mul_off, kp2);
  SPGAN_CHECK_LAUNCH("spgan_coord_taps_pack");
  return 0;
}

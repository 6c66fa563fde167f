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
#include "pack_math.cuh"

#include <stdlib.h>

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

// ---- vectorised producer for the 256-feature configuration -----------------------------------------------------------------
// sphere_pack_seg_kernel above is instruction-issue bound (ncu: ~210 warp instructions per (pixel, tap), IPC 1.9, 19 % of the
// HBM write roofline): 32 scalar gathers with 32 address computations, scalar blends and 2-byte splits per lane and tap.  The
// reference's flat (1,B*C)++(1,B*3) concatenation makes group g read channels [259 g, 259 g + 259) of the flat list, so its K
// columns are neither aligned to 16 bytes in the NHWC source nor confined to one sample; a first vectorised version that
// realigned in registers (3 loads per corner, per-lane boundary paths) was SLOWER (347 instructions per task).  Hence two steps:
//   concat_repack : xg[g][pixel][k] = the k-th flat-concat channel of group g, raw (features and raw coordinate planes), rows
//                   of SPV_LD floats (16-byte aligned) — the reference's concat made explicit once per layer (2 x 80 MB at
//                   35 x 35, B = 64, against 742 MB of operand written);
//   sphere_pack_v3: one warp per (pixel, tap), lane = 8 consecutive K columns: two aligned 128-bit loads per corner, packed
//                   FFMA2 blend, coordinate encoding only in the lanes of the one group per call that holds coordinate
//                   planes, packed bf16x2 / f16x2 conversions, one 16-byte store per plane.  A CTA owns SP_PX pixels of ONE
//                   group: the corner table of its 9 * SP_PX (pixel, tap) pairs and the group's modulation row sit in smem.
constexpr int SP_PX = 16;
constexpr int SP_TASKS = SP_PX * 9;

// xg[(g*HW + p)*SPV_LD + k] for k < SPV_LD; one thread per element, k fastest (coalesced writes, near-coalesced reads).
__global__ void __launch_bounds__(256) concat_repack_kernel(float* __restrict__ xg, const float* __restrict__ xh,
                                                           const float* __restrict__ coords,
                                                           const uint32_t* __restrict__ chan_map, int C, int nc, int HW,
                                                           int cmap_ld, int px_per_cta) {
  __shared__ uint32_t s_map[SPV_LD];
  const int g = blockIdx.y;
  const int Ct = C + nc;
  for (int k = threadIdx.x; k < SPV_LD; k += blockDim.x) s_map[k] = k < Ct ? __ldg(chan_map + (int64_t)g * cmap_ld + k) : 0xFFFFFFFFu;
  __syncthreads();
  const int p0 = blockIdx.x * px_per_cta;
  const int npx = min(px_per_cta, HW - p0);
  for (int idx = threadIdx.x; idx < npx * SPV_LD; idx += blockDim.x) {
    const int r = idx / SPV_LD, k = idx - r * SPV_LD;
    const uint32_t m = s_map[k];
    float v = 0.f;
    if (m != 0xFFFFFFFFu) {
      const uint32_t bs = (m >> 15) & 0xFFFFu, cs = m & 0x7FFFu;
      v = (m >> 31) ? __ldg(coords + ((int64_t)bs * nc + cs) * HW + p0 + r) : __ldg(xh + ((int64_t)bs * HW + p0 + r) * C + cs);
    }
    xg[((int64_t)g * HW + p0 + r) * SPV_LD + k] = v;
  }
}

template <bool kF16>
__global__ void __launch_bounds__(256) sphere_pack_v3_kernel(uint16_t* __restrict__ out, uint16_t* __restrict__ out2,
                                                            const float* __restrict__ xg, const float* __restrict__ grid,
                                                            const float* __restrict__ in_mul,
                                                            const uint32_t* __restrict__ chan_map, int B, int Ct, int H, int W,
                                                            int grid_group, int cmap_ld, int kp2, int blocks_per_group) {
  constexpr int C = SPV_C;
  __shared__ int s_off[4][SP_TASKS];
  __shared__ float s_wt[4][SP_TASKS];
  __shared__ __align__(16) float s_mul[SPV_LD];
  const int Cx = Ct - C;
  const int HW = H * W;
  const int g = blockIdx.x / blocks_per_group;
  const int p0 = (blockIdx.x - g * blocks_per_group) * SP_PX;
  const int npx = min(SP_PX, HW - p0);
  const int ntask = npx * 9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < ntask; e += blockDim.x) {
    const int r = e / 9, t = e - r * 9;
    const int p = p0 + r;
    const int py = p / W, px = p - py * W;
    const TapCorners cn = tap_corners(grid, g / grid_group, H, W, py, px, t / 3, t - (t / 3) * 3);
    s_off[0][e] = cn.o_nw * SPV_LD;
    s_off[1][e] = cn.o_ne * SPV_LD;
    s_off[2][e] = cn.o_sw * SPV_LD;
    s_off[3][e] = cn.o_se * SPV_LD;
    s_wt[0][e] = cn.w_nw;
    s_wt[1][e] = cn.w_ne;
    s_wt[2][e] = cn.w_sw;
    s_wt[3][e] = cn.w_se;
  }
  for (int k = threadIdx.x; k < SPV_LD; k += blockDim.x) s_mul[k] = k < Ct ? (in_mul ? __ldg(in_mul + (int64_t)g * Ct + k) : 1.f) : 0.f;
  // this lane's K columns: [4 lane, 4 lane + 4) and [128 + 4 lane, 128 + 4 lane + 4), so that every warp-wide 128-bit load
  // covers 512 contiguous bytes of a source row (8-column runs per lane touched every 32-byte sector twice: ncu L1 77 %).
  // Coordinate-plane columns among them: 2 bits each (0 feature, 1 tanh, 2 cos pi, 3 sin pi).
  const uint32_t* mrow = chan_map + (int64_t)g * cmap_ld;
  const int ka = 4 * lane, kb = C / 2 + 4 * lane;
  uint32_t kinds = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t m = __ldg(mrow + (j < 4 ? ka + j : kb + j - 4));
    if (m != 0xFFFFFFFFu && (m >> 31)) kinds |= (1u + (m & 3u)) << (2 * j);
  }
  __syncthreads();
  const float4 m0 = *reinterpret_cast<const float4*>(&s_mul[ka]);
  const float4 m1 = *reinterpret_cast<const float4*>(&s_mul[kb]);
  const unsigned long long mm[4] = {pack_f2(m0.x, m0.y), pack_f2(m0.z, m0.w), pack_f2(m1.x, m1.y), pack_f2(m1.z, m1.w)};
  const float* gbase = xg + (int64_t)g * HW * SPV_LD;
  const int64_t plane_elems = (int64_t)B * HW * 9 * C;
  for (int e = warp; e < ntask; e += 8) {
    int o[4];
    float wt[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o[c] = s_off[c][e];
      wt[c] = s_wt[c][e];
    }
    float4 xa[4], xb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      xa[c] = __ldg(reinterpret_cast<const float4*>(gbase + o[c] + ka));
      xb[c] = __ldg(reinterpret_cast<const float4*>(gbase + o[c] + kb));
    }
    const unsigned long long w0 = pack_f2(wt[0], wt[0]), w1 = pack_f2(wt[1], wt[1]), w2 = pack_f2(wt[2], wt[2]),
                             w3 = pack_f2(wt[3], wt[3]);
    unsigned long long acc[4];
    // ((a*w_nw + b*w_ne) + c*w_sw) + d*w_se per channel: the association of the scalar kernels in their contracted form
    acc[0] = fma2(pack_f2(xa[3].x, xa[3].y), w3, fma2(pack_f2(xa[2].x, xa[2].y), w2, fma2(pack_f2(xa[1].x, xa[1].y), w1, mul2(pack_f2(xa[0].x, xa[0].y), w0))));
    acc[1] = fma2(pack_f2(xa[3].z, xa[3].w), w3, fma2(pack_f2(xa[2].z, xa[2].w), w2, fma2(pack_f2(xa[1].z, xa[1].w), w1, mul2(pack_f2(xa[0].z, xa[0].w), w0))));
    acc[2] = fma2(pack_f2(xb[3].x, xb[3].y), w3, fma2(pack_f2(xb[2].x, xb[2].y), w2, fma2(pack_f2(xb[1].x, xb[1].y), w1, mul2(pack_f2(xb[0].x, xb[0].y), w0))));
    acc[3] = fma2(pack_f2(xb[3].z, xb[3].w), w3, fma2(pack_f2(xb[2].z, xb[2].w), w2, fma2(pack_f2(xb[1].z, xb[1].w), w1, mul2(pack_f2(xb[0].z, xb[0].w), w0))));
    float v[8];
    if (kinds != 0) {  // only the lanes of the group that holds the coordinate planes
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack_f2(acc[j], v[2 * j], v[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = encode_coord(v[j], (int)((kinds >> (2 * j)) & 3u));
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = pack_f2(v[2 * j], v[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[j] = mul2(acc[j], mm[j]);
      unpack_f2(acc[j], v[2 * j], v[2 * j + 1]);
    }
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_pair<kF16>(v[2 * j], v[2 * j + 1], h[j], l[j]);
    const int r = e / 9, t = e - r * 9;
    uint16_t* dst = out + (((int64_t)g * HW + p0 + r) * 9 + t) * C;
    *reinterpret_cast<uint2*>(dst + ka) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(dst + kb) = make_uint2(h[2], h[3]);
    *reinterpret_cast<uint2*>(dst + plane_elems + ka) = make_uint2(l[0], l[1]);
    *reinterpret_cast<uint2*>(dst + plane_elems + kb) = make_uint2(l[2], l[3]);
  }
  // ---- tail operand: columns C .. Ct-1 of every (pixel, tap) of this CTA and the zero padding, one element per thread
  // (inside the warp loop above the three tail lanes cost every task ~70 instructions)
  if (Cx > 0) {
    const int64_t plane2_elems = (int64_t)B * HW * kp2;
    for (int idx = threadIdx.x; idx < npx * kp2; idx += blockDim.x) {
      const int r = idx / kp2, col = idx - r * kp2;
      uint16_t hh = 0, ll = 0;
      if (col < 9 * Cx) {
        const int t = col / Cx, j = col - t * Cx;
        const int e = r * 9 + t;
        const uint32_t m = __ldg(mrow + C + j);
        const float* sp = gbase + C + j;
        float val = __ldg(sp + s_off[3][e]) * s_wt[3][e] +
                    (__ldg(sp + s_off[2][e]) * s_wt[2][e] + (__ldg(sp + s_off[1][e]) * s_wt[1][e] + __ldg(sp + s_off[0][e]) * s_wt[0][e]));
        if (m != 0xFFFFFFFFu && (m >> 31)) val = encode_coord(val, 1 + (int)(m & 3u));
        val *= s_mul[C + j];
        split16<kF16>(val, hh, ll);
      }
      const int64_t off = ((int64_t)g * HW + p0 + r) * kp2 + col;
      out2[off] = hh;
      out2[plane2_elems + off] = ll;
    }
  }
}

// One CTA per (sample, output row): the kh input rows of encoded, modulated coordinate values are computed ONCE into shared
// memory (kh * W * nc values; the first version evaluated tanh / cos / sin per output element, 49 times per value, and was
// issue-bound at 96 us for a 20 MB operand), then the im2col copy pairs them up and splits them, two columns per thread.
template <bool kF16>
__global__ void __launch_bounds__(256) coord_taps_pack_kernel(uint16_t* __restrict__ out2, const float* __restrict__ coords,
                                                             const float* __restrict__ in_mul, int64_t rows, int nc, int H,
                                                             int W, int kh, int kw, int My, int Mx, int mul_ld, int mul_off,
                                                             int kp2) {
  extern __shared__ float s_enc[];  // [kh][W][nc]
  const int b = blockIdx.x / My, y = blockIdx.x - b * My;
  for (int idx = threadIdx.x; idx < kh * W * nc; idx += blockDim.x) {
    const int j = idx % nc;
    const int q = idx / nc;
    const int ty = q / W, x = q - ty * W;
    float v = __ldg(coords + (((int64_t)b * nc + j) * H + (y + ty)) * W + x);
    v = encode_coord(v, j + 1);
    if (in_mul) v *= __ldg(in_mul + (int64_t)b * mul_ld + mul_off + j);
    s_enc[idx] = v;
  }
  __syncthreads();
  const int used = kh * kw * nc;
  const int half = kp2 >> 1;
  const int64_t plane = rows * kp2;
  uint16_t* orow = out2 + ((int64_t)b * My + y) * Mx * kp2;
  for (int idx = threadIdx.x; idx < Mx * half; idx += blockDim.x) {
    const int x = idx / half, col = (idx - x * half) * 2;
    float v[2] = {0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int c = col + u;
      if (c < used) {
        const int t = c / nc, j = c - t * nc;
        const int ty = t / kw, tx = t - ty * kw;
        v[u] = s_enc[(ty * W + x + tx) * nc + j];
      }
    }
    uint32_t h, l;
    split_pair<kF16>(v[0], v[1], h, l);
    *reinterpret_cast<uint32_t*>(orow + (int64_t)x * kp2 + col) = h;
    *reinterpret_cast<uint32_t*>(orow + plane + (int64_t)x * kp2 + col) = l;
  }
}

}  // namespace

extern "C" int spgan_sphere_concat_repack(float* xg, const float* x_nhwc, const float* coords, const uint32_t* chan_map, int B,
                                          int C, int H, int W, int cmap_ld, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && H >= 0 && W >= 0 && C == SPV_C, "spgan_sphere_concat_repack: needs C = %d features, got %d", SPV_C, C);
  const int nc = coords ? 3 : 0;
  SPGAN_CHECK_ARG(C + nc <= SPV_LD && cmap_ld >= C + nc, "spgan_sphere_concat_repack: channel count / chan_map stride out of range");
  if (B == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(xg && x_nhwc && chan_map, "spgan_sphere_concat_repack: null pointer");
  SPGAN_CHECK_ARG(B <= 65535, "spgan_sphere_concat_repack: batch %d > 65535", B);
  const int rp_px = 8;
  concat_repack_kernel<<<dim3((H * W + rp_px - 1) / rp_px, B), 256, 0, (cudaStream_t)stream>>>(xg, x_nhwc, coords, chan_map, C, nc, H * W, cmap_ld, rp_px);
  SPGAN_CHECK_LAUNCH("spgan_sphere_concat_repack");
  return 0;
}

extern "C" int64_t spgan_sphere_pack_seg_scratch(int B, int C, int H, int W) {
  return C == SPV_C ? (int64_t)B * H * W * SPV_LD : 0;
}

extern "C" int spgan_sphere_pack_seg(uint16_t* out, uint16_t* out2, const float* x_nhwc, const float* coords,
                                     const float* grid, const float* in_mul, const uint32_t* chan_map, int B, int C, int H,
                                     int W, int grid_group, int Cm, int cmap_ld, int kp2, int fmt, float* scratch, void* stream) {
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
  static const bool force_v1 = getenv("SPGAN_SPHERE_PACK_V1") != nullptr;  // diagnostics: A/B against the scalar producer
  if (!force_v1 && scratch != nullptr && C == SPV_C && Cm == SPV_C && C + nc <= SPV_LD && (((uintptr_t)scratch) & 15) == 0 &&
      (((uintptr_t)out) & 15) == 0 && (int64_t)H * W * SPV_LD < (1LL << 31)) {
    const int bpg = (H * W + SP_PX - 1) / SP_PX;
    SPGAN_CHECK_ARG((int64_t)B * bpg < (1LL << 31), "spgan_sphere_pack_seg: grid too large");
    const int rp_px = 8;
    concat_repack_kernel<<<dim3((H * W + rp_px - 1) / rp_px, B), 256, 0, st>>>(scratch, x_nhwc, coords, chan_map, C, nc, H * W, cmap_ld, rp_px);
    SPGAN_CHECK_LAUNCH("spgan_sphere_pack_seg (concat repack)");
    if (fmt)
      sphere_pack_v3_kernel<true><<<B * bpg, 256, 0, st>>>(out, out2, scratch, grid, in_mul, chan_map, B, C + nc, H, W, grid_group, cmap_ld, kp2, bpg);
    else
      sphere_pack_v3_kernel<false><<<B * bpg, 256, 0, st>>>(out, out2, scratch, grid, in_mul, chan_map, B, C + nc, H, W, grid_group, cmap_ld, kp2, bpg);
    SPGAN_CHECK_LAUNCH("spgan_sphere_pack_seg");
    return 0;
  }
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
  SPGAN_CHECK_ARG((int64_t)B * My < (1LL << 31) && (((uintptr_t)out2) & 3) == 0, "spgan_coord_taps_pack: grid too large / misaligned output");
  const size_t smem = (size_t)kh * W * nc * sizeof(float);
  SPGAN_CHECK_ARG(smem <= 48 * 1024, "spgan_coord_taps_pack: %d rows of %d pixels do not fit the staging buffer", kh, W);
  if (fmt)
    coord_taps_pack_kernel<true><<<B * My, 256, smem, st>>>(out2, coords, in_mul, rows, nc, H, W, kh, kw, My, Mx, mul_ld, mul_off, kp2);
  else
    coord_taps_pack_kernel<false><<<B * My, 256, smem, st>>>(out2, coords, in_mul, rows, nc, H, W, kh, kw, My, Mx, mul_ld, mul_off, kp2);
  SPGAN_CHECK_LAUNCH("spgan_coord_taps_pack");
  return 0;
}

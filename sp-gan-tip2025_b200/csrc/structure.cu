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
// sphere_pack_seg_kernel above is instruction-issue bound (~210 warp instructions per (pixel, tap): 32 scalar gathers, 32
// address computations, scalar blends and 2-byte splits), 19 % of the HBM write roofline.  Here one warp handles one (pixel,
// tap) with lane = one run of 8 consecutive K columns: when those 8 columns are 8 consecutive feature channels of one sample
// (every lane of every group except the few at the reference's flat-concat sample / coordinate boundaries, decided from
// chan_map once per CTA) the four corners are fetched with 2-3 aligned 128-bit loads each, blended with packed FFMA2, and
// stored as one 16-byte hi and one 16-byte lo vector (512 contiguous bytes per warp and plane).  A CTA owns SP2_PX pixels of
// ONE group, so the corner table of its 9 * SP2_PX (pixel, tap) pairs and the group's modulation row live in shared memory.
constexpr int SP2_PX = 16;
constexpr int SP2_TASKS = SP2_PX * 9;

__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 8 consecutive floats starting M elements (0..3) past the 16-byte aligned pointer p, as four packed pairs.
template <int M>
__device__ __forceinline__ void load8(const float* __restrict__ p, unsigned long long (&x)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  float w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, 0.f, 0.f, 0.f, 0.f};
  if (M > 0) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(p) + 2);
    w[8] = c.x;
    w[9] = c.y;
    w[10] = c.z;
    w[11] = c.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j] = pack_f2(w[2 * j + M], w[2 * j + 1 + M]);
}

// ((a*w_nw + b*w_ne) + c*w_sw) + d*w_se per channel, the association of the scalar kernels (their compiler-contracted form).
template <int M>
__device__ __forceinline__ void blend8(const float* __restrict__ base, const int (&o)[4], const float (&wt)[4],
                                       unsigned long long (&acc)[4]) {
  unsigned long long x0[4], x1[4], x2[4], x3[4];
  load8<M>(base + o[0], x0);
  load8<M>(base + o[1], x1);
  load8<M>(base + o[2], x2);
  load8<M>(base + o[3], x3);
  const unsigned long long w0 = pack_f2(wt[0], wt[0]), w1 = pack_f2(wt[1], wt[1]), w2 = pack_f2(wt[2], wt[2]),
                           w3 = pack_f2(wt[3], wt[3]);
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = fma2(x3[j], w3, fma2(x2[j], w2, fma2(x1[j], w1, mul2(x0[j], w0))));
}

template <bool kF16>
__device__ __forceinline__ void split_store8(uint16_t* __restrict__ hi_p, uint16_t* __restrict__ lo_p, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint16_t h0, l0, h1, l1;
    split16<kF16>(v[2 * j], h0, l0);
    split16<kF16>(v[2 * j + 1], h1, l1);
    h[j] = pack2x16(h0, h1);
    l[j] = pack2x16(l0, l1);
  }
  *reinterpret_cast<uint4*>(hi_p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_p) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool kF16>
__global__ void __launch_bounds__(256) sphere_pack_v2_kernel(uint16_t* __restrict__ out, uint16_t* __restrict__ out2,
                                                            const float* __restrict__ xh, const float* __restrict__ coords,
                                                            const float* __restrict__ grid, const float* __restrict__ in_mul,
                                                            const uint32_t* __restrict__ chan_map, int B, int nc, int H, int W,
                                                            int grid_group, int cmap_ld, int kp2, int blocks_per_group) {
  constexpr int C = 256;  // features per sample = main K columns per tap
  __shared__ int s_off[4][SP2_TASKS];
  __shared__ float s_wt[4][SP2_TASKS];
  __shared__ __align__(16) float s_mul[C + 32];
  __shared__ uint32_t s_map[C];
  const int Ct = C + nc;
  const int Cx = Ct - C;
  const int HW = H * W;
  const int g = blockIdx.x / blocks_per_group;
  const int p0 = (blockIdx.x - g * blocks_per_group) * SP2_PX;
  const int npx = min(SP2_PX, HW - p0);
  const int ntask = npx * 9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- corner table and modulation row
  for (int e = threadIdx.x; e < ntask; e += blockDim.x) {
    const int r = e / 9, t = e - r * 9;
    const int p = p0 + r;
    const int py = p / W, px = p - py * W;
    const TapCorners cn = tap_corners(grid, g / grid_group, H, W, py, px, t / 3, t - (t / 3) * 3);
    s_off[0][e] = cn.o_nw;
    s_off[1][e] = cn.o_ne;
    s_off[2][e] = cn.o_sw;
    s_off[3][e] = cn.o_se;
    s_wt[0][e] = cn.w_nw;
    s_wt[1][e] = cn.w_ne;
    s_wt[2][e] = cn.w_sw;
    s_wt[3][e] = cn.w_se;
  }
  for (int k = threadIdx.x; k < C + 32; k += blockDim.x) s_mul[k] = (k < Ct && in_mul) ? __ldg(in_mul + (int64_t)g * Ct + k) : (k < Ct ? 1.f : 0.f);
  // ---- this lane's 8 columns: one run of consecutive feature channels of one sample?
  const uint32_t* mrow = chan_map + (int64_t)g * cmap_ld;
  const int k0 = 8 * lane;
  uint32_t mw[8];
  {
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(mrow + k0));
    const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(mrow + k0) + 1);
    mw[0] = q0.x; mw[1] = q0.y; mw[2] = q0.z; mw[3] = q0.w;
    mw[4] = q1.x; mw[5] = q1.y; mw[6] = q1.z; mw[7] = q1.w;
  }
  bool run = (mw[0] >> 31) == 0;  // a feature (an all-ones padding entry has bit 31 set)
#pragma unroll
  for (int j = 1; j < 8; ++j) run = run && mw[j] == mw[0] + (uint32_t)j;  // same sample (bits 15..30), channel + j (no carry below)
  const int cs0 = (int)(mw[0] & 0x7FFFu);
  const int mis = cs0 & 3;
  run = run && (cs0 - mis + (mis ? 12 : 8) <= C);
  const float* fbase = xh + ((int64_t)((mw[0] >> 15) & 0xFFFFu) * HW) * C + (cs0 - mis);
  const unsigned slowmask = __ballot_sync(0xffffffffu, !run);
  if (warp == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_map[k0 + j] = mw[j];
  }
  __syncthreads();
  const int64_t plane_elems = (int64_t)B * HW * 9 * C;
  for (int e = warp; e < ntask; e += 8) {
    int o[4];
    float wt[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o[c] = s_off[c][e];
      wt[c] = s_wt[c][e];
    }
    const int r = e / 9, t = e - r * 9;
    uint16_t* dst = out + (((int64_t)g * HW + p0 + r) * 9 + t) * C;
    if (run) {
      int oc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) oc[c] = o[c] * C;
      unsigned long long acc[4];
      switch (mis) {
        case 0: blend8<0>(fbase, oc, wt, acc); break;
        case 1: blend8<1>(fbase, oc, wt, acc); break;
        case 2: blend8<2>(fbase, oc, wt, acc); break;
        default: blend8<3>(fbase, oc, wt, acc); break;
      }
      const float4 m0 = *reinterpret_cast<const float4*>(&s_mul[k0]);
      const float4 m1 = *reinterpret_cast<const float4*>(&s_mul[k0 + 4]);
      acc[0] = mul2(acc[0], pack_f2(m0.x, m0.y));
      acc[1] = mul2(acc[1], pack_f2(m0.z, m0.w));
      acc[2] = mul2(acc[2], pack_f2(m1.x, m1.y));
      acc[3] = mul2(acc[3], pack_f2(m1.z, m1.w));
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) unpack_f2(acc[j], v[2 * j], v[2 * j + 1]);
      split_store8<kF16>(dst + k0, dst + k0 + plane_elems, v);
    }
    // boundary runs (sample change of the flat concat, coordinate planes, padding): one COLUMN per lane, 8 lanes per run,
    // so that the one such run nearly every group has costs ~30 instructions per task instead of a serial 8-column loop
    for (unsigned sm = slowmask; sm; sm &= sm - 1) {
      const int k = 8 * (__ffs(sm) - 1) + (lane & 7);
      if (lane < 8) {
        const uint32_t m = s_map[k];
        float val = 0.f;
        if (m != 0xFFFFFFFFu) {
          const uint32_t bs = (m >> 15) & 0xFFFFu, cs = m & 0x7FFFu;
          const bool is_coord = (m >> 31) != 0;
          const float* sp = is_coord ? coords + ((int64_t)bs * nc + cs) * HW : xh + (int64_t)bs * HW * C + cs;
          const int st = is_coord ? 1 : C;
          val = __ldg(sp + o[0] * st) * wt[0] + __ldg(sp + o[1] * st) * wt[1] + __ldg(sp + o[2] * st) * wt[2] +
                __ldg(sp + o[3] * st) * wt[3];
          if (is_coord) {
            if (cs == 0) val = tanhf(val);
            else if (cs == 1) val = cosf(val * 3.14159274101257324f);
            else if (cs == 2) val = sinf(val * 3.14159274101257324f);
          }
          val *= s_mul[k];
        }
        uint16_t h, l;
        split16<kF16>(val, h, l);
        dst[k] = h;
        dst[k + plane_elems] = l;
      }
    }
  }
  // ---- tail columns (channels C .. Ct-1 of the group, all taps) and their zero padding
  if (Cx > 0) {
    const int64_t plane2_elems = (int64_t)B * HW * kp2;
    for (int idx = threadIdx.x; idx < npx * kp2; idx += blockDim.x) {
      const int r = idx / kp2, col = idx - r * kp2;
      uint16_t h = 0, l = 0;
      if (col < 9 * Cx) {
        const int t = col / Cx, j = col - t * Cx;
        const int e = r * 9 + t;
        const uint32_t m = __ldg(mrow + C + j);
        float val = 0.f;
        if (m != 0xFFFFFFFFu) {
          const uint32_t bs = (m >> 15) & 0xFFFFu, cs = m & 0x7FFFu;
          const bool is_coord = (m >> 31) != 0;
          const float* sp = is_coord ? coords + ((int64_t)bs * nc + cs) * HW : xh + (int64_t)bs * HW * C + cs;
          const int st = is_coord ? 1 : C;
          val = __ldg(sp + s_off[0][e] * st) * s_wt[0][e] + __ldg(sp + s_off[1][e] * st) * s_wt[1][e] +
                __ldg(sp + s_off[2][e] * st) * s_wt[2][e] + __ldg(sp + s_off[3][e] * st) * s_wt[3][e];
          if (is_coord) {
            if (cs == 0) val = tanhf(val);
            else if (cs == 1) val = cosf(val * 3.14159274101257324f);
            else if (cs == 2) val = sinf(val * 3.14159274101257324f);
          }
          val *= s_mul[C + j];
        }
        split16<kF16>(val, h, l);
      }
      const int64_t off = ((int64_t)g * HW + p0 + r) * kp2 + col;
      out2[off] = h;
      out2[plane2_elems + off] = l;
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
  static const bool force_v1 = getenv("SPGAN_SPHERE_PACK_V1") != nullptr;  // diagnostics: A/B against the scalar producer
  if (!force_v1 && C == 256 && Cm == 256 && cmap_ld % 4 == 0 && (((uintptr_t)chan_map) & 15) == 0 && (((uintptr_t)x_nhwc) & 15) == 0 &&
      (((uintptr_t)out) & 15) == 0) {
    const int bpg = (H * W + SP2_PX - 1) / SP2_PX;
    SPGAN_CHECK_ARG((int64_t)B * bpg < (1LL << 31), "spgan_sphere_pack_seg: grid too large");
    if (fmt)
      sphere_pack_v2_kernel<true><<<B * bpg, 256, 0, st>>>(out, out2, x_nhwc, coords, grid, in_mul, chan_map, B, nc, H, W, grid_group, cmap_ld, kp2, bpg);
    else
      sphere_pack_v2_kernel<false><<<B * bpg, 256, 0, st>>>(out, out2, x_nhwc, coords, grid, in_mul, chan_map, B, nc, H, W, grid_group, cmap_ld, kp2, bpg);
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
  const int nblk = grid_for(rows * kp2, 256, 4);
  if (fmt)
    coord_taps_pack_kernel<true><<<nblk, 256, 0, st>>>(out2, coords, in_mul, rows, nc, H, W, kw, kh * kw, My, Mx, mul_ld, mul_off, kp2);
  else
    coord_taps_pack_kernel<false><<<nblk, 256, 0, st>>>(out2, coords, in_mul, rows, nc, H, W, kw, kh * kw, My, Mx, mul_ld, mul_off, kp2);
  SPGAN_CHECK_LAUNCH("spgan_coord_taps_pack");
  return 0;
}

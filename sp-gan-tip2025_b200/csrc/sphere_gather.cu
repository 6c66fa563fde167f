// L1: SphereNet-style distortion-aware gather — bilinear sampling of a (B, C, H, W) feature at a 3x3 tap grid
// (B|1, 3H, 3W, 2), border padding, align_corners=True — its index-only variant, and the reference's surrogate
// backward (3x3 block mean * 0.1).  Index arithmetic is ATen's fp32 sequence, evaluated with explicit
// round-to-nearest intrinsics so that no FMA contraction can change a floor():
//     ix = ((gx + 1) / 2) * (W - 1);  ix = min(W - 1, max(ix, 0));  x0 = floor(ix)
// HBM roofline: writes 9x the input (4*B*C*9*H*W bytes) + reads input once (the 36 corner reads per input pixel
// hit L1/L2) + the grid.
#include "common.cuh"
#include "stream_stage.cuh"

namespace {

struct Corner {
  int x0, y0, x1, y1;
  float nw, ne, sw, se;
};

__device__ __forceinline__ float unnormalize_clip(float g, int size) {
  // (g + 1) / 2: halving is exact, so the multiply is bit-identical to ATen's division and ~10 instructions shorter
  float v = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
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

// Streamed gather (the default): persistent CTAs; a work item = CC (8 or 4) consecutive channel planes of one sample x a
// slice of the output positions.  The planes are one contiguous run of CC*H*W floats, fetched by ONE bulk-async copy
// (stream_stage.cuh) while the previous item is sampled; the CTA then transposes them inside shared memory to
// [pixel][channel] (conflict-free 128-bit stores), so that a corner of a position is ONE 128-bit shared-memory read per four
// channels at an immediate offset, and the blend of two channels is one packed FFMA2.  A thread owns output positions: it
// evaluates the corner table of a position once and reuses it for the CC channels.  ~10 instructions per output instead of
// ~17, and no scattered L1 requests (the L1-gather kernel above: 4 LDG x 2-4 wavefronts per 32 outputs, 0.26 of the HBM
// peak).  Stores are 128-byte coalesced rows of the 3H x 3W output planes.  Same corner arithmetic and blend order as the
// kernel above: bit-identical outputs (tests/test_gpu_ops.py).
constexpr int GS_THREADS = 512;

struct GatherStream {
  int B, C, H, W, grid_batch, encode;
  int chunks, psplit, pslice, raw_floats;
  int64_t out_bstride, out_coff, nitems;
  uintptr_t limit;
};

__device__ __forceinline__ unsigned long long gs_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

template <int CCQ, bool ENC>  // channel quads per item: 2 (8 channels) or 1 (4 channels); ENC: coordinate encoding (C == 3)
__global__ void __launch_bounds__(GS_THREADS, 2) sphere_gather_stream_kernel(float* __restrict__ out,
                                                                            const float* __restrict__ z,
                                                                            const float* __restrict__ grid, GatherStream q) {
  constexpr int CC = 4 * CCQ;
  constexpr int TS = CCQ == 1 ? 4 : 12;  // floats per pixel row of the transposed tile: an odd number of 16-byte units
  extern __shared__ __align__(128) float gs_smem[];
  __shared__ uint64_t bar;
  float* raw = gs_smem;
  float* tile = gs_smem + q.raw_floats;
  const int tid = threadIdx.x;
  if (tid == 0) {
    stream_stage::bar_init(stream_stage::smem_addr(&bar), 1);
    stream_stage::fence_bar_init();
  }
  __syncthreads();
  const int HW = q.H * q.W;
  const int opix = 9 * HW;
  const int per_sample = q.chunks * q.psplit;
  auto issue = [&](int64_t item) {
    const int b = (int)(item / per_sample);
    const int c0 = ((int)(item - (int64_t)b * per_sample) / q.psplit) * CC;
    const int nc = min(CC, q.C - c0);
    stream_stage::issue_chunk(raw, z + ((int64_t)b * q.C + c0) * HW, nc * HW, q.limit, &bar);
  };
  int64_t item = blockIdx.x;
  if (tid == 0 && item < q.nitems) issue(item);
  const uint32_t tile_addr = stream_stage::smem_addr(tile);
  for (uint32_t k = 0; item < q.nitems; item += gridDim.x, ++k) {
    const int b = (int)(item / per_sample);
    const int rem = (int)(item - (int64_t)b * per_sample);
    const int c0 = (rem / q.psplit) * CC;
    const int slice = rem - (rem / q.psplit) * q.psplit;
    const int nc = min(CC, q.C - c0);
    const float* src = z + ((int64_t)b * q.C + c0) * HW;
    const float* st = raw + stream_stage::chunk_shift(src);
    stream_stage::bar_wait(stream_stage::smem_addr(&bar), k & 1);
    // transpose [channel][pixel] -> [pixel][channel]; channels beyond the tensor read as zero
    for (int e = tid; e < HW * CCQ; e += GS_THREADS) {
      const int qd = e / HW, pix = e - qd * HW;
      float4 v;
      v.x = 4 * qd + 0 < nc ? st[(4 * qd + 0) * HW + pix] : 0.f;
      v.y = 4 * qd + 1 < nc ? st[(4 * qd + 1) * HW + pix] : 0.f;
      v.z = 4 * qd + 2 < nc ? st[(4 * qd + 2) * HW + pix] : 0.f;
      v.w = 4 * qd + 3 < nc ? st[(4 * qd + 3) * HW + pix] : 0.f;
      *reinterpret_cast<float4*>(tile + pix * TS + 4 * qd) = v;
    }
    __syncthreads();  // tile complete, raw planes consumed
    if (tid == 0 && item + gridDim.x < q.nitems) issue(item + gridDim.x);
    const float2* gp = reinterpret_cast<const float2*>(grid) + (int64_t)(q.grid_batch == 1 ? 0 : b) * opix;
    float* ob = out + ((int64_t)b * q.out_bstride + q.out_coff + c0) * opix;
    const int p_end = min(opix, (slice + 1) * q.pslice);
    // the tap grid comes from L2 (~1 us away): keep the loads of the next two positions in flight while one is blended
    const int p_first = slice * q.pslice + tid;
    float2 g_1 = p_first < p_end ? __ldg(gp + p_first) : make_float2(0.f, 0.f);
    float2 g_2 = p_first + GS_THREADS < p_end ? __ldg(gp + p_first + GS_THREADS) : g_1;
    for (int p = p_first; p < p_end; p += GS_THREADS) {
      const float2 g = g_1;
      g_1 = g_2;
      if (p + 2 * GS_THREADS < p_end) g_2 = __ldg(gp + p + 2 * GS_THREADS);
      const Corner cn = corners(g.x, g.y, q.H, q.W);
      const uint32_t a_n = tile_addr + (uint32_t)(cn.y0 * q.W) * (TS * 4u), a_s = tile_addr + (uint32_t)(cn.y1 * q.W) * (TS * 4u);
      const uint32_t a_nw = a_n + (uint32_t)cn.x0 * (TS * 4u), a_ne = a_n + (uint32_t)cn.x1 * (TS * 4u);
      const uint32_t a_sw = a_s + (uint32_t)cn.x0 * (TS * 4u), a_se = a_s + (uint32_t)cn.x1 * (TS * 4u);
      const unsigned long long w_nw = gs_pack(cn.nw, cn.nw), w_ne = gs_pack(cn.ne, cn.ne), w_sw = gs_pack(cn.sw, cn.sw),
                               w_se = gs_pack(cn.se, cn.se);
      float* o = ob + p;
#pragma unroll
      for (int qd = 0; qd < CCQ; ++qd) {
        unsigned long long nw0, nw1, ne0, ne1, sw0, sw1, se0, se1;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nw0), "=l"(nw1) : "r"(a_nw + 16u * qd));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(ne0), "=l"(ne1) : "r"(a_ne + 16u * qd));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(sw0), "=l"(sw1) : "r"(a_sw + 16u * qd));
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(se0), "=l"(se1) : "r"(a_se + 16u * qd));
        unsigned long long r0, r1;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r0) : "l"(nw0), "l"(w_nw));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r1) : "l"(nw1), "l"(w_nw));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r0) : "l"(ne0), "l"(w_ne));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r1) : "l"(ne1), "l"(w_ne));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r0) : "l"(sw0), "l"(w_sw));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r1) : "l"(sw1), "l"(w_sw));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r0) : "l"(se0), "l"(w_se));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r1) : "l"(se1), "l"(w_se));
        float v[4];
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(r0));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(r1));
        if (ENC && qd == 0) {  // channels 0, 1, 2 of the coordinate tensor (C == 3: c0 == 0)
          v[0] = tanhf(v[0]);
          v[1] = cosf(v[1] * 3.14159274101257324f);
          v[2] = sinf(v[2] * 3.14159274101257324f);
        }
        if (nc == CC) {  // running pointer: one 64-bit add per store instead of a multiply-add chain per channel
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __stcs(o, v[j]);
            o += opix;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (4 * qd + j < nc) __stcs(o, v[j]);
            o += opix;
          }
        }
      }
    }
    __syncthreads();  // every read of the tile is done before the next item overwrites it
  }
}

constexpr int64_t GS_SMEM_MAX = 220 * 1024;

// Host-side plan of a streamed gather: channels per item, position slices, shared-memory bytes, grid.  Pure arithmetic,
// exported as spgan_sphere_gather_plan so that the CPU tests can sweep it for its invariants.
bool plan_gather_stream(int B, int C, int H, int W, int encode, GatherStream& q, int& cc_out, size_t& smem, unsigned& grid_n,
                        int& resident_out) {
  const int64_t HW = (int64_t)H * W;
  if (9 * HW >= (1LL << 26)) return false;
  const int nsm = SPGAN_NUM_SMS;
  auto bytes = [&](int cc) {
    const int64_t raw = (cc * HW + 8 + 31) / 32 * 32;
    return (raw + HW * (cc == 8 ? 12 : 4)) * 4;
  };
  int cc = (C >= 8 && !encode && bytes(8) <= GS_SMEM_MAX) ? 8 : 4;
  if (const char* e = getenv("SPGAN_GS_CC")) cc = atoi(e) == 4 ? 4 : cc;  // diagnostics
  if (bytes(cc) > GS_SMEM_MAX) return false;  // planes too large to stage: L1-gather kernel
  const int resident = (2 * (bytes(cc) + 1024) <= GS_SMEM_MAX + 4096) ? 2 : 1;
  const int64_t ncta = (int64_t)nsm * resident;
  const int chunks = (C + cc - 1) / cc;
  // position slices per (sample, channel group): the smallest split that balances the items over the persistent CTAs
  int psplit = 1;
  double best = -1.0;
  for (int ps = 1; ps <= 4; ++ps) {
    const int64_t items = (int64_t)B * chunks * ps;
    const double eff = (double)items / (double)(((items + ncta - 1) / ncta) * ncta);
    const double score = eff * (1.0 - 0.02 * (ps - 1));  // each slice re-stages the planes
    if (score > best + 1e-9) {
      best = score;
      psplit = ps;
    }
  }
  if (const char* e = getenv("SPGAN_GS_PSPLIT")) psplit = atoi(e) >= 1 ? atoi(e) : psplit;  // diagnostics
  q.B = B; q.C = C; q.H = H; q.W = W; q.grid_batch = 1; q.encode = encode;
  q.chunks = chunks;
  q.psplit = psplit;
  q.pslice = (int)((9 * HW + psplit - 1) / psplit);
  q.raw_floats = (int)((cc * HW + 8 + 31) / 32 * 32);
  q.out_bstride = C; q.out_coff = 0;
  q.nitems = (int64_t)B * chunks * psplit;
  q.limit = 0;
  cc_out = cc;
  smem = (size_t)bytes(cc);
  grid_n = (unsigned)(q.nitems < ncta ? q.nitems : ncta);
  resident_out = resident;
  return true;
}

bool launch_gather_stream(float* out, const float* z, const float* grid, int B, int C, int H, int W, int grid_batch,
                          int64_t out_bstride, int64_t out_coff, int encode, cudaStream_t st) {
  if ((((uintptr_t)z) & 15) != 0) return false;
  GatherStream q;
  int cc = 0, resident = 0;
  size_t smem = 0;
  unsigned grid_n = 0;
  if (!plan_gather_stream(B, C, H, W, encode, q, cc, smem, grid_n, resident)) return false;
  q.grid_batch = grid_batch;
  q.out_bstride = out_bstride; q.out_coff = out_coff;
  q.limit = ((uintptr_t)(z + (int64_t)B * C * H * W)) & ~(uintptr_t)15;
  static bool done_a[64] = {false}, done_b[64] = {false}, done_c[64] = {false};
  if (!spgan_allow_smem(sphere_gather_stream_kernel<2, false>, (int)GS_SMEM_MAX, done_a) ||
      !spgan_allow_smem(sphere_gather_stream_kernel<1, false>, (int)GS_SMEM_MAX, done_b) ||
      !spgan_allow_smem(sphere_gather_stream_kernel<1, true>, (int)GS_SMEM_MAX, done_c))
    return false;
  if (encode) sphere_gather_stream_kernel<1, true><<<grid_n, GS_THREADS, smem, st>>>(out, z, grid, q);
  else if (cc == 8) sphere_gather_stream_kernel<2, false><<<grid_n, GS_THREADS, smem, st>>>(out, z, grid, q);
  else sphere_gather_stream_kernel<1, false><<<grid_n, GS_THREADS, smem, st>>>(out, z, grid, q);
  return true;
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

extern "C" int spgan_sphere_gather(float* out, const float* z, const float* grid, int B, int C, int H, int W,
                                   int grid_batch, int64_t out_bstride, int64_t out_coff, int encode, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "spgan_sphere_gather: negative size");
  if (B == 0 || C == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && z && grid, "spgan_sphere_gather: null pointer");
  SPGAN_CHECK_ARG(grid_batch == 1 || grid_batch == B, "spgan_sphere_gather: grid batch %d must be 1 or %d", grid_batch, B);
  SPGAN_CHECK_ARG(!encode || C == 3, "spgan_sphere_gather: coordinate encoding expects 3 channels, got %d", C);
  SPGAN_CHECK_ARG((((uintptr_t)grid) & 7) == 0, "spgan_sphere_gather: grid must be 8-byte aligned");
  const int64_t total = (int64_t)B * 9 * H * W;
  if (!spgan_legacy_hbm() &&
      launch_gather_stream(out, z, grid, B, C, H, W, grid_batch, out_bstride, out_coff, encode, (cudaStream_t)stream)) {
    SPGAN_CHECK_LAUNCH("spgan_sphere_gather");
    return 0;
  }
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

extern "C" int spgan_sphere_gather_plan(int B, int C, int H, int W, int encode, int32_t* plan) {
  SPGAN_CHECK_ARG(plan != nullptr, "spgan_sphere_gather_plan: null plan");
  SPGAN_CHECK_ARG(B >= 1 && C >= 1 && H >= 1 && W >= 1, "spgan_sphere_gather_plan: sizes must be positive");
  for (int i = 0; i < 12; ++i) plan[i] = 0;
  GatherStream q;
  int cc = 0, resident = 0;
  size_t smem = 0;
  unsigned grid_n = 0;
  if (!plan_gather_stream(B, C, H, W, encode, q, cc, smem, grid_n, resident)) return 0;
  plan[0] = 1;
  plan[1] = cc; plan[2] = q.chunks; plan[3] = q.psplit; plan[4] = q.pslice; plan[5] = q.raw_floats; plan[6] = (int32_t)smem;
  plan[7] = resident; plan[8] = (int32_t)grid_n;
  plan[9] = (int32_t)(q.nitems < 2147483647LL ? q.nitems : 2147483647LL);
  plan[10] = (int32_t)GS_SMEM_MAX;
  plan[11] = GS_THREADS;
  return 0;
}

// K2/K3: upfirdn2d over (planes, H, W) fp32 images — pad, zero-stuff by `up`, FIR with the flipped kernel,
// keep every `down`-th sample.  HBM-bound stencil: algorithmic traffic 4*(in_h*in_w + out_h*out_w) bytes per plane.
//   * streamed kernels (the default: fir_stream_kernel<K, UP, DOWN>, K <= 4, up or down by two or neither): persistent CTAs,
//     whole planes / row bands staged by bulk-async copies into a two-stage ring (stream_stage.cuh), one output column per
//     thread.  The Blur of the generator (3x3, pad 0) and of the discriminator (4x4, pad 2 / 1), Upsample, Downsample and
//     their gradients.
//   Fallbacks (input not 16-byte aligned, rows too long for a stage, other geometries; SPGAN_LEGACY_HBM_KERNELS=1):
//   * band kernel (up = down = 1, K x K with K <= 4, rows that fit the staging buffer): a CTA filters R full-width output rows.
//   * tiled kernel (same filters, very wide images): one 64x64 output tile per CTA.
//   * polyphase kernel (up, down in {1, 2}, kernels up to 4x4).
//   * generic kernel: any up/down/pad/kernel up to 16x16 (reads through L1/L2).
#include "common.cuh"
#include "stream_stage.cuh"

namespace {

struct UfdParams {
  int in_h, in_w, out_h, out_w;
  int kh, kw;
  int up_x, up_y, down_x, down_y;
  int pad_x0, pad_y0;
};

__global__ void __launch_bounds__(256) upfirdn2d_generic(float* __restrict__ out, const float* __restrict__ x,
                                                        const float* __restrict__ kernel, int64_t planes, UfdParams p) {
  __shared__ float kf[256];  // flipped kernel
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    int ky = i / p.kw, kx = i - ky * p.kw;
    kf[i] = kernel[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  const int64_t per_plane = (int64_t)p.out_h * p.out_w;
  const int64_t total = planes * per_plane;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = idx / per_plane;
    const int rem = (int)(idx - plane * per_plane);
    const int oy = rem / p.out_w, ox = rem - oy * p.out_w;
    const float* xp = x + plane * (int64_t)p.in_h * p.in_w;
    float acc = 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      const int Y = oy * p.down_y + ky - p.pad_y0;
      if (Y < 0 || Y % p.up_y != 0) continue;
      const int iy = Y / p.up_y;
      if (iy >= p.in_h) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        const int X = ox * p.down_x + kx - p.pad_x0;
        if (X < 0 || X % p.up_x != 0) continue;
        const int ix = X / p.up_x;
        if (ix >= p.in_w) continue;
        acc += kf[ky * p.kw + kx] * __ldg(xp + (int64_t)iy * p.in_w + ix);
      }
    }
    out[idx] = acc;
  }
}

// Polyphase FIR for up, down in {1, 2} and kernels up to 4x4 (Upsample / Downsample of models/ops.py:32-79 and the gradients of
// each other): the generic kernel below spends ~25 emulated-division instructions per tap on `% up` / `/ up` and decodes a
// 64-bit flat index per output (0.03-0.08 of the HBM roofline in the configs[4] sweep).  Here up / down are template
// parameters (parity tests and shifts), a CTA owns a 32 x 32 output tile of one plane, a thread four outputs of one column;
// neighbouring outputs share their inputs through L1.  Measured (tools/microbench.py sweep): 0.2-0.3 of the HBM roofline for
// down = 2 and 0.09 for up = 2 (2-3x the generic kernel) — still instruction-bound on per-tap predicates (a predicate-first,
// branch-free variant was 3x slower: 64 predicated loads per thread).  Off the hot path: spgan.yaml reaches up = 2 only through
// ToRGB's 3-channel skip upsample (models/spgan_ops.py:36-65); a quad-per-thread phase decomposition is the next step.
template <int UP, int DOWN>
__global__ void __launch_bounds__(256) upfirdn2d_poly(float* __restrict__ out, const float* __restrict__ x,
                                                     const float* __restrict__ kernel, int64_t planes, int tiles_x, UfdParams p) {
  __shared__ float kf[16];  // flipped kernel, row stride 4
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    kf[threadIdx.x] = (ky < p.kh && kx < p.kw) ? __ldg(kernel + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)) : 0.f;
  }
  __syncthreads();
  const int tile_y = blockIdx.x / tiles_x, tile_x = blockIdx.x - tile_y * tiles_x;
  const int ox = tile_x * 32 + (threadIdx.x & 31);
  const int oy0 = tile_y * 32 + (threadIdx.x >> 5);
  if (ox >= p.out_w) return;
  for (int64_t plane = blockIdx.y; plane < planes; plane += gridDim.y) {
    const float* __restrict__ xp = x + plane * (int64_t)p.in_h * p.in_w;
    float* __restrict__ op = out + plane * (int64_t)p.out_h * p.out_w;
    // this column's taps: input column and validity per kx (the same for the four outputs)
    int ixs[4];
    bool okx[4];
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int X = ox * DOWN + kx - p.pad_x0;
      okx[kx] = kx < p.kw && X >= 0 && (UP == 1 || (X & 1) == 0) && (X / UP) < p.in_w;
      ixs[kx] = X / UP;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int oy = oy0 + 8 * i;
      if (oy >= p.out_h) break;
      float acc = 0.f;
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const int Y = oy * DOWN + ky - p.pad_y0;
        if (ky >= p.kh || Y < 0 || (UP == 2 && (Y & 1))) continue;  // warp-uniform: a warp is one output row
        const int iy = Y / UP;
        if (iy >= p.in_h) continue;
        const float* row = xp + (int64_t)iy * p.in_w;
#pragma unroll
        for (int kx = 0; kx < 4; ++kx)
          if (okx[kx]) acc += kf[ky * 4 + kx] * __ldg(row + ixs[kx]);
      }
      op[(int64_t)oy * p.out_w + ox] = acc;
    }
  }
}

// Tiled FIR kernels: 64x64 output tile per CTA (256 threads as 64 columns x 4 row groups, 16 rows per thread).
// A CTA's lifetime is one global-load latency + one store; the tile size is what keeps enough bytes in flight per SM
// (8 resident CTAs x ~17 KB) for the kernel to sit on the HBM roofline rather than on the load latency.
constexpr int TILE = 64;
constexpr int TROWS = 4;  // thread rows

// Tiled FIR (up = down = 1, K x K, zero padding): 64x64 output tile per CTA.  All of a thread's ~17 input loads are issued
// before the first shared-memory store (>= 64 KB in flight per SM: the kernel sits on HBM bandwidth, not on load latency),
// and each thread filters a vertical strip of 16 rows with a sliding window (K shared-memory reads per output, not K*K).
template <int K>
__global__ void __launch_bounds__(256) upfirdn2d_tiled(float* __restrict__ out, const float* __restrict__ x,
                                                      const float* __restrict__ kernel, int tiles_x, int tiles_y,
                                                      UfdParams p) {
  constexpr int IN = TILE + K - 1;
  constexpr int NLOAD = (IN * IN + 255) / 256;
  __shared__ float tile[IN][IN + 1];
  __shared__ float kf[K * K];
  if (threadIdx.x < K * K) {
    int ky = threadIdx.x / K, kx = threadIdx.x - ky * K;
    kf[threadIdx.x] = kernel[(K - 1 - ky) * K + (K - 1 - kx)];
  }
  const int tiles_per_plane = tiles_x * tiles_y;
  const int64_t plane = blockIdx.x / tiles_per_plane;
  const int t = blockIdx.x - (int)(plane * tiles_per_plane);
  const int oy0 = (t / tiles_x) * TILE, ox0 = (t % tiles_x) * TILE;
  const float* xp = x + plane * (int64_t)p.in_h * p.in_w;
  const int iy0 = oy0 - p.pad_y0, ix0 = ox0 - p.pad_x0;
  const int rows_needed = min(IN, p.out_h - oy0 + K - 1);
  const int cols_needed = min(IN, p.out_w - ox0 + K - 1);
  float v[NLOAD];
#pragma unroll
  for (int u = 0; u < NLOAD; ++u) {
    const int e = threadIdx.x + u * 256;
    const int r = e / IN, c = e - r * IN;
    const int iy = iy0 + r, ix = ix0 + c;
    const bool ok = r < rows_needed && c < cols_needed && iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
    v[u] = ok ? __ldcs(xp + (int64_t)iy * p.in_w + ix) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < NLOAD; ++u) {
    const int e = threadIdx.x + u * 256;
    const int r = e / IN, c = e - r * IN;
    if (r < IN) tile[r][c] = v[u];
  }
  __syncthreads();
  float w[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) w[i] = kf[i];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 columns x 4 strips of 16 rows
  float* op = out + plane * (int64_t)p.out_h * p.out_w;
  const int ox = ox0 + tx;
  if (ox >= p.out_w) return;
  constexpr int STRIP = TILE / TROWS;
  const int ly0 = ty * STRIP;
  float win[K][K];
#pragma unroll
  for (int ky = 0; ky < K - 1; ++ky)
#pragma unroll
    for (int kx = 0; kx < K; ++kx) win[ky + 1][kx] = tile[ly0 + ky][tx + kx];
#pragma unroll
  for (int r = 0; r < STRIP; ++r) {
    const int oy = oy0 + ly0 + r;
    if (oy >= p.out_h) break;
#pragma unroll
    for (int ky = 0; ky < K - 1; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) win[ky][kx] = win[ky + 1][kx];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) win[K - 1][kx] = tile[ly0 + r + K - 1][tx + kx];
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) acc += w[ky * K + kx] * win[ky][kx];
    __stcs(op + (int64_t)oy * p.out_w + ox, acc);
  }
}

// Band FIR (up = down = 1, K x K, zero padding, rows narrow enough to stage): a CTA owns R output rows of one plane at
// FULL width.  The R + K - 1 input rows are staged in shared memory with their zero borders (so the stencil needs no
// bounds checks), all loads of a batch in flight before the first shared-memory store; each thread then filters vertical
// quads of outputs (K * (K + 3) shared-memory reads and 4 K^2 FMAs per 4 outputs), lanes along the row (coalesced stores).
// Unlike the 64 x 64 tiles above, every lane does useful work whatever the image size (tiles waste 35 % of the lanes at
// 103 x 103).
constexpr int FB_THREADS = 256;
constexpr int FB_LD = 8;  // loads per thread and batch

template <int K>
__global__ void __launch_bounds__(FB_THREADS) upfirdn2d_band(float* __restrict__ out, const float* __restrict__ x,
                                                            const float* __restrict__ kernel, UfdParams p, int R,
                                                            int bands, int SW, FastDiv dinw, FastDiv doutw) {
  extern __shared__ float stage[];  // [(R + K - 1)][SW], column pad_x0 + c holds input column c
  __shared__ float kf[K * K];
  if (threadIdx.x < K * K) {
    int ky = threadIdx.x / K, kx = threadIdx.x - ky * K;
    kf[threadIdx.x] = kernel[(K - 1 - ky) * K + (K - 1 - kx)];
  }
  const int64_t plane = blockIdx.x / bands;
  const int band = blockIdx.x - (int)(plane * bands);
  const int oy0 = band * R;
  const int rows_out = min(R, p.out_h - oy0);
  const int rows_in = rows_out + K - 1;
  const int iy0 = oy0 - p.pad_y0;
  const float* xp = x + plane * (int64_t)p.in_h * p.in_w;
  // zero borders: columns [0, pad_x0) and [pad_x0 + in_w, SW) of every staged row
  const int border = SW - p.in_w;
  for (int e = threadIdx.x; e < rows_in * border; e += FB_THREADS) {
    const int r = e / border, c = e - r * border;
    stage[r * SW + (c < p.pad_x0 ? c : p.in_w + c)] = 0.f;
  }
  const int n = rows_in * p.in_w;
  for (int e0 = 0; e0 < n; e0 += FB_LD * FB_THREADS) {
    float v[FB_LD];
    int dst[FB_LD];
#pragma unroll
    for (int u = 0; u < FB_LD; ++u) {
      const int e = e0 + u * FB_THREADS + threadIdx.x;
      dst[u] = -1;
      v[u] = 0.f;
      if (e < n) {
        const uint32_t r = fdiv((uint32_t)e, dinw);
        const int c = e - (int)r * p.in_w;
        const int iy = iy0 + (int)r;
        dst[u] = (int)r * SW + p.pad_x0 + c;
        if (iy >= 0 && iy < p.in_h) v[u] = __ldcs(xp + (int64_t)iy * p.in_w + c);
      }
    }
#pragma unroll
    for (int u = 0; u < FB_LD; ++u)
      if (dst[u] >= 0) stage[dst[u]] = v[u];
  }
  __syncthreads();
  float w[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) w[i] = kf[i];
  float* op = out + plane * (int64_t)p.out_h * p.out_w;
  const int quad_rows = (rows_out + 3) >> 2;
  const int nquads = quad_rows * p.out_w;
  for (int j = threadIdx.x; j < nquads; j += FB_THREADS) {
    const uint32_t qd = fdiv((uint32_t)j, doutw);
    const int ox = j - (int)qd * p.out_w;
    const int ly = 4 * (int)qd;
    const float* sp = stage + ly * SW + ox;
    float r[K + 3][K];
#pragma unroll
    for (int dy = 0; dy < K + 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < K; ++dx) r[dy][dx] = (ly + dy < rows_in) ? sp[dy * SW + dx] : 0.f;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < K; ++ky)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float wv = w[ky * K + kx];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += wv * r[ky + i][kx];
      }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (ly + i < rows_out) __stcs(op + (int64_t)(oy0 + ly + i) * p.out_w + ox, acc[i]);
  }
}

template <int K>
bool launch_band(float* out, const float* x, const float* kernel, int64_t planes, const UfdParams& p, int pad_x1,
                 cudaStream_t st) {
  const int SW = p.pad_x0 + p.in_w + (pad_x1 > 0 ? pad_x1 : 0);
  if (p.out_w - 1 + K - 1 >= SW) return false;  // the stencil of the last output column must stay inside the staged row
  // rows per band: a multiple of 4 (quads) that keeps the 256 threads busy and the staging inside 48 KB
  int R = 0;
  double best = -1.0;
  for (int cand = 4; cand <= 64; cand += 4) {
    if ((int64_t)(cand + K - 1) * SW > 12000) break;
    const int quads = (cand / 4) * p.out_w;
    const double util = (double)quads / (double)(((quads + FB_THREADS - 1) / FB_THREADS) * FB_THREADS);
    const int nb = (p.out_h + cand - 1) / cand;
    const double tail = (double)p.out_h / (double)(nb * cand);
    const double halo = (double)cand / (double)(cand + K - 1);
    const double score = util * tail * halo;
    if (score > best) {
      best = score;
      R = cand;
    }
  }
  if (R == 0) return false;
  const int bands = (p.out_h + R - 1) / R;
  const int64_t blocks = planes * bands;
  if (blocks > 2147483647LL) return false;
  const size_t smem = (size_t)(R + K - 1) * SW * sizeof(float);
  upfirdn2d_band<K><<<(unsigned)blocks, FB_THREADS, smem, st>>>(out, x, kernel, p, R, bands, SW, make_fastdiv((uint32_t)p.in_w),
                                                               make_fastdiv((uint32_t)p.out_w));
  return true;
}

// Streamed FIR (K x K taps, K <= 4, zero padding; up = down = 1, or up 2, or down 2): the default for the generator's Blur
// (3x3, pad 0), the discriminator's Blur (4x4, pad 2 / 1), Upsample / Downsample (models/ops.py:32-79) and their gradients.
// Persistent CTAs (two per SM); a work item is a group of P whole planes (small images) or a band of R output rows of one
// plane (large images) — in both cases ONE contiguous run of floats, fetched by one bulk-async copy into a two-stage ring
// (stream_stage.cuh) while the previous item is filtered.
//
// The rows are staged exactly as they lie in memory (no zero borders).  A thread owns ONE output column for the whole
// kernel (threads = G groups x columns; the groups share the 8-row strips of an item), so everything that depends on the
// column is set up once per item: each of its K taps is a shared-memory address plus a row stride, and a tap that falls
// outside the image points at a zero word with stride 0 — no column masks, no bounds checks.  For up = down = 1 the columns
// whose K taps all lie inside the image sit on the leading warps and use one running address with immediate tap offsets;
// the few border columns run the general form on one trailing warp, so no warp executes both.  The inner loop is a sliding
// window over the staged rows: K shared-memory reads per row, taps held as packed pairs so that a K x K stencil is
// K * (K/2) FFMA2 (+ K FFMA for K = 3) per output.  Only the first / last strips of an image (rows in the zero padding)
// take a variant with a clamped row index and a zero multiplier.  Up 2: see fir_strip_up2.  ncu
// (profiles/r2_ncu_hbm_kernels.txt): the first streamed version decoded (plane, strip, column) and rebuilt its column state
// for every 8 outputs — 33 / 44 instructions per output for K = 3 / 4 at 0.85 / 0.61 of the HBM peak; now 28 / 39 warp-level
// (idle lanes included) at 0.88 / 0.65.  The launch plan is host arithmetic (plan_fir_stream), swept on the CPU by
// tests/test_stream_plans.py together with a numpy restatement of the index arithmetic below.
constexpr int FS_MAX_THREADS = 320;
constexpr int FS_STRIP = 8;

struct FirStream {
  int in_h, in_w, out_h, out_w, pad_x0, pad_y0;
  int P;       // planes per item (bands == 1)
  int bands;   // bands per plane (P == 1 when bands > 1)
  int R;       // output rows per band
  int strips;  // ceil(R / FS_STRIP)
  int G;       // thread groups sharing the strips of an item (threads g * out_w + column)
  int kh, kw;  // kernel size (<= K: smaller kernels are zero-padded)
  int n_int, int_lo;             // interior columns [int_lo, int_lo + n_int): all K taps inside the image (0 = uniform mapping)
  int n_bord, Gb, border_base;   // border columns, their thread groups, first thread of the border warp
  FastDiv d_int;
  int stage_floats;
  int64_t planes, nitems;
  uintptr_t limit;  // 16-byte floor of the end of x
  FastDiv d_ow, d_strips;
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned long long fs_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}

template <int K>
struct FirTaps {
  unsigned long long w2[K][K / 2 > 0 ? K / 2 : 1];  // packed column pairs of every tap row
  float ws[K];                                      // last column when K is odd
};

// One strip of FS_STRIP output rows of one column.  a0[kx] = shared-memory address of tap column kx in staged row 0 of the
// plane (or of the zero word), rs[kx] = its row stride in bytes (0 for the zero word).  CHECK = false: every window row is
// staged and every output row exists (interior strips).  CHECK = true (first / last strips): rows outside the image are
// read from a clamped row and multiplied by zero, missing output rows are not stored — still branch-free.
template <int K, int DOWN, bool CHECK, bool INTERIOR>
__device__ __forceinline__ void fir_strip(const FirTaps<K>& t, const uint32_t (&a0)[K], const uint32_t (&rs)[K],
                                          float* __restrict__ o, int row0, int nrows, int out_w, int rows_left) {
  constexpr int NP = K / 2;
  constexpr bool ODD = (K & 1) != 0;
  // INTERIOR: all K taps of the column lie inside the image, so they sit at a0[0] + 4 * kx — ONE running address with
  // immediate offsets (LDS takes register + immediate only) instead of K running addresses
  uint32_t a[K];
#pragma unroll
  for (int kx = 0; kx < K; ++kx) a[kx] = CHECK ? a0[kx] : a0[kx] + (uint32_t)row0 * rs[kx];
  unsigned long long win2[K][NP > 0 ? NP : 1];
  float wins[K];
#pragma unroll
  for (int r = 0; r < (FS_STRIP - 1) * DOWN + K; ++r) {
    float v[K];
    if (CHECK) {
      const int ry = row0 + r;  // staged row index; outside [0, nrows) = zero padding
      const float mr = (ry >= 0 && ry < nrows) ? 1.f : 0.f;
      const uint32_t rc = (uint32_t)min(max(ry, 0), nrows - 1);
      if (INTERIOR) {
        const uint32_t ar = a[0] + rc * rs[0];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) v[kx] = lds_f32(ar + 4u * kx) * mr;
      } else {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) v[kx] = lds_f32(a[kx] + rc * rs[kx]) * mr;
      }
    } else if (INTERIOR) {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) v[kx] = lds_f32(a[0] + 4u * kx);
      a[0] += rs[0];
    } else {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        v[kx] = lds_f32(a[kx]);
        a[kx] += rs[kx];
      }
    }
#pragma unroll
    for (int pi = 0; pi < NP; ++pi) win2[r % K][pi] = fs_pack(v[2 * pi], v[2 * pi + 1]);
    if (ODD) wins[r % K] = v[K - 1];
    if (r >= K - 1 && (r - (K - 1)) % DOWN == 0) {  // down-sampling keeps every DOWN-th window position
      const int i = (r - (K - 1)) / DOWN;
      unsigned long long acc2 = 0ull;
      float accs = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int pi = 0; pi < NP; ++pi)
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2) : "l"(t.w2[ky][pi]), "l"(win2[(i * DOWN + ky) % K][pi]));
        if (ODD) accs = fmaf(t.ws[ky], wins[(i * DOWN + ky) % K], accs);
      }
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2));
      const float res = ODD ? (lo + hi) + accs : lo + hi;
      if (!CHECK || i < rows_left) __stcs(o, res);
      o += out_w;
    }
  }
}

// All strips of one output column that fall to thread group `grp` of `groups` (UP == 1).
template <int K, int DOWN, bool INTERIOR>
__device__ __forceinline__ void fir_column(const FirTaps<K>& taps, const FirStream& q, int ox, int grp, int groups, int units,
                                           uint32_t st_addr, uint32_t zero_addr, uint32_t plane_bytes, float* __restrict__ op,
                                           int oy0, int rows_out, int iy_lo, int nrows) {
  const uint32_t row_bytes = (uint32_t)q.in_w * 4u;
  uint32_t a0[K], rs[K];
#pragma unroll
  for (int kx = 0; kx < K; ++kx) {
    const int ix = ox * DOWN - q.pad_x0 + kx;
    const bool ok = INTERIOR || (ix >= 0 && ix < q.in_w);
    a0[kx] = ok ? st_addr + (uint32_t)ix * 4u : zero_addr;
    rs[kx] = ok ? row_bytes : 0u;
  }
  for (int u = grp; u < units; u += groups) {
    const int pl = (int)fdiv((uint32_t)u, q.d_strips);
    const int ly0 = (u - pl * q.strips) * FS_STRIP;  // first output row of the strip inside the band
    const int rows_left = rows_out - ly0;
    if (rows_left <= 0) continue;
    uint32_t ap[K];
    const uint32_t poff = (uint32_t)pl * plane_bytes;
#pragma unroll
    for (int kx = 0; kx < K; ++kx) ap[kx] = a0[kx] + (rs[kx] ? poff : 0u);
    const int row0 = (oy0 + ly0) * DOWN - q.pad_y0 - iy_lo;  // staged row of window row 0
    float* o = op + (int64_t)pl * q.out_h * q.out_w + (int64_t)ly0 * q.out_w + ox;
    if (row0 >= 0 && row0 + (FS_STRIP - 1) * DOWN + K - 1 < nrows && rows_left >= FS_STRIP)
      fir_strip<K, DOWN, false, INTERIOR>(taps, ap, rs, o, row0, nrows, q.out_w, rows_left);
    else
      fir_strip<K, DOWN, true, INTERIOR>(taps, ap, rs, o, row0, nrows, q.out_w, rows_left);
  }
}

// Up-sampling by two (zero stuffing): an output sees only the taps whose parity matches its position — two columns and two
// rows of a 4 x 4 kernel.  A thread owns one output column, so its column parity, its two input columns and its 4 x 2 taps
// are fixed; consecutive output rows alternate between tap rows (0, 2) and (1, 3) and advance one input row every second
// output.  ODD0 = parity of the first output row's tap origin (uniform over an item).  a0 / rs as in fir_strip.
template <bool ODD0>
__device__ __forceinline__ void fir_strip_up2(const float (&wc)[4][2], const uint32_t (&a0)[2], const uint32_t (&rs)[2],
                                              float* __restrict__ o, int j0, int nrows, int out_w, int rows_left) {
  constexpr int NR = ODD0 ? FS_STRIP / 2 + 1 : FS_STRIP / 2 + 2;
  float v[NR][2];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int ry = j0 + r;
    const float mr = (ry >= 0 && ry < nrows) ? 1.f : 0.f;
    const uint32_t rc = (uint32_t)min(max(ry, 0), nrows - 1);
    v[r][0] = lds_f32(a0[0] + rc * rs[0]) * mr;
    v[r][1] = lds_f32(a0[1] + rc * rs[1]) * mr;
  }
#pragma unroll
  for (int i = 0; i < FS_STRIP; ++i) {
    const int A = ODD0 ? (i >> 1) : ((i + 1) >> 1);
    const int par = ((ODD0 ? 1 : 0) + i) & 1;
    float res = wc[par][0] * v[A][0];
    res = fmaf(wc[par][1], v[A][1], res);
    res = fmaf(wc[par + 2][0], v[A + 1][0], res);
    res = fmaf(wc[par + 2][1], v[A + 1][1], res);
    if (i < rows_left) __stcs(o, res);
    o += out_w;
  }
}

template <int K, int UP, int DOWN>
__global__ void __launch_bounds__(FS_MAX_THREADS, 2) fir_stream_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                                      const float* __restrict__ kernel, FirStream q) {
  extern __shared__ __align__(128) float fs_smem[];
  __shared__ uint64_t bars[2];
  __shared__ float kf[K * K];
  __shared__ float zero_word;
  const int tid = threadIdx.x;
  if (tid < K * K) {  // flipped taps, zero-padded to K x K when the kernel is smaller (up / down variants)
    const int ky = tid / K, kx = tid - ky * K;
    kf[tid] = (ky < q.kh && kx < q.kw) ? kernel[(q.kh - 1 - ky) * q.kw + (q.kw - 1 - kx)] : 0.f;
  }
  if (tid == 0) {
    zero_word = 0.f;
    stream_stage::bar_init(stream_stage::smem_addr(&bars[0]), 1);
    stream_stage::bar_init(stream_stage::smem_addr(&bars[1]), 1);
    stream_stage::fence_bar_init();
  }
  __syncthreads();
  const int64_t plane_floats = (int64_t)q.in_h * q.in_w;
  // item -> (first plane, plane count, first output row, output rows, first / last staged input row)
  auto decode = [&](int64_t item, int64_t& plane0, int& np, int& oy0, int& rows_out, int& iy_lo, int& iy_hi) {
    if (q.bands == 1) {
      plane0 = item * q.P;
      np = (int)min((int64_t)q.P, q.planes - plane0);
      oy0 = 0;
      rows_out = q.out_h;
      iy_lo = 0;
      iy_hi = q.in_h - 1;
    } else {
      plane0 = item / q.bands;
      const int band = (int)(item - plane0 * q.bands);
      np = 1;
      oy0 = band * q.R;
      rows_out = min(q.R, q.out_h - oy0);
      const int y_first = oy0 * DOWN - q.pad_y0, y_last = (oy0 + rows_out - 1) * DOWN - q.pad_y0 + K - 1;
      iy_lo = max(UP == 2 ? (y_first + 1) >> 1 : y_first, 0);
      iy_hi = min(UP == 2 ? y_last >> 1 : y_last, q.in_h - 1);
    }
  };
  auto issue = [&](int64_t item, int s) {
    int64_t plane0;
    int np, oy0, rows_out, iy_lo, iy_hi;
    decode(item, plane0, np, oy0, rows_out, iy_lo, iy_hi);
    const int n = (int)((np - 1) * plane_floats) + (iy_hi - iy_lo + 1) * q.in_w;
    stream_stage::issue_chunk(fs_smem + s * q.stage_floats, x + plane0 * plane_floats + (int64_t)iy_lo * q.in_w, n, q.limit,
                              &bars[s]);
  };
  FirTaps<K> taps;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
#pragma unroll
    for (int pi = 0; pi < K / 2; ++pi) taps.w2[ky][pi] = fs_pack(kf[ky * K + 2 * pi], kf[ky * K + 2 * pi + 1]);
    taps.ws[ky] = (K & 1) ? kf[ky * K + K - 1] : 0.f;
  }
  // thread -> (group, first column); wide images (out_w > threads) walk their columns with one group
  const int nt = blockDim.x;
  const int grp = (int)fdiv((uint32_t)tid, q.d_ow);
  const int col0 = tid - grp * q.out_w;
  const uint32_t zero_addr = stream_stage::smem_addr(&zero_word);
  const uint32_t row_bytes = (uint32_t)q.in_w * 4u;
  int64_t item = blockIdx.x;
  if (tid == 0 && item < q.nitems) issue(item, 0);
  for (uint32_t k = 0; item < q.nitems; item += gridDim.x, ++k) {
    const int s = k & 1;
    if (tid == 0 && item + gridDim.x < q.nitems) issue(item + gridDim.x, s ^ 1);
    int64_t plane0;
    int np, oy0, rows_out, iy_lo, iy_hi;
    decode(item, plane0, np, oy0, rows_out, iy_lo, iy_hi);
    const float* src = x + plane0 * plane_floats + (int64_t)iy_lo * q.in_w;
    const uint32_t st_addr = stream_stage::smem_addr(fs_smem + s * q.stage_floats + stream_stage::chunk_shift(src));
    float* op = out + plane0 * (int64_t)q.out_h * q.out_w + (int64_t)oy0 * q.out_w;
    const int nrows = iy_hi - iy_lo + 1;  // staged rows of the (last) plane of the item
    const int units = np * q.strips;
    stream_stage::bar_wait(stream_stage::smem_addr(&bars[s]), (k >> 1) & 1);
    if (UP == 1 && q.n_int > 0) {
      // split mapping: interior columns (all taps inside the image) on the leading warps, the few border columns on one
      // trailing warp, so that no warp runs both variants
      if (tid < q.G * q.n_int) {
        const int g2 = (int)fdiv((uint32_t)tid, q.d_int);
        fir_column<K, DOWN, true>(taps, q, q.int_lo + tid - g2 * q.n_int, g2, q.G, units, st_addr, zero_addr,
                                  (uint32_t)plane_floats * 4u, op, oy0, rows_out, iy_lo, nrows);
      } else if (tid >= q.border_base && tid - q.border_base < q.Gb * q.n_bord) {
        const int j = tid - q.border_base;
        const int g2 = j / q.n_bord, b = j - g2 * q.n_bord;
        const int ox = b < q.int_lo ? b : q.n_int + b;  // columns left of the interior, then right of it
        fir_column<K, DOWN, false>(taps, q, ox, g2, q.Gb, units, st_addr, zero_addr, (uint32_t)plane_floats * 4u, op, oy0,
                                   rows_out, iy_lo, nrows);
      }
    } else if (grp < q.G) {
      for (int ox = col0; ox < q.out_w; ox += nt) {
        if (UP == 1) {
          fir_column<K, DOWN, false>(taps, q, ox, grp, q.G, units, st_addr, zero_addr, (uint32_t)plane_floats * 4u, op, oy0,
                                     rows_out, iy_lo, nrows);
        } else {
          // up 2: this column's tap parity, its two input columns and its 4 x 2 taps
          const int c = ox - q.pad_x0;
          const int kx0 = c & 1;
          const int ix0 = (c + kx0) >> 1;
          uint32_t a0[2], rs[2];
          float wc[4][2];
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const bool ok = ix0 + b >= 0 && ix0 + b < q.in_w;
            a0[b] = ok ? st_addr + (uint32_t)(ix0 + b) * 4u : zero_addr;
            rs[b] = ok ? row_bytes : 0u;
#pragma unroll
            for (int ky = 0; ky < 4; ++ky) wc[ky][b] = kx0 ? kf[ky * K + 1 + 2 * b] : kf[ky * K + 2 * b];
          }
          const int par0 = (oy0 - q.pad_y0) & 1;  // strips start at multiples of 8 rows: one parity per item
          for (int u = grp; u < units; u += q.G) {
            const int pl = (int)fdiv((uint32_t)u, q.d_strips);
            const int ly0 = (u - pl * q.strips) * FS_STRIP;
            const int rows_left = rows_out - ly0;
            if (rows_left <= 0) continue;
            uint32_t ap[2];
            const uint32_t poff = (uint32_t)pl * (uint32_t)plane_floats * 4u;
            ap[0] = a0[0] + (rs[0] ? poff : 0u);
            ap[1] = a0[1] + (rs[1] ? poff : 0u);
            const int t0 = oy0 + ly0 - q.pad_y0;
            const int j0 = ((t0 + par0) >> 1) - iy_lo;  // staged row of the first output's upper input row
            float* o = op + (int64_t)pl * q.out_h * q.out_w + (int64_t)ly0 * q.out_w + ox;
            if (par0) fir_strip_up2<true>(wc, ap, rs, o, j0, nrows, q.out_w, rows_left);
            else fir_strip_up2<false>(wc, ap, rs, o, j0, nrows, q.out_w, rows_left);
          }
        }
      }
    }
    __syncthreads();  // every read of stage s is done before the copy of item k + 2 is issued into it
  }
}

constexpr int FS_MAX_STAGE = 11776;  // floats per stage: 46 KB, two stages and two CTAs per SM

// Host-side plan of a streamed launch: work items, stage size, thread mapping.  Pure arithmetic (no CUDA calls besides the SM
// count), exported as spgan_upfirdn2d_plan so that the CPU tests can sweep it for its invariants.
template <int K, int UP, int DOWN>
bool plan_fir_stream(int64_t planes, const UfdParams& p, FirStream& q, int& block, unsigned& grid) {
  static_assert(UP == 1 || (K == 4 && DOWN == 1), "up 2 uses the zero-padded 4 x 4 form");
  constexpr int MAX_STAGE = FS_MAX_STAGE;
  const int64_t plane_floats = (int64_t)p.in_h * p.in_w;
  const int ncta = 2 * SPGAN_NUM_SMS;
  q.in_h = p.in_h; q.in_w = p.in_w; q.out_h = p.out_h; q.out_w = p.out_w; q.pad_x0 = p.pad_x0; q.pad_y0 = p.pad_y0;
  q.kh = p.kh; q.kw = p.kw;
  q.planes = planes;
  if (plane_floats + 8 <= MAX_STAGE) {
    // whole planes: as many per item as fit, but no more than spreads the planes over all CTAs
    int64_t P = (MAX_STAGE - 8) / plane_floats;
    const int64_t spread = (planes + ncta - 1) / ncta;
    if (P > spread) P = spread;
    if (P < 1) P = 1;
    q.P = (int)P;
    q.bands = 1;
    q.R = p.out_h;
    q.nitems = (planes + P - 1) / P;
    q.stage_floats = (int)(P * plane_floats + 8 + 31) / 32 * 32;
  } else {
    // bands: R output rows need ((R - 1) * DOWN + K - 1) / UP + 2 input rows at most
    const int rows_fit = (MAX_STAGE - 8) / p.in_w;
    int R = ((rows_fit - 2) * UP - (K - 1)) / DOWN + 1;
    if (rows_fit < 3 || R < 1) return false;  // a single row group does not fit: tiled / polyphase kernel
    R = R / FS_STRIP * FS_STRIP;              // whole strips, so that every band starts on the same tap parity
    if (R < FS_STRIP) return false;
    if (R > p.out_h) R = p.out_h;
    const int bands = (p.out_h + R - 1) / R;
    if (bands > 1) {
      R = ((p.out_h + bands - 1) / bands + FS_STRIP - 1) / FS_STRIP * FS_STRIP;
    }
    q.P = 1;
    q.bands = (p.out_h + R - 1) / R;
    q.R = R;
    q.nitems = planes * q.bands;
    q.stage_floats = (rows_fit * p.in_w + 8 + 31) / 32 * 32;
  }
  q.strips = (q.R + FS_STRIP - 1) / FS_STRIP;
  // threads = G groups x out_w columns (rounded up to whole warps): the G that wastes the fewest lanes and strip rounds
  const int units = q.P * q.strips;
  int best_g = 1, best_nt = FS_MAX_THREADS;
  if (p.out_w <= FS_MAX_THREADS) {
    double best = -1.0;
    for (int g = 1; g * p.out_w <= FS_MAX_THREADS && g <= units; ++g) {
      const int nt = (g * p.out_w + 31) / 32 * 32;
      if (nt < 96 && (g + 1) * p.out_w <= FS_MAX_THREADS && g + 1 <= units) continue;  // too few warps to hide latency
      const double lanes = (double)(g * p.out_w) / nt;
      const double rounds = (double)units / (double)(((units + g - 1) / g) * g);
      const double score = lanes * rounds * (0.85 + 0.15 * nt / FS_MAX_THREADS);
      if (score > best) {
        best = score;
        best_g = g;
        best_nt = nt;
      }
    }
  }
  q.n_int = 0; q.int_lo = 0; q.n_bord = 0; q.Gb = 1; q.border_base = 0;
  q.d_int = make_fastdiv(1);
  if (UP == 1 && DOWN == 1 && p.out_w + 32 <= FS_MAX_THREADS) {  // (down 2 is bound by its strided shared-memory reads)
    // interior columns: 0 <= ox * DOWN - pad_x0 and ox * DOWN - pad_x0 + K - 1 <= in_w - 1
    const int lo = (p.pad_x0 + DOWN - 1) / DOWN;
    int hi = (p.in_w - K + p.pad_x0) >= 0 ? (p.in_w - K + p.pad_x0) / DOWN : -1;  // last interior column
    if (hi > p.out_w - 1) hi = p.out_w - 1;
    const int n_int = hi - lo + 1, n_bord = p.out_w - n_int;
    if (n_int >= 8 && n_bord <= 32) {
      double best = -1.0;
      int g_best = 0, nt_best = 0;
      const int bw = n_bord > 0 ? 32 : 0;  // one trailing warp for the border columns
      for (int g = 1; g * n_int + bw <= FS_MAX_THREADS && g <= units; ++g) {
        const int lead = (g * n_int + 31) / 32 * 32;
        if (lead + bw > FS_MAX_THREADS) break;
        if (lead + bw < 96 && (g + 1) * n_int + bw <= FS_MAX_THREADS && g + 1 <= units) continue;
        const double lanes = (double)(g * n_int) / lead;
        const double rounds = (double)units / (double)(((units + g - 1) / g) * g);
        const double score = lanes * rounds * (0.85 + 0.15 * (lead + bw) / FS_MAX_THREADS);
        if (score > best) {
          best = score;
          g_best = g;
          nt_best = lead + bw;
        }
      }
      if (g_best > 0) {
        best_g = g_best;
        best_nt = nt_best;
        q.n_int = n_int;
        q.int_lo = lo;
        q.n_bord = n_bord;
        q.border_base = nt_best - bw;
        q.Gb = n_bord > 0 ? (32 / n_bord < units ? 32 / n_bord : units) : 1;
        if (q.Gb < 1) q.Gb = 1;
        q.d_int = make_fastdiv((uint32_t)n_int);
      }
    }
  }
  if (p.out_w > FS_MAX_THREADS) {
    // wide images: every thread walks `passes` columns; size the block so that the last pass is as full as the others
    const int passes = (p.out_w + FS_MAX_THREADS - 1) / FS_MAX_THREADS;
    best_nt = ((p.out_w + passes - 1) / passes + 31) / 32 * 32;
  }
  q.G = best_g;
  q.limit = 0;
  q.d_ow = make_fastdiv((uint32_t)p.out_w);
  q.d_strips = make_fastdiv((uint32_t)q.strips);
  block = best_nt;
  grid = (unsigned)(q.nitems < ncta ? q.nitems : ncta);
  return true;
}

template <int K, int UP, int DOWN>
bool launch_fir_stream(float* out, const float* x, const float* kernel, int64_t planes, const UfdParams& p,
                       cudaStream_t st) {
  if ((((uintptr_t)x) & 15) != 0) return false;
  FirStream q;
  int block = 0;
  unsigned grid = 0;
  if (!plan_fir_stream<K, UP, DOWN>(planes, p, q, block, grid)) return false;
  q.limit = ((uintptr_t)(x + planes * (int64_t)p.in_h * p.in_w)) & ~(uintptr_t)15;
  const size_t smem = (size_t)2 * q.stage_floats * sizeof(float);
  static bool attr_done[64] = {false};
  if (!spgan_allow_smem(fir_stream_kernel<K, UP, DOWN>, 2 * FS_MAX_STAGE * 4 + 1024, attr_done)) return false;
  fir_stream_kernel<K, UP, DOWN><<<grid, block, smem, st>>>(out, x, kernel, q);
  return true;
}

// Which streamed kernel serves a geometry: K * 100 + UP * 10 + DOWN, or 0 (tiled / band / polyphase / generic kernels).
int fir_stream_variant(const UfdParams& p, int up_x, int up_y, int down_x, int down_y, int pad_x1, int pad_y1) {
  const bool unit = up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && p.kh == p.kw && p.pad_x0 >= 0 && p.pad_y0 >= 0;
  if (unit && p.kh >= 2 && p.kh <= 4 && p.pad_x0 <= p.kh - 1 && p.pad_y0 <= p.kh - 1 && pad_x1 <= p.kh - 1 &&
      pad_y1 <= p.kh - 1 && p.in_w >= 1)
    return p.kh * 100 + 11;
  // Upsample / Downsample of models/ops.py:32-79 and each other's gradient: up or down by two, kernels up to 4 x 4
  const bool small_k = p.kh >= 1 && p.kw >= 1 && p.kh <= 4 && p.kw <= 4 && p.pad_x0 >= 0 && p.pad_y0 >= 0 && p.pad_x0 <= 3 &&
                       p.pad_y0 <= 3 && pad_x1 <= 3 && pad_y1 <= 3 && p.in_w >= 1;
  if (small_k && up_x == 2 && up_y == 2 && down_x == 1 && down_y == 1) return 421;
  if (small_k && up_x == 1 && up_y == 1 && down_x == 2 && down_y == 2) return 412;
  return 0;
}

// Fused tail of the upsampling StyledConv: interleave the four polyphase planes of the transposed conv on the fly,
// 3x3 FIR, + noise + bias, leaky-ReLU * scale.  HBM traffic = one read of the planes (+ one halo row per band) and one
// write of the result, instead of scatter-write + FIR read/write + activation read/write.
//
// A CTA owns a band of A output-row PAIRS of one (sample, channel) plane at full width.  The plane rows it needs are
// contiguous in each of the four polyphase planes, so staging is four flat coalesced copies (no index arithmetic, all
// loads of a thread in flight before the first shared-memory store).  Each thread then produces vertical output pairs
// (2a, ox), (2a + 1, ox): 12 shared-memory reads and 18 FMAs per 2 outputs, lanes along ox (coalesced stores; the two
// column parities sit 16 banks apart, so the reads are conflict-free).  ~30 instructions per output instead of ~65 for
// a one-output-per-thread interleaving kernel, which was issue-bound at 27 % of the HBM roofline.
constexpr int UB_THREADS = 256;
constexpr int UB_LD = 4;  // staged loads per thread and plane on the fast path ((A + 2) * Wq <= 768)

__global__ void __launch_bounds__(UB_THREADS) upblur_act_kernel(float* __restrict__ out, const float* __restrict__ pp,
                                                               const float* __restrict__ kernel,
                                                               const float* __restrict__ noise,
                                                               const float* __restrict__ noise_w,
                                                               const float* __restrict__ bias, int64_t channels, int Hq,
                                                               int Wq, int oh, int ow, int A, int bands, int pstride,
                                                               FastDiv dow, float alpha, float scale) {
  extern __shared__ float stage[];  // [4][pstride]
  __shared__ float kf[9];
  if (threadIdx.x < 9) {
    int ky = threadIdx.x / 3, kx = threadIdx.x - ky * 3;
    kf[threadIdx.x] = kernel[(2 - ky) * 3 + (2 - kx)];
  }
  const int64_t plane = blockIdx.x / bands;
  const int band = blockIdx.x - (int)(plane * bands);
  const int a0 = band * A;                       // first output-row pair of the band
  const int rows = min(A + 2, Hq - a0);          // plane rows to stage (A pairs + halo; + 1 when A is odd: quads)
  const int n = rows * Wq;
  const int64_t Q = (int64_t)Hq * Wq;
  const float* src = pp + plane * 4 * Q + (int64_t)a0 * Wq;
  if (n <= UB_LD * UB_THREADS) {
    float v[4][UB_LD];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int u = 0; u < UB_LD; ++u) {
        const int e = u * UB_THREADS + threadIdx.x;
        v[ph][u] = e < n ? __ldcs(src + ph * Q + e) : 0.f;
      }
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int u = 0; u < UB_LD; ++u) {
        const int e = u * UB_THREADS + threadIdx.x;
        if (e < n) stage[ph * pstride + e] = v[ph][u];
      }
  } else {
    for (int ph = 0; ph < 4; ++ph)
      for (int e = threadIdx.x; e < n; e += UB_THREADS) stage[ph * pstride + e] = __ldcs(src + ph * Q + e);
  }
  __syncthreads();
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = kf[i];
  const int64_t b = plane / channels, c = plane - b * channels;
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const float bv = bias ? __ldg(bias + c) : 0.f;
  const int64_t opl = (int64_t)oh * ow;
  const float* np = noise ? noise + b * opl : nullptr;
  float* op = out + plane * opl;
  // Each thread produces a vertical QUAD of outputs (rows 4q .. 4q+3 of the band at one column): 18 shared-memory reads
  // and 36 FMAs per 4 outputs.  Column parity selects the plane pair once per quad; every other address term is a
  // warp-uniform constant, so a read is one LDS with an immediate-style offset.
  const int pair_rows = min(A, (oh + 1) / 2 - a0);
  const int quad_rows = (pair_rows + 1) >> 1;
  const int nquads = quad_rows * ow;
  const int r1o = Wq, r2o = 2 * Wq;          // plane-row offsets
  const int odd_row = 2 * pstride;           // planes 2, 3 hold the odd interleaved rows
  for (int j = threadIdx.x; j < nquads; j += UB_THREADS) {
    const uint32_t qd = fdiv((uint32_t)j, dow);
    const int ox = j - (int)qd * ow;
    const int a = 2 * (int)qd;               // first row pair of the quad inside the band
    const int oy0 = 2 * (a0 + a);
    const int p0 = ox & 1;
    // column taps: dx = 0 -> (parity p0, col ox>>1), dx = 1 -> (parity 1-p0, col (ox+1)>>1), dx = 2 -> (p0, (ox>>1)+1)
    const float* e0 = stage + p0 * pstride + a * Wq + (ox >> 1);
    const float* e1 = stage + (1 - p0) * pstride + a * Wq + ((ox + 1) >> 1);
    // interleaved rows oy0 + dy, dy = 0..5: plane row a + (dy >> 1), odd rows in planes 2, 3
    float r[6][3];
#pragma unroll
    for (int dy = 0; dy < 6; ++dy) {
      const int off = (dy & 1) * odd_row + (dy >> 1 == 0 ? 0 : (dy >> 1 == 1 ? r1o : r2o));
      r[dy][0] = e0[off];
      r[dy][1] = e1[off];
      r[dy][2] = e0[off + 1];
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float wv = w[ky * 3 + kx];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += wv * r[ky + i][kx];
      }
    const int64_t o0 = (int64_t)oy0 * ow + ox;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (oy0 + i < oh) {
        const int64_t o = o0 + (int64_t)i * ow;
        const float v = acc[i] + bv + (np ? nw * __ldg(np + o) : 0.f);
        __stcs(op + o, (v > 0.f ? v : v * alpha) * scale);
      }
    }
  }
}

}  // namespace

extern "C" int spgan_upblur_act(float* out, const float* pp, const float* kernel, const float* noise,
                                const float* noise_w, const float* bias, int64_t batch, int64_t channels, int zh, int zw,
                                int Hq, int Wq, float alpha, float scale, void* stream) {
  SPGAN_CHECK_ARG(batch >= 0 && channels >= 0 && zh >= 0 && zw >= 0, "spgan_upblur_act: negative size");
  SPGAN_CHECK_ARG(Hq * 2 >= zh && Wq * 2 >= zw, "spgan_upblur_act: polyphase planes %dx%d too small for %dx%d", Hq, Wq, zh, zw);
  const int oh = zh - 2, ow = zw - 2;
  if (batch * channels == 0 || oh <= 0 || ow <= 0) return 0;
  SPGAN_CHECK_ARG(out && pp && kernel, "spgan_upblur_act: null pointer");
  SPGAN_CHECK_ARG((noise == nullptr) == (noise_w == nullptr), "spgan_upblur_act: noise and noise_w go together");
  SPGAN_CHECK_ARG((int64_t)zh * zw < (1LL << 30), "spgan_upblur_act: image %dx%d too large", zh, zw);
  const int pair_rows = (oh + 1) / 2;
  // band height A (output-row pairs per CTA): the even A that keeps the 256 threads busiest (quads per CTA just below
  // a multiple of 256, little waste in the last band) among those whose staging fits the fast path and 48 KB
  int A = 2;
  double best = -1.0;
  for (int cand = 2; cand <= 64; cand += 2) {
    if ((cand + 2) * Wq > UB_LD * UB_THREADS && cand > 2) break;
    if ((cand + 2) * Wq + 50 > 3072) break;
    const int quads = (cand / 2) * ow;
    const double util = (double)quads / (double)(((quads + UB_THREADS - 1) / UB_THREADS) * UB_THREADS);
    const int nb = (pair_rows + cand - 1) / cand;
    const double tail = (double)pair_rows / (double)(nb * cand);
    const double halo = (double)cand / (double)(cand + 2);  // staged rows that are not halo
    const double score = util * tail * (0.5 + 0.5 * halo);
    if (score > best) {
      best = score;
      A = cand;
    }
  }
  SPGAN_CHECK_ARG((A + 2) * Wq + 50 <= 3072, "spgan_upblur_act: rows of %d pixels exceed the staging buffer", zw);
  const int bands = (pair_rows + A - 1) / A;
  const int pstride = ((A + 2) * Wq + 2 + 31) / 32 * 32 + 16;  // planes 16 banks apart: column parities never collide
  const size_t smem = (size_t)4 * pstride * sizeof(float);
  SPGAN_CHECK_ARG(smem <= 48 * 1024, "spgan_upblur_act: rows of %d pixels exceed the staging buffer", zw);
  const int64_t blocks = batch * channels * bands;
  SPGAN_CHECK_ARG(blocks <= 2147483647LL, "spgan_upblur_act: too many bands");
  upblur_act_kernel<<<(unsigned)blocks, UB_THREADS, smem, (cudaStream_t)stream>>>(
      out, pp, kernel, noise, noise_w, bias, channels, Hq, Wq, oh, ow, A, bands, pstride, make_fastdiv((uint32_t)ow), alpha,
      scale);
  SPGAN_CHECK_LAUNCH("spgan_upblur_act");
  return 0;
}

extern "C" int spgan_upfirdn2d(float* out, const float* x, const float* kernel, int64_t planes, int in_h, int in_w,
                               int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                               int pad_y0, int pad_y1, void* stream) {
  SPGAN_CHECK_ARG(planes >= 0 && in_h >= 0 && in_w >= 0, "spgan_upfirdn2d: negative size");
  SPGAN_CHECK_ARG(kh >= 1 && kw >= 1 && kh * kw <= 256, "spgan_upfirdn2d: kernel %dx%d unsupported (max 256 taps)", kh, kw);
  SPGAN_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "spgan_upfirdn2d: up/down must be >= 1");
  UfdParams p;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  const int full_h = in_h * up_y + pad_y0 + pad_y1 - kh;
  const int full_w = in_w * up_x + pad_x0 + pad_x1 - kw;
  p.out_h = full_h >= 0 ? full_h / down_y + 1 : 0;
  p.out_w = full_w >= 0 ? full_w / down_x + 1 : 0;
  if (planes == 0 || p.out_h <= 0 || p.out_w <= 0) return 0;
  SPGAN_CHECK_ARG(out && x && kernel, "spgan_upfirdn2d: input and kernel must be CUDA tensors (null pointer)");
  cudaStream_t st = (cudaStream_t)stream;
  const bool unit = up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh == kw && pad_x0 >= 0 && pad_y0 >= 0;
  const int tiles_x = (p.out_w + TILE - 1) / TILE, tiles_y = (p.out_h + TILE - 1) / TILE;
  const int64_t blocks = planes * tiles_x * tiles_y;
  bool done = false;
  switch (spgan_legacy_hbm() ? 0 : fir_stream_variant(p, up_x, up_y, down_x, down_y, pad_x1, pad_y1)) {
    case 211: done = launch_fir_stream<2, 1, 1>(out, x, kernel, planes, p, st); break;
    case 311: done = launch_fir_stream<3, 1, 1>(out, x, kernel, planes, p, st); break;
    case 411: done = launch_fir_stream<4, 1, 1>(out, x, kernel, planes, p, st); break;
    case 421: done = launch_fir_stream<4, 2, 1>(out, x, kernel, planes, p, st); break;
    case 412: done = launch_fir_stream<4, 1, 2>(out, x, kernel, planes, p, st); break;
    default: break;
  }
  if (!done && unit && kh >= 2 && kh <= 4 && pad_x0 <= kh - 1 && pad_y0 <= kh - 1 && p.in_w >= 1) {
    if (kh == 2) done = launch_band<2>(out, x, kernel, planes, p, pad_x1, st);
    if (kh == 3) done = launch_band<3>(out, x, kernel, planes, p, pad_x1, st);
    if (kh == 4) done = launch_band<4>(out, x, kernel, planes, p, pad_x1, st);
  }
  if (done) {
  } else if (unit && kh >= 2 && kh <= 4 && blocks <= 2147483647LL) {
    if (kh == 2) upfirdn2d_tiled<2><<<(unsigned)blocks, 256, 0, st>>>(out, x, kernel, tiles_x, tiles_y, p);
    if (kh == 3) upfirdn2d_tiled<3><<<(unsigned)blocks, 256, 0, st>>>(out, x, kernel, tiles_x, tiles_y, p);
    if (kh == 4) upfirdn2d_tiled<4><<<(unsigned)blocks, 256, 0, st>>>(out, x, kernel, tiles_x, tiles_y, p);
  } else if (up_x == up_y && down_x == down_y && up_x <= 2 && down_x <= 2 && kh <= 4 && kw <= 4) {
    const int tx = (p.out_w + 31) / 32, ty = (p.out_h + 31) / 32;
    dim3 grid((unsigned)(tx * ty), (unsigned)(planes < 65535 ? planes : 65535));
    if (up_x == 1 && down_x == 1) upfirdn2d_poly<1, 1><<<grid, 256, 0, st>>>(out, x, kernel, planes, tx, p);
    else if (up_x == 2 && down_x == 1) upfirdn2d_poly<2, 1><<<grid, 256, 0, st>>>(out, x, kernel, planes, tx, p);
    else if (up_x == 1 && down_x == 2) upfirdn2d_poly<1, 2><<<grid, 256, 0, st>>>(out, x, kernel, planes, tx, p);
    else upfirdn2d_poly<2, 2><<<grid, 256, 0, st>>>(out, x, kernel, planes, tx, p);
  } else {
    const int64_t total = planes * p.out_h * p.out_w;
    upfirdn2d_generic<<<grid_for(total, 256, 8, 8), 256, 0, st>>>(out, x, kernel, planes, p);
  }
  SPGAN_CHECK_LAUNCH("spgan_upfirdn2d");
  return 0;
}

extern "C" int spgan_upfirdn2d_plan(int64_t planes, int in_h, int in_w, int kh, int kw, int up, int down, int pad_x0, int pad_x1,
                                    int pad_y0, int pad_y1, int32_t* plan) {
  SPGAN_CHECK_ARG(plan != nullptr, "spgan_upfirdn2d_plan: null plan");
  SPGAN_CHECK_ARG(planes >= 1 && in_h >= 1 && in_w >= 1 && kh >= 1 && kw >= 1 && up >= 1 && down >= 1,
                  "spgan_upfirdn2d_plan: sizes must be positive");
  for (int i = 0; i < 20; ++i) plan[i] = 0;
  UfdParams p;
  p.in_h = in_h; p.in_w = in_w; p.kh = kh; p.kw = kw;
  p.up_x = up; p.up_y = up; p.down_x = down; p.down_y = down;
  p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  const int full_h = in_h * up + pad_y0 + pad_y1 - kh, full_w = in_w * up + pad_x0 + pad_x1 - kw;
  p.out_h = full_h >= 0 ? full_h / down + 1 : 0;
  p.out_w = full_w >= 0 ? full_w / down + 1 : 0;
  plan[15] = p.out_h;
  plan[16] = p.out_w;
  if (p.out_h <= 0 || p.out_w <= 0) return 0;
  const int variant = fir_stream_variant(p, up, up, down, down, pad_x1, pad_y1);
  FirStream q;
  int block = 0;
  unsigned grid = 0;
  bool ok = false;
  switch (variant) {
    case 211: ok = plan_fir_stream<2, 1, 1>(planes, p, q, block, grid); break;
    case 311: ok = plan_fir_stream<3, 1, 1>(planes, p, q, block, grid); break;
    case 411: ok = plan_fir_stream<4, 1, 1>(planes, p, q, block, grid); break;
    case 421: ok = plan_fir_stream<4, 2, 1>(planes, p, q, block, grid); break;
    case 412: ok = plan_fir_stream<4, 1, 2>(planes, p, q, block, grid); break;
    default: break;
  }
  if (!ok) return 0;
  plan[0] = variant;
  plan[1] = q.P; plan[2] = q.bands; plan[3] = q.R; plan[4] = q.strips; plan[5] = q.G; plan[6] = block;
  plan[7] = q.stage_floats; plan[8] = q.n_int; plan[9] = q.int_lo; plan[10] = q.n_bord; plan[11] = q.Gb;
  plan[12] = q.border_base;
  plan[13] = (int32_t)(q.nitems < 2147483647LL ? q.nitems : 2147483647LL);
  plan[14] = (int32_t)grid;
  plan[17] = FS_MAX_STAGE;
  plan[18] = FS_MAX_THREADS;
  plan[19] = FS_STRIP;
  return 0;
}

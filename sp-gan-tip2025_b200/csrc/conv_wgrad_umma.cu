// Weight gradient of the conv passes on tcgen05 / TMEM (replaces cuDNN wgrad of the reference's F.conv2d /
// F.conv_transpose2d calls, models/ops.py:617, 634, 175; models/spgan_ops_gs.py:814).
//
//   dW[t][o][c] = sum_q G'[g_phase][q][o] * X'[phase_t][q + off_t][c],     q = (b*Hl + i)*Wl + j  (flattened lattice point)
//
// G' = out_mul * g and X' = in_mul * x are the SAME channels-last bf16 hi/lo packs the forward GEMM consumes
// (spgan_pack_act: [2][phases * B*Hl*Wl][Cp]) on one common lattice.  The contraction runs over the ROW index of both
// packs, so both operands are "MN-major" for tcgen05 (instruction-descriptor bits 15/16): a tile is a stack of TMA boxes
// of 64 pixel rows x 64 channels (128-byte swizzle), and a conv tap is a shift of the X' row coordinate.  (A
// pixel-contiguous K-major layout does not work: the tap shift would land on the INNER TMA coordinate, which must be a
// multiple of 16 bytes — an odd offset raises an illegal-instruction fault, see tools/probes/tma_unaligned.cu.)
// G' is zero at lattice points that are not outputs of the pass, which makes the row wrap-around of the flattened index
// harmless.  Strided convs read X' from its polyphase planes (phase_t), the parity passes of the transposed conv read
// G' from its polyphase planes (g_phase).
//
// GEMM shape: M = Cout (128-row tiles), N = Cin (256-column tiles), K = B*Hl*Wl split into `ksplit` chunks so that
// taps * tiles * ksplit work items fill the 148 SMs; items are ordered K-chunk-major, so the CTAs running at the same
// time read the same slice of both operands (L2 reuse).  Each item writes an fp32 partial tile; wgrad_reduce_kernel
// sums the K-chunks in a fixed order (deterministic) and scatters into the native (Cout, Cin, kh, kw) layout.
// Kernel anatomy as conv_gemm_kernel: warp 0 TMA producer, warp 1 single-thread tcgen05.mma issuer with two TMEM
// accumulator stages, warps 2..5 epilogue.  Roofline: tensor pipe, 2*Q*Cout*Cin*ntaps FLOP per launch (x3 issued in
// bf16x3 mode).
#include "umma_common.cuh"

namespace {

constexpr int CHUNK_BYTES = 64 * GEMM_BLOCK_K * 2;  // one TMA box: 64 pixel rows x 64 channels of bf16 = 8 KiB

struct WgradParams {
  int32_t Q;  // contraction length (lattice points per phase plane)
  int32_t O, C, Cs;  // M extent, N extent, padded row stride of the partial tiles
  int32_t ntaps;
  int32_t tap_row[SPGAN_MAX_TAPS];  // row offset of the tap in the X' pack: phase_t * Q + off_t
  int32_t g_row0;                   // row offset of this pass's phase plane in the G' pack
  int32_t m_tiles, n_tiles, ksplit, kb_per_split, kblocks;
  uint32_t lbo, sbo, kstep;  // MN-major shared-memory descriptor: chunk stride, 8-row group stride, bytes per K = 16 step
};

// MN-major, 128-byte swizzle (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 contiguous
// channels per row, rows 128 bytes apart, 8-row groups SBO = 1024 B apart, 64-channel chunks LBO = one TMA box (8 KiB)
// apart, and a K = 16 step advances the start address by two 8-row groups (2048 B).  Verified on a B200 against the
// exact-fp32 kernel (5e-6); the swapped / other encodings give O(1) errors.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int n) {
  return umma_idesc_bf16(n) | (1u << 15) | (1u << 16);  // A and B MN-major
}

template <int kPasses>
struct WgSmem {
  static constexpr int kStageBytes = (kPasses == 3 ? 2 : 1) * (A_TILE_BYTES + B_TILE_BYTES);
  static constexpr int kStages = (kPasses == 3) ? 2 : 4;
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kTotal = kTileBytes + 256 + 1024;
};

template <int kPasses>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
conv_wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX,
                       const WgradParams wp, float* __restrict__ partial) {
  using S = WgSmem<kPasses>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S::kTileBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles = wp.m_tiles * wp.n_tiles;
  const int items = wp.ksplit * wp.ntaps * tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // item -> (k-chunk, tap, m tile, n tile); K-chunk-major so that concurrent CTAs share operand slices in L2
  auto decode = [&](int item, int& ks, int& t, int& m0, int& n0) {
    const int nt = item % wp.n_tiles;
    int r = item / wp.n_tiles;
    const int mt = r % wp.m_tiles;
    r /= wp.m_tiles;
    t = r % wp.ntaps;
    ks = r / wp.ntaps;
    m0 = mt * GEMM_BLOCK_M;
    n0 = nt * GEMM_BLOCK_N;
  };

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int ks, t, m0, n0;
        decode(item, ks, t, m0, n0);
        const int kb0 = ks * wp.kb_per_split;
        const int kb1 = min(kb0 + wp.kb_per_split, wp.kblocks);
        const int xrow0 = wp.tap_row[t];
        int n_eff = wp.C - n0;
        n_eff = n_eff > GEMM_BLOCK_N ? GEMM_BLOCK_N : ((n_eff + 15) & ~15);
        const int nchunks = (n_eff + 63) >> 6;
        constexpr int kPlanes = kPasses == 3 ? 2 : 1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const int q = kb * GEMM_BLOCK_K;
          mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(kPlanes * (2 + nchunks) * CHUNK_BYTES));
#pragma unroll
          for (int pl = 0; pl < kPlanes; ++pl) {
            const uint32_t a_dst = sa + pl * A_TILE_BYTES;
            const uint32_t b_dst = sa + kPlanes * A_TILE_BYTES + pl * B_TILE_BYTES;
            tma_load_3d(a_dst, &tmG, full_bar(stage), m0, wp.g_row0 + q, pl);
            tma_load_3d(a_dst + CHUNK_BYTES, &tmG, full_bar(stage), m0 + 64, wp.g_row0 + q, pl);
            for (int i = 0; i < nchunks; ++i)
              tma_load_3d(b_dst + i * CHUNK_BYTES, &tmX, full_bar(stage), n0 + 64 * i, xrow0 + q, pl);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++titer) {
        int ks, t, m0, n0;
        decode(item, ks, t, m0, n0);
        int n_eff = wp.C - n0;
        n_eff = n_eff > GEMM_BLOCK_N ? GEMM_BLOCK_N : ((n_eff + 15) & ~15);
        const uint32_t idesc = umma_idesc_bf16_mn(n_eff);
        const int as = titer & 1;
        const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * GEMM_BLOCK_N);
        const int kb0 = ks * wp.kb_per_split;
        const int kb1 = min(kb0 + wp.kb_per_split, wp.kblocks);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t a_hi = sa;
          const uint32_t a_lo = sa + A_TILE_BYTES;
          const uint32_t b_hi = sa + (kPasses == 3 ? 2 : 1) * A_TILE_BYTES;
          const uint32_t b_lo = b_hi + B_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            const uint32_t koff = k * wp.kstep;  // 16 pixel rows = two 8-row groups further down the tile
            const uint64_t da_hi = umma_desc_mn_sw128(a_hi + koff, wp.lbo, wp.sbo);
            const uint64_t db_hi = umma_desc_mn_sw128(b_hi + koff, wp.lbo, wp.sbo);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            if (kPasses == 3) {
              const uint64_t da_lo = umma_desc_mn_sw128(a_lo + koff, wp.lbo, wp.sbo);
              const uint64_t db_lo = umma_desc_mn_sw128(b_lo + koff, wp.lbo, wp.sbo);
              tc_mma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
              tc_mma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
            }
          }
          tc_commit(empty_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull_bar(as));
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    int titer = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++titer) {
      int ks, t, m0, n0;
      decode(item, ks, t, m0, n0);
      int n_eff = wp.C - n0;
      n_eff = n_eff > GEMM_BLOCK_N ? GEMM_BLOCK_N : ((n_eff + 15) & ~15);
      const int as = titer & 1;
      const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
      const int o = m0 + quarter * 32 + lane;
      float* prow = partial + (((int64_t)ks * wp.ntaps + t) * wp.O + o) * wp.Cs + n0;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * GEMM_BLOCK_N);
      for (int c0 = 0; c0 < n_eff; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (o < wp.O) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            if (n0 + c0 + k < wp.Cs)  // Cs is a multiple of 4: whole float4 groups are inside the padded row
              *reinterpret_cast<float4*>(prow + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

struct WTapList {
  int32_t w[SPGAN_MAX_TAPS];
};

// dw[o*ws_o + c*ws_c + tap_w[t]] (+)= scale * sum_s partial[s][t][o][c]; fixed summation order.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(float* __restrict__ dw, const float* __restrict__ partial,
                                                          int ksplit, int ntaps, int O, int C, int Cs, int64_t ws_o,
                                                          int64_t ws_c, WTapList taps, float scale, int accumulate) {
  const int64_t total = (int64_t)ntaps * O * C;
  const int64_t slab = (int64_t)ntaps * O * Cs;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t r = idx / C;
    const int o = (int)(r % O);
    const int t = (int)(r / O);
    const float* src = partial + ((int64_t)t * O + o) * Cs + c;
    float acc = 0.f;
    for (int s = 0; s < ksplit; ++s) acc += __ldg(src + s * slab);
    float* dst = dw + (int64_t)o * ws_o + (int64_t)c * ws_c + taps.w[t];
    *dst = accumulate ? *dst + acc * scale : acc * scale;
  }
}

template <int kPasses>
int launch_wgrad(const CUtensorMap& tmG, const CUtensorMap& tmX, const WgradParams& wp, float* partial, cudaStream_t st) {
  using S = WgSmem<kPasses>;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPGAN_CUDA(cudaGetDevice(&dev), "spgan_conv_wgrad_gemm");
  if (dev < 64 && !attr_set[dev]) {
    SPGAN_CUDA(cudaFuncSetAttribute(conv_wgrad_gemm_kernel<kPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
               "spgan_conv_wgrad_gemm (shared memory opt-in)");
    attr_set[dev] = true;
  }
  const int items = wp.ksplit * wp.ntaps * wp.m_tiles * wp.n_tiles;
  const int grid = items < SPGAN_NUM_SMS ? items : SPGAN_NUM_SMS;
  conv_wgrad_gemm_kernel<kPasses><<<grid, GEMM_THREADS, S::kTotal, st>>>(tmG, tmX, wp, partial);
  SPGAN_CHECK_LAUNCH("spgan_conv_wgrad_gemm");
  return 0;
}

// K split: enough work items for ~2 waves of the 148 SMs, at least 8 k-blocks (512 lattice points) per item.
void plan_split(int Q, int ntaps, int m_tiles, int n_tiles, int* kblocks, int* ksplit, int* kb_per_split) {
  const int kb = (Q + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  const int tiles = ntaps * m_tiles * n_tiles;
  int want = (2 * SPGAN_NUM_SMS + tiles - 1) / tiles;
  int max_split = kb / 8;
  if (max_split < 1) max_split = 1;
  if (want > max_split) want = max_split;
  if (want < 1) want = 1;
  const int per = (kb + want - 1) / want;
  *kblocks = kb;
  *kb_per_split = per;
  *ksplit = (kb + per - 1) / per;
}

}  // namespace

extern "C" int64_t spgan_conv_wgrad_gemm_workspace(const SpganConvPass* p) {
  if (p == nullptr || p->ntaps < 1 || p->Cout < 1 || p->Cin < 1) return 0;
  const int64_t Q = (int64_t)p->B * p->H * p->W;
  if (Q < 1 || Q >= (1LL << 30)) return 0;
  const int m_tiles = (p->Cout + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, n_tiles = (p->Cin + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
  int kblocks, ksplit, per;
  plan_split((int)Q, p->ntaps, m_tiles, n_tiles, &kblocks, &ksplit, &per);
  const int Cs = (p->Cin + 3) / 4 * 4;
  return (int64_t)ksplit * p->ntaps * p->Cout * Cs;
}

extern "C" int spgan_conv_wgrad_gemm(const SpganConvPass* p, float* dw, const uint16_t* g_packed, int g_phases,
                                     int g_phase, int gp_cols, const uint16_t* x_packed, int x_phases,
                                     const int32_t* tap_phase, int xp_cols, float* workspace, int64_t workspace_elems,
                                     int accumulate, void* stream) {
  SPGAN_CHECK_ARG(p != nullptr, "spgan_conv_wgrad_gemm: null pass descriptor");
  SPGAN_CHECK_ARG(p->precision == 1 || p->precision == 2, "spgan_conv_wgrad_gemm: precision must be 1 (bf16x3) or 2 (bf16), got %d",
                  p->precision);
  SPGAN_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= SPGAN_MAX_TAPS, "spgan_conv_wgrad_gemm: %d taps unsupported", p->ntaps);
  SPGAN_CHECK_ARG(p->B >= 0 && p->H >= 0 && p->W >= 0 && p->Cout >= 0 && p->Cin >= 0, "spgan_conv_wgrad_gemm: negative size");
  const int64_t Q = (int64_t)p->B * p->H * p->W;
  if (Q == 0 || p->Cout == 0 || p->Cin == 0) return 0;
  SPGAN_CHECK_ARG(p->Cout >= 16 && p->Cin >= 16, "spgan_conv_wgrad_gemm: Cout=%d / Cin=%d < 16 belong on the SIMT path", p->Cout, p->Cin);
  SPGAN_CHECK_ARG(g_phases >= 1 && g_phase >= 0 && g_phase < g_phases && x_phases >= 1, "spgan_conv_wgrad_gemm: bad phase arguments");
  SPGAN_CHECK_ARG(Q * (g_phases > x_phases ? g_phases : x_phases) < 2147483647LL - 65536, "spgan_conv_wgrad_gemm: too many lattice points");
  SPGAN_CHECK_ARG(gp_cols >= p->Cout && gp_cols % 64 == 0 && xp_cols >= p->Cin && xp_cols % 64 == 0,
                  "spgan_conv_wgrad_gemm: packed widths (%d, %d) must be multiples of 64 covering Cout=%d / Cin=%d", gp_cols, xp_cols,
                  p->Cout, p->Cin);
  SPGAN_CHECK_ARG(dw && g_packed && x_packed && workspace, "spgan_conv_wgrad_gemm: null pointer");
  SPGAN_CHECK_ARG(((((uintptr_t)g_packed) | ((uintptr_t)x_packed) | ((uintptr_t)workspace)) & 15) == 0,
                  "spgan_conv_wgrad_gemm: packed operands and workspace must be 16-byte aligned");
  WgradParams wp;
  wp.Q = (int32_t)Q;
  wp.O = p->Cout;
  wp.C = p->Cin;
  wp.Cs = (p->Cin + 3) / 4 * 4;
  wp.ntaps = p->ntaps;
  wp.g_row0 = 0;  // the G map below is a view of this pass's phase plane only: rows >= Q are zero-filled by TMA
  for (int t = 0; t < SPGAN_MAX_TAPS; ++t) wp.tap_row[t] = 0;
  for (int t = 0; t < p->ntaps; ++t) {
    const int ph = tap_phase ? tap_phase[t] : 0;
    SPGAN_CHECK_ARG(ph >= 0 && ph < x_phases, "spgan_conv_wgrad_gemm: tap %d has phase %d of %d", t, ph, x_phases);
    const int off = p->tap_dy[t] * p->W + p->tap_dx[t];
    SPGAN_CHECK_ARG(off >= 0, "spgan_conv_wgrad_gemm: tap %d has a negative lattice offset (pad the lattice)", t);
    wp.tap_row[t] = (int32_t)(ph * Q + off);
  }
  wp.m_tiles = (p->Cout + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  wp.n_tiles = (p->Cin + GEMM_BLOCK_N - 1) / GEMM_BLOCK_N;
  plan_split((int)Q, p->ntaps, wp.m_tiles, wp.n_tiles, &wp.kblocks, &wp.ksplit, &wp.kb_per_split);
  wp.lbo = CHUNK_BYTES;
  wp.sbo = 1024;
  wp.kstep = 2048;
  const int64_t need = (int64_t)wp.ksplit * wp.ntaps * wp.O * wp.Cs;
  SPGAN_CHECK_ARG(workspace_elems >= need, "spgan_conv_wgrad_gemm: workspace holds %lld floats, %lld needed",
                  (long long)workspace_elems, (long long)need);

  CUtensorMap tmG, tmX;
  {
    // the last K block reads up to 63 rows past Q: they must contribute nothing, so the map ends at this phase plane
    const cuuint64_t rows = (cuuint64_t)g_phases * Q;
    cuuint64_t dims[3] = {(cuuint64_t)gp_cols, (cuuint64_t)Q, 2};
    cuuint64_t strides[2] = {(cuuint64_t)gp_cols * 2, rows * gp_cols * 2};
    cuuint32_t box[3] = {64, GEMM_BLOCK_K, 1};
    const uint16_t* gbase = g_packed + (int64_t)g_phase * Q * gp_cols;
    if (int e = encode_bf16_map(&tmG, gbase, 3, dims, strides, box, "spgan_conv_wgrad_gemm (G map)")) return e;
  }
  {
    const cuuint64_t rows = (cuuint64_t)x_phases * Q;
    cuuint64_t dims[3] = {(cuuint64_t)xp_cols, rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)xp_cols * 2, rows * xp_cols * 2};
    cuuint32_t box[3] = {64, GEMM_BLOCK_K, 1};
    if (int e = encode_bf16_map(&tmX, x_packed, 3, dims, strides, box, "spgan_conv_wgrad_gemm (X map)")) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  int e = p->precision == 1 ? launch_wgrad<3>(tmG, tmX, wp, workspace, st) : launch_wgrad<1>(tmG, tmX, wp, workspace, st);
  if (e) return e;
  WTapList taps;
  for (int t = 0; t < SPGAN_MAX_TAPS; ++t) taps.w[t] = t < p->ntaps ? p->tap_w[t] : 0;
  const int64_t total = (int64_t)p->ntaps * p->Cout * p->Cin;
  wgrad_reduce_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(dw, workspace, wp.ksplit, wp.ntaps, wp.O, wp.C, wp.Cs, p->ws_o,
                                                              p->ws_c, taps, p->out_scale, accumulate);
  SPGAN_CHECK_LAUNCH("spgan_conv_wgrad_gemm (reduce)");
  return 0;
}

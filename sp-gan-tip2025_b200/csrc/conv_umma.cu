// tcgen05 / TMEM implicit-GEMM convolution for sm_100a and its operand packers.
//
//   Y[p, o] = sum_t sum_k A[p + off_t, k] * Wp[t][o][k]          (see include/spgan_b200.h)
//
// A is a channels-last bf16 copy of the (modulated) activation, flattened over (sample, row, column); a conv tap is a
// constant row offset, so every operand tile is one TMA box (128B-swizzled, K-major) and the contraction is issued as
// tcgen05.mma.cta_group::1.kind::f16 (M = 128, N <= 256, K = 16) with the fp32 accumulator in TMEM.
//
// Kernel anatomy (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: one elected lane streams {A_hi, A_lo, B_hi, B_lo} boxes through a ring of smem stages
//   warp 1      TMEM allocator + MMA issuer: one lane issues 4 (K=16 steps) x {hi*hi, hi*lo, lo*hi} MMAs per stage,
//               tcgen05.commit releases the stage / publishes the accumulator
//   warps 2..9  epilogue: tcgen05.ld their 32-lane TMEM quarter (two warps per quarter, alternate 32-column blocks), apply
//               demodulation, noise, bias, leaky-ReLU, residual, scatter to the NCHW fp32 output.  Two accumulator stages
//               (2 x 256 TMEM columns) overlap the epilogue of tile i with the mainloop of tile i+1.
// Roofline: tensor pipe.  FLOPs per launch = 2 * rows * Cout * ntaps * kp (x3 issued MMAs in bf16x3 mode).
#include "gemm_epilogue.cuh"
#include "sphere_taps.cuh"

#include <atomic>

namespace {

constexpr int FWD_THREADS = 320;  // warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue

// ------------------------------------------------------------------------------------------------ GEMM kernel
// kBlockN = 256 for the large layers; 128 halves the tile so that the small layers (structure synthesiser, first texture
// layers: 70..300 M tiles of 128 rows) spread over the 148 SMs without a half-empty last wave.  The MMA rate is the same
// (N = 128 takes half the cycles of N = 256); the A tile is simply fetched once per N tile from L2.
template <int kPasses, int kBlockN>
struct GemmSmem {
  static constexpr int kBTileBytes = kBlockN * GEMM_BLOCK_K * 2;
  static constexpr int kAPlanes = kPasses >= 2 ? 2 : 1;  // 2 passes: (A_hi + A_lo) * B_hi; 3 passes add A_hi * B_lo
  static constexpr int kBPlanes = kPasses == 3 ? 2 : 1;
  static constexpr int kStageBytes = kAPlanes * A_TILE_BYTES + kBPlanes * kBTileBytes;
  static constexpr int kStages = (196608 / kStageBytes) < 4 ? (196608 / kStageBytes) : 4;
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kBarrierBytes + 1024;  // + alignment slack
};

template <int kPasses, int kBlockN>
__global__ void __launch_bounds__(FWD_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmParams gp,
                 const GemmSinks sk) {
  using S = GemmSmem<kPasses, kBlockN>;
  constexpr int kStages = S::kStages;
  constexpr int kBTile = S::kBTileBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S::kTileBytes;
  // barrier slots (8 bytes each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = gp.m_tiles * gp.n_tiles;
  const int kmain = gp.ntaps * gp.kblocks;
  const int kiters = kmain + gp.k2blocks;  // + the blocks of the second K segment (A2 x W2), if any

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
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

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      if (gp.k2blocks > 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / gp.n_tiles) * GEMM_BLOCK_M;
        const int n0 = (tile % gp.n_tiles) * kBlockN;
        // im2col mode: the tile's first output pixel (sample b0, row i0, column j0); TMA walks 128 base pixels from there
        // inside the (My, Mx) bounding box, across rows and samples, and adds the tap offset to each
        int b0 = 0, i0 = 0, j0 = 0;
        if (gp.im2col) {
          const int plane = gp.Hl * gp.Wl;
          b0 = m0 / plane;
          const int r = m0 - b0 * plane;
          i0 = r / gp.Wl;
          j0 = r - i0 * gp.Wl;
        }
        for (int t = 0; t < gp.ntaps; ++t) {
          const int row = m0 + gp.tap_off[t];
          for (int kb = 0; kb < gp.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sa = smem_base + stage * S::kStageBytes;
            mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes);
            const uint32_t sb = sa + S::kAPlanes * A_TILE_BYTES;
            if (gp.im2col) {
              const int img = b0 + gp.tap_img[t];
              tma_load_im2col(sa, &tmA, full_bar(stage), kb * GEMM_BLOCK_K, j0, i0, img, gp.tap_ox[t], gp.tap_oy[t]);
              if (S::kAPlanes == 2)
                tma_load_im2col(sa + A_TILE_BYTES, &tmA, full_bar(stage), kb * GEMM_BLOCK_K, j0, i0, img + gp.img_lo,
                                gp.tap_ox[t], gp.tap_oy[t]);
            } else {
              tma_load_3d(sa, &tmA, full_bar(stage), kb * GEMM_BLOCK_K, row, 0);
              if (S::kAPlanes == 2) tma_load_3d(sa + A_TILE_BYTES, &tmA, full_bar(stage), kb * GEMM_BLOCK_K, row, 1);
            }
            tma_load_4d(sb, &tmB, full_bar(stage), kb * GEMM_BLOCK_K, n0, t, 0);
            if (S::kBPlanes == 2) tma_load_4d(sb + kBTile, &tmB, full_bar(stage), kb * GEMM_BLOCK_K, n0, t, 1);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        // second K segment: rows of A2 are the tile's M rows themselves (no tap displacement), one weight slab
        for (int kb = 0; kb < gp.k2blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          mbar_arrive_expect_tx(full_bar(stage), S::kStageBytes);
          const uint32_t sb = sa + S::kAPlanes * A_TILE_BYTES;
          tma_load_3d(sa, &tmA2, full_bar(stage), kb * GEMM_BLOCK_K, m0, 0);
          if (S::kAPlanes == 2) tma_load_3d(sa + A_TILE_BYTES, &tmA2, full_bar(stage), kb * GEMM_BLOCK_K, m0, 1);
          tma_load_4d(sb, &tmB2, full_bar(stage), kb * GEMM_BLOCK_K, n0, 0, 0);
          if (S::kBPlanes == 2) tma_load_4d(sb + kBTile, &tmB2, full_bar(stage), kb * GEMM_BLOCK_K, n0, 0, 1);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
        const int n0 = (tile % gp.n_tiles) * kBlockN;
        int n_eff = gp.Cout - n0;
        n_eff = n_eff > kBlockN ? kBlockN : ((n_eff + 15) & ~15);
        const uint32_t idesc = umma_idesc_16(n_eff, gp.a_f16 != 0, gp.b_f16 != 0);
        const int as = titer & 1;
        const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kBlockN);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t a_hi = sa;
          const uint32_t a_lo = sa + A_TILE_BYTES;
          const uint32_t b_hi = sa + S::kAPlanes * A_TILE_BYTES;
          const uint32_t b_lo = b_hi + kBTile;
          // the last K block of a tap may be partial (kp a multiple of 16, not of 64): TMA zero-fills the box, the MMAs
          // of the all-zero K steps are simply not issued
          const int ksteps = (it < kmain && (it + 1) % gp.kblocks == 0) ? gp.last_ksteps : GEMM_BLOCK_K / GEMM_UMMA_K;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            if (k >= ksteps) break;
            const uint32_t koff = k * GEMM_UMMA_K * 2;  // bytes inside the 128-byte swizzle row
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff);
            const uint64_t db_hi = umma_desc_sw128(b_hi + koff);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (kPasses == 3) tc_mma_f16(d_tmem, da_hi, umma_desc_sw128(b_lo + koff), idesc, 1u);
            if (kPasses >= 2) tc_mma_f16(d_tmem, umma_desc_sw128(a_lo + koff), db_hi, idesc, 1u);
          }
          tc_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc_commit(tfull_bar(as));  // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (warps 2..9)
    // Eight warps: warp w may touch TMEM lanes 32*(w & 3)..+31; the two warps of a lane quarter take alternate blocks of 32
    // accumulator columns.  With one warp per scheduler the epilogue of the few-tap passes (parity passes of the
    // transposed conv, 1x1 shortcuts) was issue-latency bound at ~56 instructions per column; here the per-column work is
    // a multiply, a pointer bump and a store, the per-channel factors come in as 128-bit loads, and the rare terms (noise,
    // bias, activation, residual) are behind one warp-uniform branch.
    int titer = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
      const int as = titer & 1;
      const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
      gemm_epilogue_tile<kBlockN>(gp, sk, (tile / gp.n_tiles) * GEMM_BLOCK_M, tile % gp.n_tiles, warp, lane, tfull_bar(as), aphase, tmem_base + (uint32_t)(as * kBlockN));
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

// ------------------------------------------------------------------------------------------------ CTA-pair GEMM kernel
// cta_group::2: the two CTAs of a cluster (one TPC) compute ONE 256 x 256 tile.  CTA r holds rows [128 r, 128 r + 128) of the
// A tile and rows [128 r, 128 r + 128) of the 256-row weight tile; the leader (rank 0) issues tcgen05.mma.cta_group::2 with
// M = 256, which reads both CTAs' shared memory and writes both CTAs' TMEM.  Per CTA and K block that is 32 KiB of weights less
// than the single-CTA kernel fetches from L2 (a third of the bf16x3 stage), so the ring holds 3 / 4 / 6 stages instead of
// 2 / 3 / 4 and the operand traffic — which is what the power-limited part pays for — drops by a third.
//   both CTAs   warp 0: TMA for its own A rows and its own half of the weight tile, completing on the LEADER's full barrier
//   leader      warp 1: waits the full barrier (bytes of both CTAs), issues the MMAs, commits (multicast) to both CTAs' empty /
//                       accumulator-full barriers
//   both CTAs   warps 2..9: epilogue of their own 128 accumulator rows; accumulator-empty arrives go to the leader's barrier
template <int kPasses>
struct Gemm2Smem {
  static constexpr int kBHalfBytes = 128 * GEMM_BLOCK_K * 2;  // this CTA's half of the 256-row weight tile
  static constexpr int kAPlanes = kPasses >= 2 ? 2 : 1;
  static constexpr int kBPlanes = kPasses == 3 ? 2 : 1;
  static constexpr int kStageBytes = kAPlanes * A_TILE_BYTES + kBPlanes * kBHalfBytes;
  static constexpr int kStages = (196608 / kStageBytes) < 6 ? (196608 / kStageBytes) : 6;
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kBarrierBytes + 1024;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma2_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n,
                                                 int w_off, int h_off) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)w_off), "h"((uint16_t)h_off)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc2_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kPasses>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FWD_THREADS, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const GemmParams gp,
                  const GemmSinks sk) {
  using S = Gemm2Smem<kPasses>;
  constexpr int kStages = S::kStages;
  constexpr int kBHalf = S::kBHalfBytes;
  constexpr int kPairN = 256;
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
  const uint32_t rank = cluster_rank();
  const bool leader = rank == 0;
  const int m_pairs = (gp.m_tiles + 1) / 2;
  const int n_tiles = gp.Cout / kPairN;  // the host launches this kernel only for Cout % 256 == 0
  const int num_tiles = m_pairs * n_tiles;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int kmain = gp.ntaps * gp.kblocks;
  const int kiters = kmain + gp.k2blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // leader: its own arrive.expect_tx; the bytes of both CTAs complete on it
      mbar_init(empty_bar(s), 1);  // one multicast commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 16);  // leader: 8 epilogue warps of each CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      if (gp.k2blocks > 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cid; tile < num_tiles; tile += ncl) {
        const int m0 = (tile / n_tiles) * (2 * GEMM_BLOCK_M) + (int)rank * GEMM_BLOCK_M;
        const int n0 = (tile % n_tiles) * kPairN + (int)rank * 128;  // this CTA's half of the weight rows
        int b0 = 0, i0 = 0, j0 = 0;
        if (gp.im2col) {
          const int plane = gp.Hl * gp.Wl;
          b0 = m0 / plane;
          const int r = m0 - b0 * plane;
          i0 = r / gp.Wl;
          j0 = r - i0 * gp.Wl;
        }
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t sb = sa + S::kAPlanes * A_TILE_BYTES;
          const uint32_t fb = map_to_cta(full_bar(stage), 0);
          if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * S::kStageBytes);
          if (it < kmain) {
            const int t = it / gp.kblocks, kb = it - t * gp.kblocks;
            if (gp.im2col) {
              const int img = b0 + gp.tap_img[t];
              tma2_load_im2col(sa, &tmA, fb, kb * GEMM_BLOCK_K, j0, i0, img, gp.tap_ox[t], gp.tap_oy[t]);
              if (S::kAPlanes == 2)
                tma2_load_im2col(sa + A_TILE_BYTES, &tmA, fb, kb * GEMM_BLOCK_K, j0, i0, img + gp.img_lo, gp.tap_ox[t], gp.tap_oy[t]);
            } else {
              const int row = m0 + gp.tap_off[t];
              tma2_load_3d(sa, &tmA, fb, kb * GEMM_BLOCK_K, row, 0);
              if (S::kAPlanes == 2) tma2_load_3d(sa + A_TILE_BYTES, &tmA, fb, kb * GEMM_BLOCK_K, row, 1);
            }
            tma2_load_4d(sb, &tmB, fb, kb * GEMM_BLOCK_K, n0, t, 0);
            if (S::kBPlanes == 2) tma2_load_4d(sb + kBHalf, &tmB, fb, kb * GEMM_BLOCK_K, n0, t, 1);
          } else {
            const int kb = it - kmain;
            tma2_load_3d(sa, &tmA2, fb, kb * GEMM_BLOCK_K, m0, 0);
            if (S::kAPlanes == 2) tma2_load_3d(sa + A_TILE_BYTES, &tmA2, fb, kb * GEMM_BLOCK_K, m0, 1);
            tma2_load_4d(sb, &tmB2, fb, kb * GEMM_BLOCK_K, n0, 0, 0);
            if (S::kBPlanes == 2) tma2_load_4d(sb + kBHalf, &tmB2, fb, kb * GEMM_BLOCK_K, n0, 0, 1);
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
    // ===================================================================== MMA issuer (leader only)
    if (leader && lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      // M = 256 (both CTAs' 128 rows), N = 256
      const uint32_t idesc = (1u << 4) | ((gp.a_f16 ? 0u : 1u) << 7) | ((gp.b_f16 ? 0u : 1u) << 10) | ((uint32_t)(kPairN >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      for (int tile = cid; tile < num_tiles; tile += ncl, ++titer) {
        const int as = titer & 1;
        const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * kPairN);
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * S::kStageBytes;
          const uint32_t a_hi = sa;
          const uint32_t a_lo = sa + A_TILE_BYTES;
          const uint32_t b_hi = sa + S::kAPlanes * A_TILE_BYTES;
          const uint32_t b_lo = b_hi + kBHalf;
          const int ksteps = (it < kmain && (it + 1) % gp.kblocks == 0) ? gp.last_ksteps : GEMM_BLOCK_K / GEMM_UMMA_K;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            if (k >= ksteps) break;
            const uint32_t koff = k * GEMM_UMMA_K * 2;
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff);
            const uint64_t db_hi = umma_desc_sw128(b_hi + koff);
            tc2_mma_f16(d_tmem, da_hi, db_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (kPasses == 3) tc2_mma_f16(d_tmem, da_hi, umma_desc_sw128(b_lo + koff), idesc, 1u);
            if (kPasses >= 2) tc2_mma_f16(d_tmem, umma_desc_sw128(a_lo + koff), db_hi, idesc, 1u);
          }
          tc2_commit(empty_bar(stage));  // both CTAs' stage `stage` is free once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        tc2_commit(tfull_bar(as));  // both CTAs' accumulator halves are complete
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (both CTAs, their own 128 rows)
    int titer = 0;
    for (int tile = cid; tile < num_tiles; tile += ncl, ++titer) {
      const int as = titer & 1;
      const uint32_t aphase = (uint32_t)(titer >> 1) & 1u;
      const int m0 = (tile / n_tiles) * (2 * GEMM_BLOCK_M) + (int)rank * GEMM_BLOCK_M;
      gemm_epilogue_tile<kPairN>(gp, sk, m0, tile % n_tiles, warp, lane, tfull_bar(as), aphase, tmem_base + (uint32_t)(as * kPairN));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty_bar(as));
        else mbar_arrive_cluster(map_to_cta(tempty_bar(as), 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it or read its shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ operand packers
// (v0, v1) -> packed 16-bit hi pair and lo pair (lo = round(v - hi)); same values as two split16 calls
template <bool kF16>
__device__ __forceinline__ void split_pair_k(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  if (kF16) {
    v0 = fminf(fmaxf(v0, -65504.f), 65504.f);
    v1 = fminf(fmaxf(v1, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
    const float h0 = __uint_as_float(hb << 16), h1 = __uint_as_float(hb & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - h0, v1 - h1);
    hi = hb;
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// One CTA: 64 channels x 64 consecutive lattice points of one sample (and one polyphase plane).  Reads coalesced along
// pixels, writes coalesced along channels (8 channels = 16 bytes per lane and plane).  With step s > 1 the lattice point
// (i, j) of phase (py, px) holds source pixel (i*s + py - pad_y, j*s + px - pad_x): a strided conv becomes a stride-1
// conv whose taps pick their phase plane (rows [ph * B*Hl*Wl, (ph+1) * B*Hl*Wl) of the packed matrix).
template <bool kF16>
__global__ void __launch_bounds__(256) pack_act_kernel(uint16_t* __restrict__ out, const float* __restrict__ x,
                                                      const float* __restrict__ in_mul, int B, int C, int H, int W,
                                                      int Cp, int pad_y, int pad_x, int Hl, int Wl, int step) {
  // (pad_y, pad_x) = top/left offset of the image inside the (Hl, Wl) lattice
  __shared__ float tile[64][65];
  const int plane_l = Hl * Wl;
  const int q0 = blockIdx.x * 64;  // lattice point within the sample
  const int c0 = blockIdx.y * 64;
  const int ph = blockIdx.z / B;
  const int b = blockIdx.z - ph * B;
  const int py = ph / step, px = ph - py * step;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  {
    const int q = q0 + tx;
    int sy = -1, sx = -1;
    if (q < plane_l) {
      const int i = q / Wl, j = q - i * Wl;
      sy = i * step + py - pad_y;
      sx = j * step + px - pad_x;
    }
    const bool inside = sy >= 0 && sy < H && sx >= 0 && sx < W;
    // all 16 loads of a thread are in flight before the first shared-memory store (bytes in flight cover the latency)
    float v[16], m[16];
    const int64_t cstep = (int64_t)H * W;
    const float* xp = x + ((int64_t)b * C + c0 + ty) * cstep + (int64_t)sy * W + sx;  // one pointer, bumped per channel
    const float* mp = in_mul ? in_mul + (int64_t)b * C + c0 + ty : nullptr;
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const bool ok = inside && c0 + ty + 4 * u < C;
      v[u] = ok ? __ldcs(xp + 4 * u * cstep) : 0.f;
      m[u] = (ok && mp) ? __ldg(mp + 4 * u) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) tile[ty + 4 * u][tx] = v[u] * m[u];
  }
  __syncthreads();
  // write phase: one item = (lattice point, run of 8 channels) -> one 16-byte store per plane (a warp covers 4 points x 128
  // contiguous bytes); pairwise 16-bit conversions.  (One bf16x2 per lane and eight 2-byte-pair splits per thread made this
  // phase the issue bottleneck: 0.65 of the HBM roofline.)
  const int64_t rows_total = (int64_t)step * step * B * plane_l;
  uint16_t* out_lo = out + rows_total * Cp;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int item = threadIdx.x + 256 * k;
    const int pp = item >> 3, chunk = item & 7;
    const int q = q0 + pp;
    const int cbase = c0 + 8 * chunk;
    if (q >= plane_l || cbase >= Cp) continue;  // Cp is a multiple of 16: whole 8-channel runs
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = tile[8 * chunk + 2 * j][pp], v1 = tile[8 * chunk + 2 * j + 1][pp];
      split_pair_k<kF16>(v0, v1, h[j], l[j]);
    }
    const int64_t off = (((int64_t)ph * B + b) * plane_l + q) * Cp + cbase;
    *reinterpret_cast<uint4*>(out + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(out_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

struct TapList {
  int32_t w[SPGAN_MAX_TAPS];
};

// out[plane][t][o][c] (or [plane][o][t*Cp + c] when merged); one thread per (t, o, c) element, c fastest.
template <bool kF16>
__global__ void __launch_bounds__(256) pack_weight_kernel(uint16_t* __restrict__ out, const float* __restrict__ w,
                                                         int Cout, int Cin, int64_t ws_o, int64_t ws_c, int ntaps,
                                                         TapList taps, int Cp, int merged) {
  const int64_t total = (int64_t)ntaps * Cout * Cp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cp);
    const int64_t r = idx / Cp;
    const int o = (int)(r % Cout);
    const int t = (int)(r / Cout);
    float v = 0.f;
    if (c < Cin) v = __ldg(w + (int64_t)o * ws_o + (int64_t)c * ws_c + taps.w[t]);
    uint16_t hi, lo;
    split16<kF16>(v, hi, lo);
    const int64_t dst = merged ? ((int64_t)o * ntaps + t) * Cp + c : idx;
    out[dst] = hi;
    out[total + dst] = lo;
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(float* __restrict__ out, const float* __restrict__ x, int C,
                                                          int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int cc = ty; cc < 32; cc += 8) {
    const int c = c0 + cc, p = p0 + tx;
    tile[cc][tx] = (c < C && p < HW) ? __ldg(x + ((int64_t)b * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const int p = p0 + pp, c = c0 + tx;
    if (p < HW && c < C) out[((int64_t)b * HW + p) * C + c] = tile[tx][pp];
  }
}

// One warp per (group g, pixel p, tap t); lanes sweep the channel pairs (2*lane, 2*lane + 1) + 64*j.  Which gathered
// plane feeds channel k of group g comes from the host-built table chan_map[g][k] (the reference's flat-concat
// quirk lives there), so the inner loop has no integer divisions.  The corner indices and weights are computed once
// per source sample; corner reads are coalesced over channels in the NHWC staging copy and each lane writes one bf16x2
// to the hi plane and one to the lo plane (128 B per warp and plane).
template <bool kF16>
__global__ void __launch_bounds__(256) sphere_pack_kernel(uint16_t* __restrict__ out, const float* __restrict__ xh,
                                                         const float* __restrict__ coords, const float* __restrict__ grid,
                                                         const float* __restrict__ in_mul,
                                                         const uint32_t* __restrict__ chan_map, int B, int C, int nc,
                                                         int H, int W, int grid_batch, int Cp) {
  const int Ct = C + nc;
  const int HW = H * W;
  const int64_t plane_elems = (int64_t)B * HW * 9 * Cp;
  const int64_t warps_total = (int64_t)B * HW * 9;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warp_stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wid = warp0; wid < warps_total; wid += warp_stride) {
    const int t = (int)(wid % 9);
    const int64_t r = wid / 9;
    const int p = (int)(r % HW);
    const int g = (int)(r / HW);
    const int py = p / W, px = p - py * W;
    const int ty = t / 3, tx = t - ty * 3;
    int cur_bs = -1;
    TapCorners cn;
    cn.o_nw = cn.o_ne = cn.o_sw = cn.o_se = 0;
    cn.w_nw = cn.w_ne = cn.w_sw = cn.w_se = 0.f;
    uint16_t* orow = out + wid * Cp;  // row (g, p), columns [t*Cp, (t+1)*Cp): wid = (g*HW + p)*9 + t
    const uint32_t* mrow = chan_map + (int64_t)g * Cp;
    const float* mulrow = in_mul ? in_mul + (int64_t)g * Ct : nullptr;
    for (int k0 = 2 * lane; k0 < Cp; k0 += 64) {
      const uint2 mm = __ldg(reinterpret_cast<const uint2*>(mrow + k0));
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t m = u ? mm.y : mm.x;
        float val = 0.f;
        if (m != 0xFFFFFFFFu) {
          const int bs = (int)((m >> 15) & 0xFFFFu);
          const int cs = (int)(m & 0x7FFFu);
          const bool is_coord = (m >> 31) != 0;
          if (bs != cur_bs) {
            cur_bs = bs;
            cn = tap_corners(grid, grid_batch == 1 ? 0 : bs, H, W, py, px, ty, tx);
          }
          float a, b2, c2, d2;
          if (!is_coord) {
            const float* src = xh + (int64_t)bs * HW * C + cs;
            a = __ldg(src + (int64_t)cn.o_nw * C);
            b2 = __ldg(src + (int64_t)cn.o_ne * C);
            c2 = __ldg(src + (int64_t)cn.o_sw * C);
            d2 = __ldg(src + (int64_t)cn.o_se * C);
          } else {
            const float* src = coords + ((int64_t)bs * nc + cs) * HW;
            a = __ldg(src + cn.o_nw);
            b2 = __ldg(src + cn.o_ne);
            c2 = __ldg(src + cn.o_sw);
            d2 = __ldg(src + cn.o_se);
          }
          val = a * cn.w_nw + b2 * cn.w_ne + c2 * cn.w_sw + d2 * cn.w_se;
          if (is_coord) {
            if (cs == 0) val = tanhf(val);
            else if (cs == 1) val = cosf(val * 3.14159274101257324f);
            else if (cs == 2) val = sinf(val * 3.14159274101257324f);
          }
          if (mulrow) val *= __ldg(mulrow + k0 + u);
        }
        v[u] = val;
      }
      uint16_t h0, l0, h1, l1;
      split16<kF16>(v[0], h0, l0);
      split16<kF16>(v[1], h1, l1);
      *reinterpret_cast<uint32_t*>(orow + k0) = pack2x16(h0, h1);
      *reinterpret_cast<uint32_t*>(orow + plane_elems + k0) = pack2x16(l0, l1);
    }
  }
}

// Shared-grid specialisation (test / panorama mode: one sampling grid for the whole batch, grid_batch == 1) with
// Cp = 64 * KITER known at compile time.  One warp per (group g, pixel p) walks all 9 taps: lanes 0..8 fetch the nine
// grid entries and compute the corner sets in parallel (one grid latency per pixel instead of one per tap, broadcast by
// shuffle), the channel map and the modulation are decoded once per pixel, and per tap all 8 * KITER corner loads of a
// lane are issued before the first use.  The general kernel above pays three dependent global latencies per
// (pixel, tap) with 16 warps per SM and sat at ~15 % of the HBM write roofline.
template <int KITER, bool kF16>
__global__ void __launch_bounds__(256, 2) sphere_pack_shared_kernel(uint16_t* __restrict__ out,
                                                                const float* __restrict__ xh,
                                                                const float* __restrict__ coords,
                                                                const float* __restrict__ grid,
                                                                const float* __restrict__ in_mul,
                                                                const uint32_t* __restrict__ chan_map, int B, int C,
                                                                int nc, int H, int W) {
  constexpr int Cp = 64 * KITER;
  const int Ct = C + nc;
  const int HW = H * W;
  const int64_t plane_elems = (int64_t)B * HW * 9 * Cp;
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
    if (lane < 9) mine = tap_corners(grid, 0, H, W, py, px, lane / 3, lane % 3);
    const uint32_t* mrow = chan_map + (int64_t)g * Cp;
    const float* mulrow = in_mul ? in_mul + (int64_t)g * Ct : nullptr;
    // decode this lane's 2 * KITER channels once: 32-bit element offset of the source plane, the map word (flags) and
    // the modulation
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
    uint16_t* obase = out + wid * 9 * Cp;  // row (g, p): 9 taps x Cp columns
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
      uint16_t* orow = obase + t * Cp;
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
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
std::atomic<long long>* launch_counter() {
  static std::atomic<long long> c{0};
  return &c;
}

template <int kPasses, int kBlockN>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                const GemmParams& gp, const GemmSinks& sk, cudaStream_t st) {
  using S = GemmSmem<kPasses, kBlockN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPGAN_CUDA(cudaGetDevice(&dev), "spgan_conv_gemm");
  if (dev < 64 && !attr_set[dev]) {
    SPGAN_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<kPasses, kBlockN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
               "spgan_conv_gemm (shared memory opt-in)");
    attr_set[dev] = true;
  }
  const int tiles = gp.m_tiles * gp.n_tiles;
  const int grid = tiles < SPGAN_NUM_SMS ? tiles : SPGAN_NUM_SMS;
  conv_gemm_kernel<kPasses, kBlockN><<<grid, FWD_THREADS, S::kTotal, st>>>(tmA, tmB, tmA2, tmB2, gp, sk);
  SPGAN_CHECK_LAUNCH("spgan_conv_gemm");
  launch_counter()->fetch_add(1);
  return 0;
}

template <int kPasses>
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA2, const CUtensorMap& tmB2,
                 const GemmParams& gp, const GemmSinks& sk, cudaStream_t st) {
  using S = Gemm2Smem<kPasses>;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPGAN_CUDA(cudaGetDevice(&dev), "spgan_conv_gemm");
  if (dev < 64 && !attr_set[dev]) {
    SPGAN_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<kPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
               "spgan_conv_gemm (shared memory opt-in, CTA-pair kernel)");
    attr_set[dev] = true;
  }
  const int ptiles = ((gp.m_tiles + 1) / 2) * (gp.Cout / 256);
  const int clusters = ptiles < SPGAN_NUM_SMS / 2 ? ptiles : SPGAN_NUM_SMS / 2;
  conv_gemm2_kernel<kPasses><<<2 * clusters, FWD_THREADS, S::kTotal, st>>>(tmA, tmB, tmA2, tmB2, gp, sk);
  SPGAN_CHECK_LAUNCH("spgan_conv_gemm (CTA-pair kernel)");
  launch_counter()->fetch_add(1);
  return 0;
}

// 0 = never use the CTA-pair kernel, 1 = where its tiling fills the machine at least as well (default), 2 = wherever legal
std::atomic<int>* pair_mode() {
  static std::atomic<int> m{1};
  return &m;
}

// N tile choice.  N = 128 costs ~25 % more time per FLOP than N = 256 (measured on the 55x55 layer: the A tile is fetched
// twice as often and the chip is power-limited), so it is used only where N = 256 leaves the machine badly under-filled:
// at most 1.5 waves of tiles and a wave efficiency gain of more than a third.
int pick_block_n(int64_t m_tiles, int cout) {
  if (cout <= 128) return 128;
  auto eff = [&](int bn) {
    const int64_t tiles = m_tiles * ((cout + bn - 1) / bn);
    const int64_t waves = (tiles + SPGAN_NUM_SMS - 1) / SPGAN_NUM_SMS;
    return (double)tiles / (double)(waves * SPGAN_NUM_SMS);
  };
  const int64_t tiles256 = m_tiles * ((cout + 255) / 256);
  return (2 * tiles256 <= 3 * SPGAN_NUM_SMS && eff(128) > 1.33 * eff(256)) ? 128 : 256;
}

// M extent and N tile of a pass: shared by the launcher and by spgan_conv_gemm_rgb_slots.
struct GemmShape {
  bool im2col;
  bool pair;  // cta_group::2 kernel: 256 x 256 tiles over CTA pairs
  int phases;
  int64_t rows_m;
  int m_tiles, block_n, n_tiles;
};

GemmShape gemm_shape(const SpganConvPass* p, int64_t a_rows, int kblocks) {
  GemmShape g;
  const int64_t rows = (int64_t)p->B * p->H * p->W;
  g.phases = rows > 0 ? (int)(a_rows / rows) : 1;
  // im2col A loads whenever the pass has lattice points that are not outputs (unpadded 3x3 / 7x7 convs, padded convs on
  // their bordered lattice, parity passes): M then runs over the B*My*Mx outputs only.  The bounding-box corner of a
  // rank-4 map is an 8-bit field, hence the 128 limit; passes outside it keep the flat row-offset loads.
  g.im2col = (p->My < p->H || p->Mx < p->W) && p->My <= p->H && p->Mx <= p->W && p->H - p->My <= 128 &&
             p->W - p->Mx <= 128 && (int64_t)2 * g.phases * p->B < (1LL << 31);
  g.rows_m = g.im2col ? (int64_t)p->B * p->My * p->Mx : rows;
  g.m_tiles = (int)((g.rows_m + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  g.block_n = pick_block_n(g.m_tiles, p->Cout);
  g.n_tiles = (p->Cout + g.block_n - 1) / g.block_n;
  // CTA-pair kernel (cta_group::2).  Measured per shape on the B200 (tools/probes/pair_ab.py, bf16x3): 8-16 % faster from
  // ~130 M tiles of 128 rows upwards (less weight traffic per SM, deeper ring), 15-30 % SLOWER below ~70 M tiles (half as many
  // work items, cluster launch + two cluster barriers per launch), neutral for 1x1 convs (4 K blocks: epilogue-bound).
  g.pair = false;
  const int mode = pair_mode()->load();
  if (mode != 0 && p->Cout % 256 == 0 && g.m_tiles >= 2) {
    g.pair = mode == 2 || (g.m_tiles >= 128 && p->ntaps * kblocks >= 8);
    if (g.pair) {
      g.block_n = 256;
      g.n_tiles = p->Cout / 256;
    }
  }
  return g;
}

}  // namespace

extern "C" int64_t spgan_gemm_launch_count(void) { return (int64_t)launch_counter()->load(); }

extern "C" int spgan_set_option(int key, int value) {
  SPGAN_CHECK_ARG(key == 1, "spgan_set_option: unknown key %d", key);
  SPGAN_CHECK_ARG(value >= 0 && value <= 2, "spgan_set_option: CTA-pair mode must be 0 (off), 1 (auto) or 2 (wherever legal), got %d", value);
  pair_mode()->store(value);
  return 0;
}
void spgan_internal_count_gemm_launch() { launch_counter()->fetch_add(1); }

extern "C" int spgan_pack_act(uint16_t* out, const float* x, const float* in_mul, int B, int C, int H, int W, int Cp,
                              int pad_y, int pad_x, int Hl, int Wl, int step, int fmt, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0 && pad_y >= 0 && pad_x >= 0, "spgan_pack_act: negative size");
  SPGAN_CHECK_ARG(step >= 1 && step <= 8, "spgan_pack_act: step %d unsupported", step);
  SPGAN_CHECK_ARG(step > 1 || (Hl >= H + pad_y && Wl >= W + pad_x), "spgan_pack_act: lattice %dx%d smaller than the padded image", Hl, Wl);
  SPGAN_CHECK_ARG(Cp >= C && Cp % 16 == 0, "spgan_pack_act: Cp=%d must be a multiple of 16 and >= C=%d", Cp, C);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_pack_act: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  if (B == 0 || Cp == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && x, "spgan_pack_act: null pointer");
  SPGAN_CHECK_ARG((((uintptr_t)out) & 15) == 0, "spgan_pack_act: the packed operand must be 16-byte aligned");
  SPGAN_CHECK_ARG(B * step * step <= 65535, "spgan_pack_act: batch %d x %d phases > 65535", B, step * step);
  dim3 grid((Hl * Wl + 63) / 64, (Cp + 63) / 64, B * step * step);
  if (fmt)
    pack_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, in_mul, B, C, H, W, Cp, pad_y, pad_x, Hl, Wl, step);
  else
    pack_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, in_mul, B, C, H, W, Cp, pad_y, pad_x, Hl, Wl, step);
  SPGAN_CHECK_LAUNCH("spgan_pack_act");
  return 0;
}

extern "C" int spgan_pack_weight(uint16_t* out, const float* w, int Cout, int Cin, int64_t ws_o, int64_t ws_c, int ntaps,
                                 const int32_t* tap_w, int Cp, int merged, int fmt, void* stream) {
  SPGAN_CHECK_ARG(Cout >= 0 && Cin >= 0, "spgan_pack_weight: negative size");
  SPGAN_CHECK_ARG(ntaps >= 1 && ntaps <= SPGAN_MAX_TAPS, "spgan_pack_weight: %d taps unsupported", ntaps);
  SPGAN_CHECK_ARG(Cp >= Cin && Cp % 16 == 0, "spgan_pack_weight: Cp=%d must be a multiple of 16 and >= Cin=%d", Cp, Cin);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_pack_weight: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  if (Cout == 0 || Cp == 0) return 0;
  SPGAN_CHECK_ARG(out && w && tap_w, "spgan_pack_weight: null pointer");
  TapList taps;
  for (int t = 0; t < ntaps; ++t) taps.w[t] = tap_w[t];
  const int64_t total = (int64_t)ntaps * Cout * Cp;
  if (fmt)
    pack_weight_kernel<true><<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, w, Cout, Cin, ws_o, ws_c, ntaps, taps, Cp, merged);
  else
    pack_weight_kernel<false><<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, w, Cout, Cin, ws_o, ws_c, ntaps, taps, Cp, merged);
  SPGAN_CHECK_LAUNCH("spgan_pack_weight");
  return 0;
}

extern "C" int spgan_nchw_to_nhwc(float* out, const float* x, int B, int C, int H, int W, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "spgan_nchw_to_nhwc: negative size");
  if (B == 0 || C == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && x, "spgan_nchw_to_nhwc: null pointer");
  SPGAN_CHECK_ARG(B <= 65535, "spgan_nchw_to_nhwc: batch %d > 65535", B);
  dim3 grid((H * W + 31) / 32, (C + 31) / 32, B);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, C, H * W);
  SPGAN_CHECK_LAUNCH("spgan_nchw_to_nhwc");
  return 0;
}

extern "C" int spgan_sphere_pack(uint16_t* out, const float* x_nhwc, const float* coords, const float* grid,
                                 const float* in_mul, const uint32_t* chan_map, int B, int C, int H, int W,
                                 int grid_batch, int Cp, int fmt, void* stream) {
  SPGAN_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "spgan_sphere_pack: negative size");
  const int nc = coords ? 3 : 0;
  SPGAN_CHECK_ARG(Cp >= C + nc && Cp % 64 == 0, "spgan_sphere_pack: Cp=%d must be a multiple of 64 and >= %d", Cp, C + nc);
  SPGAN_CHECK_ARG(fmt == 0 || fmt == 1, "spgan_sphere_pack: fmt must be 0 (bf16 hi/lo) or 1 (fp16 hi/lo), got %d", fmt);
  if (B == 0 || H == 0 || W == 0) return 0;
  SPGAN_CHECK_ARG(out && x_nhwc && grid && chan_map, "spgan_sphere_pack: null pointer");
  SPGAN_CHECK_ARG(B <= 65535 && C <= 32767, "spgan_sphere_pack: B=%d / C=%d exceed the channel-map encoding", B, C);
  SPGAN_CHECK_ARG(grid_batch == 1 || grid_batch == B, "spgan_sphere_pack: grid batch %d must be 1 or %d", grid_batch, B);
  SPGAN_CHECK_ARG((int64_t)B * H * W * (C > 3 ? C : 3) < (1LL << 31), "spgan_sphere_pack: input too large for 32-bit plane offsets");
  SPGAN_CHECK_ARG((((uintptr_t)grid) & 7) == 0 && (((uintptr_t)chan_map) & 7) == 0,
                  "spgan_sphere_pack: grid and chan_map must be 8-byte aligned");
  const int64_t warps = (int64_t)B * H * W * 9;
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = grid_for((int64_t)B * H * W, 8, 2, 16);
#define SPGAN_SPHERE_SHARED(KI)                                                                                        \
  do {                                                                                                                 \
    if (fmt) sphere_pack_shared_kernel<KI, true><<<nblk, 256, 0, st>>>(out, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W); \
    else sphere_pack_shared_kernel<KI, false><<<nblk, 256, 0, st>>>(out, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W);    \
  } while (0)
  if (grid_batch == 1 && Cp == 64) SPGAN_SPHERE_SHARED(1);
  else if (grid_batch == 1 && Cp == 128) SPGAN_SPHERE_SHARED(2);
  else if (grid_batch == 1 && Cp == 192) SPGAN_SPHERE_SHARED(3);
  else if (grid_batch == 1 && Cp == 256) SPGAN_SPHERE_SHARED(4);
  else if (grid_batch == 1 && Cp == 320) SPGAN_SPHERE_SHARED(5);
  else if (fmt)
    sphere_pack_kernel<true><<<grid_for(warps, 8, 8, 8), 256, 0, st>>>(out, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W, grid_batch, Cp);
  else
    sphere_pack_kernel<false><<<grid_for(warps, 8, 8, 8), 256, 0, st>>>(out, x_nhwc, coords, grid, in_mul, chan_map, B, C, nc, H, W, grid_batch, Cp);
#undef SPGAN_SPHERE_SHARED
  SPGAN_CHECK_LAUNCH("spgan_sphere_pack");
  return 0;
}

extern "C" int spgan_conv_gemm_rgb_slots(const SpganConvPass* p, int64_t a_rows) {
  if (p == nullptr || p->Cout <= 0) return 0;
  return 2 * gemm_shape(p, a_rows, 8).n_tiles;  // the ToRGB sink rides on 512-channel convs: kp >= 512
}

extern "C" int spgan_conv_gemm_ex(const SpganConvPass* p, const SpganGemmIO* io, void* stream) {
  SPGAN_CHECK_ARG(p != nullptr && io != nullptr, "spgan_conv_gemm: null descriptor");
  SPGAN_CHECK_ARG(p->precision >= 1 && p->precision <= 3,
                  "spgan_conv_gemm: precision must be 1 (bf16x3), 2 (bf16) or 3 (fp16x2), got %d", p->precision);
  // both operands of one kind::f16 MMA must share a format: a bf16 A with an fp16 B raises an illegal-instruction fault
  // on the B200 (tools/probes/mixed_fmt.py), so there is no "bf16 hi/lo activations x fp16 weights" mode
  SPGAN_CHECK_ARG(io->fmt == (p->precision == 3 ? 1 : 0) && io->w_fmt == io->fmt,
                  "spgan_conv_gemm: operand formats (A %d, W %d) do not match precision %d (bf16 planes for 1 and 2, fp16 planes "
                  "for 3)", io->fmt, (int)io->w_fmt, p->precision);
  SPGAN_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= SPGAN_MAX_TAPS, "spgan_conv_gemm: %d taps unsupported", p->ntaps);
  SPGAN_CHECK_ARG(p->in_stride == 1, "spgan_conv_gemm: in_stride %d unsupported on the tcgen05 path", p->in_stride);
  const int kp = io->kp;
  const int64_t a_rows = io->a_rows;
  SPGAN_CHECK_ARG(kp > 0 && kp % GEMM_UMMA_K == 0, "spgan_conv_gemm: kp=%d must be a positive multiple of 16", kp);
  SPGAN_CHECK_ARG(p->B >= 0 && p->H >= 0 && p->W >= 0 && p->Cout >= 0, "spgan_conv_gemm: negative size");
  SPGAN_CHECK_ARG(p->out_stride >= 1, "spgan_conv_gemm: out_stride must be >= 1");
  const int64_t rows = (int64_t)p->B * p->H * p->W;
  SPGAN_CHECK_ARG(rows > 0 ? (a_rows >= rows && a_rows % rows == 0) : a_rows == 0,
                  "spgan_conv_gemm: a_rows=%lld is not a multiple (polyphase planes) of B*H*W=%lld", (long long)a_rows, (long long)rows);
  SPGAN_CHECK_ARG(a_rows < 2147483647LL - 65536, "spgan_conv_gemm: too many packed rows");
  SPGAN_CHECK_ARG(rows < 2147483647LL - 65536, "spgan_conv_gemm: too many lattice points");
  if (rows == 0 || p->Cout == 0 || p->My == 0 || p->Mx == 0) return 0;
  SPGAN_CHECK_ARG(p->Cout >= 16, "spgan_conv_gemm: Cout=%d < 16 belongs on the SIMT path", p->Cout);
  SPGAN_CHECK_ARG(io->a_packed && io->w_packed, "spgan_conv_gemm: null operand pointer");
  SPGAN_CHECK_ARG(io->y || io->y_packed || io->rgb_w, "spgan_conv_gemm: no output sink");
  SPGAN_CHECK_ARG(((((uintptr_t)io->a_packed) | ((uintptr_t)io->w_packed)) & 15) == 0, "spgan_conv_gemm: packed operands must be 16-byte aligned");
  const bool general = io->y == nullptr || io->y_layout != 0 || io->y_packed != nullptr || io->rgb_w != nullptr || io->residual_nhwc != nullptr;
  if (general) {
    SPGAN_CHECK_ARG(p->Cout % 32 == 0, "spgan_conv_gemm: channels-last / packed / ToRGB sinks need Cout %% 32 == 0, got %d", p->Cout);
    SPGAN_CHECK_ARG(io->residual == nullptr, "spgan_conv_gemm: residual is only supported with a plain NCHW output");
    SPGAN_CHECK_ARG(io->y == nullptr || io->y_layout == 0 || (((uintptr_t)io->y) & 15) == 0, "spgan_conv_gemm: NHWC output must be 16-byte aligned");
  }
  if (io->y_packed) {
    SPGAN_CHECK_ARG(io->y_packed_cols >= p->Cout && io->y_packed_cols % 8 == 0 && (((uintptr_t)io->y_packed) & 15) == 0,
                    "spgan_conv_gemm: packed sink needs cols %% 8 == 0, cols >= Cout and a 16-byte aligned pointer");
    SPGAN_CHECK_ARG(io->y_packed_rows >= (int64_t)p->B * p->out_H * p->out_W, "spgan_conv_gemm: packed sink has too few rows");
    SPGAN_CHECK_ARG(io->y_packed_fmt == 0 || io->y_packed_fmt == 1, "spgan_conv_gemm: packed sink format must be 0 or 1");
  }
  if (io->rgb_w) {
    SPGAN_CHECK_ARG(io->rgb_part != nullptr && io->rgb_n >= 1 && io->rgb_n <= 3, "spgan_conv_gemm: ToRGB sink needs rgb_part and 1..3 channels");
    SPGAN_CHECK_ARG((((uintptr_t)io->rgb_w) & 15) == 0, "spgan_conv_gemm: rgb_w must be 16-byte aligned");
  }

  const GemmShape gs = gemm_shape(p, a_rows, (kp + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K);
  const int phases = gs.phases;
  const bool im2col = gs.im2col;
  GemmParams gp;
  gp.B = p->B;
  gp.im2col = im2col ? 1 : 0;
  gp.img_lo = phases * p->B;
  gp.rows = (int32_t)gs.rows_m;
  gp.Hl = im2col ? p->My : p->H;
  gp.Wl = im2col ? p->Mx : p->W;
  gp.My = p->My;
  gp.Mx = p->Mx;
  for (int t = 0; t < SPGAN_MAX_TAPS; ++t) gp.tap_ox[t] = gp.tap_oy[t] = gp.tap_img[t] = 0;
  for (int t = 0; t < p->ntaps; ++t) {
    // tap_dy carries the phase plane of the tap as phase * (B*H) (see the header): split it back
    const int bh = p->B * p->H;
    const int ph = p->tap_dy[t] / bh;
    gp.tap_oy[t] = p->tap_dy[t] - ph * bh;
    gp.tap_ox[t] = p->tap_dx[t];
    gp.tap_img[t] = ph * p->B;
    if (im2col)
      SPGAN_CHECK_ARG(p->tap_dy[t] >= 0 && p->tap_dx[t] >= 0 && gp.tap_oy[t] < 256 && gp.tap_ox[t] < 256 && ph < phases,
                      "spgan_conv_gemm: tap %d (%d, %d) outside the im2col offset range", t, p->tap_dy[t], p->tap_dx[t]);
  }
  gp.Cout = p->Cout;
  gp.out_H = p->out_H;
  gp.out_W = p->out_W;
  gp.out_stride = p->out_stride;
  gp.out_off_y = p->out_off_y;
  gp.out_off_x = p->out_off_x;
  gp.out_cstride = p->out_cstride ? p->out_cstride : (int64_t)p->out_H * p->out_W;
  gp.ntaps = p->ntaps;
  gp.kblocks = (kp + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  gp.last_ksteps = (kp - (gp.kblocks - 1) * GEMM_BLOCK_K) / GEMM_UMMA_K;
  for (int t = 0; t < p->ntaps; ++t) gp.tap_off[t] = p->tap_dy[t] * p->W + p->tap_dx[t];
  for (int t = p->ntaps; t < SPGAN_MAX_TAPS; ++t) gp.tap_off[t] = 0;
  gp.m_tiles = gs.m_tiles;
  const int block_n = gs.block_n;
  gp.n_tiles = gs.n_tiles;
  gp.out_scale = p->out_scale;
  gp.act = p->act;
  gp.act_alpha = p->act_alpha;
  gp.act_gain = p->act_gain;
  gp.a_f16 = io->fmt;
  gp.b_f16 = io->w_fmt;
  gp.y_nhwc = io->y_layout != 0 ? 1 : 0;
  gp.y_bstride = io->y_bstride ? io->y_bstride : (int64_t)p->out_H * p->out_W * p->Cout;
  gp.pk_rows = io->y_packed_rows;
  gp.pk_cols = io->y_packed_cols;
  gp.pk_f16 = io->y_packed_fmt;
  gp.rgb_n = io->rgb_w ? io->rgb_n : 0;
  gp.res_bstride = io->res_bstride ? io->res_bstride : (int64_t)p->out_H * p->out_W * p->Cout;
  GemmSinks sk;
  sk.y = io->y;
  sk.out_mul = io->out_mul;
  sk.noise = io->noise;
  sk.noise_w = io->noise_w;
  sk.bias = io->bias;
  sk.residual = io->residual;
  sk.residual_nhwc = io->residual_nhwc;
  sk.y_packed = io->y_packed;
  sk.next_mul = io->next_mul;
  sk.rgb_w = io->rgb_w;
  sk.rgb_part = io->rgb_part;

  CUtensorMap tmA, tmB;
  if (im2col) {
    // (kp channels, W, H, 2 * phases * B images): hi planes first, then the lo planes
    if (int e = encode_bf16_im2col_map(&tmA, io->a_packed, (cuuint64_t)kp, (cuuint64_t)p->W, (cuuint64_t)p->H,
                                       (cuuint64_t)2 * phases * p->B, p->Mx - p->W, p->My - p->H, GEMM_BLOCK_K, GEMM_BLOCK_M,
                                       "spgan_conv_gemm (A im2col map)"))
      return e;
  } else {
    cuuint64_t dims[3] = {(cuuint64_t)kp, (cuuint64_t)a_rows, 2};
    cuuint64_t strides[2] = {(cuuint64_t)kp * 2, (cuuint64_t)a_rows * kp * 2};
    cuuint32_t box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
    if (int e = encode_bf16_map(&tmA, io->a_packed, 3, dims, strides, box, "spgan_conv_gemm (A map)")) return e;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)kp, (cuuint64_t)p->Cout, (cuuint64_t)p->ntaps, 2};
    cuuint64_t strides[3] = {(cuuint64_t)kp * 2, (cuuint64_t)p->Cout * kp * 2, (cuuint64_t)p->ntaps * p->Cout * kp * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, (cuuint32_t)(gs.pair ? 128 : block_n), 1, 1};
    if (int e = encode_bf16_map(&tmB, io->w_packed, 4, dims, strides, box, "spgan_conv_gemm (B map)")) return e;
  }
  // optional second K segment: Y += A2[p, :] * W2[o, :] over kp2 more columns (the few channels that do not fill a
  // 64-wide block per tap, gathered over all taps into one dense slab instead of padding every tap)
  CUtensorMap tmA2 = tmA, tmB2 = tmB;
  gp.k2blocks = 0;
  if (io->kp2 > 0) {
    SPGAN_CHECK_ARG(io->a2_packed && io->w2_packed, "spgan_conv_gemm: kp2 > 0 needs a2_packed and w2_packed");
    SPGAN_CHECK_ARG(io->kp2 % GEMM_BLOCK_K == 0, "spgan_conv_gemm: kp2=%d must be a multiple of 64", io->kp2);
    SPGAN_CHECK_ARG(io->a2_rows >= gs.rows_m, "spgan_conv_gemm: a2_packed has %lld rows, the pass has %lld M rows",
                    (long long)io->a2_rows, (long long)gs.rows_m);
    SPGAN_CHECK_ARG(((((uintptr_t)io->a2_packed) | ((uintptr_t)io->w2_packed)) & 15) == 0, "spgan_conv_gemm: second-segment operands must be 16-byte aligned");
    gp.k2blocks = io->kp2 / GEMM_BLOCK_K;
    {
      cuuint64_t dims[3] = {(cuuint64_t)io->kp2, (cuuint64_t)io->a2_rows, 2};
      cuuint64_t strides[2] = {(cuuint64_t)io->kp2 * 2, (cuuint64_t)io->a2_rows * io->kp2 * 2};
      cuuint32_t box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
      if (int e = encode_bf16_map(&tmA2, io->a2_packed, 3, dims, strides, box, "spgan_conv_gemm (A2 map)")) return e;
    }
    {
      cuuint64_t dims[4] = {(cuuint64_t)io->kp2, (cuuint64_t)p->Cout, 1, 2};
      cuuint64_t strides[3] = {(cuuint64_t)io->kp2 * 2, (cuuint64_t)p->Cout * io->kp2 * 2, (cuuint64_t)p->Cout * io->kp2 * 2};
      cuuint32_t box[4] = {GEMM_BLOCK_K, (cuuint32_t)(gs.pair ? 128 : block_n), 1, 1};
      if (int e = encode_bf16_map(&tmB2, io->w2_packed, 4, dims, strides, box, "spgan_conv_gemm (W2 map)")) return e;
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (gs.pair) {
    if (p->precision == 1) return launch_gemm2<3>(tmA, tmB, tmA2, tmB2, gp, sk, st);
    if (p->precision == 3) return launch_gemm2<2>(tmA, tmB, tmA2, tmB2, gp, sk, st);
    return launch_gemm2<1>(tmA, tmB, tmA2, tmB2, gp, sk, st);
  }
  if (p->precision == 1)
    return block_n == 256 ? launch_gemm<3, 256>(tmA, tmB, tmA2, tmB2, gp, sk, st) : launch_gemm<3, 128>(tmA, tmB, tmA2, tmB2, gp, sk, st);
  if (p->precision == 3)
    return block_n == 256 ? launch_gemm<2, 256>(tmA, tmB, tmA2, tmB2, gp, sk, st) : launch_gemm<2, 128>(tmA, tmB, tmA2, tmB2, gp, sk, st);
  return block_n == 256 ? launch_gemm<1, 256>(tmA, tmB, tmA2, tmB2, gp, sk, st) : launch_gemm<1, 128>(tmA, tmB, tmA2, tmB2, gp, sk, st);
}

extern "C" int spgan_conv_gemm(const SpganConvPass* p, float* y, const uint16_t* a_packed, int64_t a_rows, int kp,
                               const uint16_t* w_packed, const float* out_mul, const float* noise, const float* noise_w,
                               const float* bias, const float* residual, void* stream) {
  SPGAN_CHECK_ARG(p != nullptr, "spgan_conv_gemm: null pass descriptor");
  SpganGemmIO io = {};
  io.a_packed = a_packed;
  io.a_rows = a_rows;
  io.kp = kp;
  io.fmt = p->precision == 3 ? 1 : 0;
  io.w_fmt = io.fmt;
  io.w_packed = w_packed;
  io.out_mul = out_mul;
  io.noise = noise;
  io.noise_w = noise_w;
  io.bias = bias;
  io.residual = residual;
  io.y = y;
  SPGAN_CHECK_ARG(y != nullptr, "spgan_conv_gemm: null output pointer");
  return spgan_conv_gemm_ex(p, &io, stream);
}

// Fused spherical modulated conv for sm_100a: the bilinear gather is the A-operand PRODUCER of the tcgen05 implicit GEMM.
//
//   Y[p, o] = sum_{t < 9} sum_{k < Cp} G[p, t, k] * Wp[o][t*Cp + k]
//   G[p, t, k] = in_mul[g, k] * enc_k( sum_{corner} w_corner(p, t) * src_k[corner(p, t)] )
//
// replaces models/spgan_ops_gs.py:791-814 (F.grid_sample x2 -> tanh / cos pi / sin pi -> cat -> F.conv2d(stride 3, groups=B))
// without the 9x gathered fp32 tensor of the reference AND without the [B*H*W][9*Cp] 16-bit operand spgan_sphere_pack
// writes (452 MB at 35x35, B = 32): the operand only ever exists as 128 x 64 tiles in shared memory.
//
// One persistent CTA per SM, 320 threads:
//   warp 0      TMA: streams the weight tiles {B_hi, B_lo} of k-block (kb, t) through the smem ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (same schedule as conv_gemm_kernel)
//   warps 2..9  PRODUCERS, then epilogue.  Per tile they first compute the 128 x 9 corner table (4 corner offsets + 4
//               bilinear weights per (pixel, tap), ATen's fp32 index sequence bit for bit) into shared memory; then for every
//               k-block they gather their 16 rows from the channels-last fp32 input (L2-resident: 40 MB at 35x35, B = 32;
//               each warp-level load is 128 contiguous bytes of one corner pixel), blend, encode the coordinate channels,
//               apply the style modulation, split into two 16-bit planes and store them with the 128-byte swizzle pattern
//               the UMMA descriptor expects (fence.proxy.async, then one mbarrier arrive per warp).  After the last k-block
//               of tile i they run the epilogue of tile i-1 (its accumulator is complete by then; two TMEM stages), so
//               the tensor pipe keeps working on tile i while tile i-1 is written out.
// K order: kb-major, tap-minor, so the channel decode (the reference's flat (1,B*C)++(1,B*3) concat table, chan_map) and
// the modulation are looked up once per 64-channel block and reused for its 9 taps; the packed weight is the merged layout
// [o][t*Cp + k] of spgan_pack_weight, any K order of which is addressable by TMA.
// Roofline: tensor pipe (2 * rows * Cout * 9 * Cp FLOP, x3 issued in bf16x3); the producers need ~80 % of the MMA time of a
// stage in issue slots and ~85 B/clk/SM from L1/L2.
#include "gemm_epilogue.cuh"
#include "sphere_taps.cuh"
#include "pack_math.cuh"

#include <atomic>

namespace {

constexpr int SPH_THREADS = 320;
constexpr int SPH_PROD_WARPS = 8;
constexpr int SPH_BATCH = 8;  // rows gathered per load batch of a producer warp
constexpr int SPH_TABLE_ENTRIES = GEMM_BLOCK_M * 9;
constexpr int SPH_TABLE_BYTES = SPH_TABLE_ENTRIES * 5 * 4;  // int base + 4 float weights, structure of arrays

struct SphereIn {
  const float* xh;           // (B, H, W, C) fp32 channels-last features
  const float* coords;       // (B, nc, H, W) fp32 raw coordinate planes or null
  const float* grid;         // (1, 3H, 3W, 2) fp32 tap grid shared by the batch
  const float* in_mul;       // (B, C + nc) style modulation or null
  const uint32_t* chan_map;  // (B, Cp): bit 31 coordinate plane, bits [15,31) source sample, bits [0,15) source channel
  int32_t B, C, nc, H, W, Cp;
};

template <int kPasses, int kBlockN>
struct SphereSmem {
  static constexpr int kBTileBytes = kBlockN * GEMM_BLOCK_K * 2;
  static constexpr int kAPlanes = kPasses >= 2 ? 2 : 1;
  static constexpr int kBPlanes = kPasses == 3 ? 2 : 1;
  static constexpr int kStageBytes = kAPlanes * A_TILE_BYTES + kBPlanes * kBTileBytes;
  static constexpr int kStages = (196608 / kStageBytes) < 4 ? (196608 / kStageBytes) : 4;
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + SPH_TABLE_BYTES + kBarrierBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int kPasses, int kBlockN>
__global__ void __launch_bounds__(SPH_THREADS, 1)
sphere_gemm_kernel(const __grid_constant__ CUtensorMap tmB, const GemmParams gp, const GemmSinks sk, const SphereIn si) {
  using S = SphereSmem<kPasses, kBlockN>;
  constexpr int kStages = S::kStages;
  constexpr int kBTile = S::kBTileBytes;
  constexpr bool kF16 = kPasses == 2;  // the 2-MMA mode is the fp16 split (precision 3); 1 and 3 passes use bf16 planes
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic-address view of the aligned region
  const uint32_t table_base = smem_base + S::kTileBytes;
  const uint32_t bar_base = table_base + SPH_TABLE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  int* tbl_base = reinterpret_cast<int*>(smem_gen + S::kTileBytes);
  float* tbl_w = reinterpret_cast<float*>(tbl_base + SPH_TABLE_ENTRIES);  // [4][SPH_TABLE_ENTRIES]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = gp.m_tiles * gp.n_tiles;
  const int kblocks = si.Cp / GEMM_BLOCK_K;
  const int kiters = 9 * kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1 + SPH_PROD_WARPS);  // the TMA thread's expect_tx arrive + one arrive per producer warp
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
    // ===================================================================== TMA: weight tiles
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n0 = (tile % gp.n_tiles) * kBlockN;
        for (int kb = 0; kb < kblocks; ++kb) {
          for (int t = 0; t < 9; ++t) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t sb = smem_base + stage * S::kStageBytes + S::kAPlanes * A_TILE_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), S::kBPlanes * kBTile);
            const int kcol = t * si.Cp + kb * GEMM_BLOCK_K;
            tma_load_4d(sb, &tmB, full_bar(stage), kcol, n0, 0, 0);
            if (S::kBPlanes == 2) tma_load_4d(sb + kBTile, &tmB, full_bar(stage), kcol, n0, 0, 1);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
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
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            const uint32_t koff = k * GEMM_UMMA_K * 2;
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff);
            const uint64_t db_hi = umma_desc_sw128(b_hi + koff);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (kPasses == 3) tc_mma_f16(d_tmem, da_hi, umma_desc_sw128(b_lo + koff), idesc, 1u);
            if (kPasses >= 2) tc_mma_f16(d_tmem, umma_desc_sw128(a_lo + koff), db_hi, idesc, 1u);
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
    // ===================================================================== producers (gather -> swizzled A tiles) + epilogue
    const int tid2 = threadIdx.x - 64;
    const int w2 = tid2 >> 5;
    const int HW = si.H * si.W;
    const int Ct = si.C + si.nc;
    const float* __restrict__ xh = si.xh;
    const float* __restrict__ coords = si.coords;
    int stage = 0;
    uint32_t phase = 0;
    int titer = 0;
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
      const int m0 = (tile / gp.n_tiles) * GEMM_BLOCK_M;
      // ---- corner table of the tile's 128 pixels x 9 taps
      named_bar_sync(1, SPH_PROD_WARPS * 32);  // every producer is done reading the previous tile's table
      for (int e = tid2; e < SPH_TABLE_ENTRIES; e += SPH_PROD_WARPS * 32) {
        const int r = e / 9, t = e - r * 9;
        const int p = m0 + r;
        int base = 0;
        float w0 = 0.f, w1 = 0.f, w2f = 0.f, w3 = 0.f;
        if (p < gp.rows) {
          const int pix = p % HW;
          const int py = pix / si.W, px = pix - py * si.W;
          const TapCorners cn = tap_corners(si.grid, 0, si.H, si.W, py, px, t / 3, t - (t / 3) * 3);
          base = cn.o_nw * 4 + (cn.o_sw != cn.o_nw ? 2 : 0) + (cn.o_ne != cn.o_nw ? 1 : 0);
          w0 = cn.w_nw;
          w1 = cn.w_ne;
          w2f = cn.w_sw;
          w3 = cn.w_se;
        }
        tbl_base[e] = base;
        tbl_w[e] = w0;
        tbl_w[SPH_TABLE_ENTRIES + e] = w1;
        tbl_w[2 * SPH_TABLE_ENTRIES + e] = w2f;
        tbl_w[3 * SPH_TABLE_ENTRIES + e] = w3;
      }
      named_bar_sync(1, SPH_PROD_WARPS * 32);
      // rows [0, rb) of the tile belong to sample gA, the rest to gA + 1 (the host guarantees H*W >= 128)
      const int gA = m0 / HW;
      const int rb = (gA + 1) * HW - m0;
      for (int kb = 0; kb < kblocks; ++kb) {
        // channel decode of this lane's two columns (lane, lane + 32) for both samples: source plane offset, kind, modulation
        uint32_t soff[2][2];
        int kind[2][2];  // 0 zero padding, 1 feature, 2..4 coordinate plane 0..2 (tanh / cos pi / sin pi)
        float mulv[2][2];
#pragma unroll
        for (int sidx = 0; sidx < 2; ++sidx) {
          const int g = gA + sidx;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int k = kb * GEMM_BLOCK_K + lane + 32 * u;
            uint32_t m = 0xFFFFFFFFu;
            if (g < si.B) m = __ldg(si.chan_map + (int64_t)g * si.Cp + k);
            const bool valid = m != 0xFFFFFFFFu;
            const uint32_t bs = (m >> 15) & 0xFFFFu, cs = m & 0x7FFFu;
            const bool is_coord = valid && (m >> 31) != 0;
            kind[sidx][u] = !valid ? 0 : (is_coord ? 2 + (int)cs : 1);
            soff[sidx][u] = !valid ? 0u : (is_coord ? (bs * (uint32_t)si.nc + cs) * (uint32_t)HW : bs * (uint32_t)HW * (uint32_t)si.C + cs);
            mulv[sidx][u] = (valid && si.in_mul && k < Ct) ? __ldg(si.in_mul + (int64_t)g * Ct + k) : 1.f;
          }
        }
        // Fast path (every k-block except the one that holds the coordinate planes, i.e. one block of the LAST group under
        // the reference's flat concat): all 64 columns of both samples are feature planes or zero padding (modulation 0), so
        // the gather is branch-free: 32-bit element offsets against the uniform base pointer, 8 loads in flight per row.
        const bool fast = __all_sync(0xffffffffu, kind[0][0] <= 1 && kind[0][1] <= 1 && kind[1][0] <= 1 && kind[1][1] <= 1);
        const uint32_t soA0 = soff[0][0], soA1 = soff[0][1], soB0 = soff[1][0], soB1 = soff[1][1];
        const float mA0 = kind[0][0] ? mulv[0][0] : 0.f, mA1 = kind[0][1] ? mulv[0][1] : 0.f;
        const float mB0 = kind[1][0] ? mulv[1][0] : 0.f, mB1 = kind[1][1] ? mulv[1][1] : 0.f;
        const uint32_t lane_col = (uint32_t)(((lane >> 3) << 4) | ((lane & 7) << 1));  // byte offset of column `lane` in a 128 B row
        for (int t = 0; t < 9; ++t) {
          if (lane == 0) mbar_wait(empty_bar(stage), phase ^ 1u);
          __syncwarp();
          uint8_t* a_hi = smem_gen + stage * S::kStageBytes;
          uint8_t* a_lo = a_hi + A_TILE_BYTES;
          if (fast) {
            const uint32_t uC = (uint32_t)si.C, uWC = (uint32_t)si.W * (uint32_t)si.C;
            // 8 rows per batch: 64 independent loads in flight per lane (the producers are 8 warps on an SM that could hold
            // 64: memory-level parallelism has to come from the batch, not from occupancy); the bilinear weights are re-read
            // from the shared table when they are used instead of being held across the loads
#pragma unroll 1
            for (int i0 = 0; i0 < GEMM_BLOCK_M / SPH_PROD_WARPS; i0 += SPH_BATCH) {
              float cv[SPH_BATCH][2][4];
#pragma unroll
              for (int ii = 0; ii < SPH_BATCH; ++ii) {
                const int r = w2 * (GEMM_BLOCK_M / SPH_PROD_WARPS) + i0 + ii;
                const uint32_t base = (uint32_t)tbl_base[r * 9 + t];
                const uint32_t c0 = (base >> 2) * uC;
                const uint32_t c1 = c0 + ((base & 1u) ? uC : 0u);
                const uint32_t c2 = c0 + ((base & 2u) ? uWC : 0u);
                const uint32_t c3 = c2 + (c1 - c0);
                const bool sb = r >= rb;
                const uint32_t s0 = sb ? soB0 : soA0, s1 = sb ? soB1 : soA1;
                cv[ii][0][0] = __ldg(xh + (s0 + c0));
                cv[ii][0][1] = __ldg(xh + (s0 + c1));
                cv[ii][0][2] = __ldg(xh + (s0 + c2));
                cv[ii][0][3] = __ldg(xh + (s0 + c3));
                cv[ii][1][0] = __ldg(xh + (s1 + c0));
                cv[ii][1][1] = __ldg(xh + (s1 + c1));
                cv[ii][1][2] = __ldg(xh + (s1 + c2));
                cv[ii][1][3] = __ldg(xh + (s1 + c3));
              }
#pragma unroll
              for (int ii = 0; ii < SPH_BATCH; ++ii) {
                const int r = w2 * (GEMM_BLOCK_M / SPH_PROD_WARPS) + i0 + ii;
                const int e = r * 9 + t;
                const float w0 = tbl_w[e], w1 = tbl_w[SPH_TABLE_ENTRIES + e], w2v = tbl_w[2 * SPH_TABLE_ENTRIES + e],
                            w3 = tbl_w[3 * SPH_TABLE_ENTRIES + e];
                const bool sb = r >= rb;
                // same association as spgan_sphere_pack: ((a*w_nw + b*w_ne) + c*w_sw) + d*w_se
                float v0 = cv[ii][0][0] * w0 + cv[ii][0][1] * w1 + cv[ii][0][2] * w2v + cv[ii][0][3] * w3;
                float v1 = cv[ii][1][0] * w0 + cv[ii][1][1] * w1 + cv[ii][1][2] * w2v + cv[ii][1][3] * w3;
                v0 *= sb ? mB0 : mA0;
                v1 *= sb ? mB1 : mA1;
                uint16_t h0, l0, h1, l1;
                split16<kF16>(v0, h0, l0);
                split16<kF16>(v1, h1, l1);
                const uint32_t off0 = (uint32_t)r * 128u + (lane_col ^ ((uint32_t)(r & 7) << 4));
                const uint32_t off1 = off0 ^ 64u;  // column lane + 32: chunk index + 4
                *reinterpret_cast<uint16_t*>(a_hi + off0) = h0;
                *reinterpret_cast<uint16_t*>(a_hi + off1) = h1;
                if (S::kAPlanes == 2) {
                  *reinterpret_cast<uint16_t*>(a_lo + off0) = l0;
                  *reinterpret_cast<uint16_t*>(a_lo + off1) = l1;
                }
              }
            }
          } else {
#pragma unroll 1
          for (int i0 = 0; i0 < GEMM_BLOCK_M / SPH_PROD_WARPS; i0 += 4) {
            float cv[4][2][4];
            float wv[4][4];
            int sel[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int r = w2 * (GEMM_BLOCK_M / SPH_PROD_WARPS) + i0 + ii;
              const int e = r * 9 + t;
              const int base = tbl_base[e];
              wv[ii][0] = tbl_w[e];
              wv[ii][1] = tbl_w[SPH_TABLE_ENTRIES + e];
              wv[ii][2] = tbl_w[2 * SPH_TABLE_ENTRIES + e];
              wv[ii][3] = tbl_w[3 * SPH_TABLE_ENTRIES + e];
              const int o_nw = base >> 2;
              const int dx = base & 1, dy = (base & 2) ? si.W : 0;
              const int sidx = r < rb ? 0 : 1;
              sel[ii] = sidx;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int kd = sidx ? kind[1][u] : kind[0][u];
                const uint32_t so = sidx ? soff[1][u] : soff[0][u];
                if (kd == 0) {
                  cv[ii][u][0] = cv[ii][u][1] = cv[ii][u][2] = cv[ii][u][3] = 0.f;
                } else if (kd == 1) {
                  const float* sp = xh + so;
                  cv[ii][u][0] = __ldg(sp + (int64_t)o_nw * si.C);
                  cv[ii][u][1] = __ldg(sp + (int64_t)(o_nw + dx) * si.C);
                  cv[ii][u][2] = __ldg(sp + (int64_t)(o_nw + dy) * si.C);
                  cv[ii][u][3] = __ldg(sp + (int64_t)(o_nw + dy + dx) * si.C);
                } else {
                  const float* sp = coords + so;
                  cv[ii][u][0] = __ldg(sp + o_nw);
                  cv[ii][u][1] = __ldg(sp + o_nw + dx);
                  cv[ii][u][2] = __ldg(sp + o_nw + dy);
                  cv[ii][u][3] = __ldg(sp + o_nw + dy + dx);
                }
              }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
              const int r = w2 * (GEMM_BLOCK_M / SPH_PROD_WARPS) + i0 + ii;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                // same association as spgan_sphere_pack: ((a*w_nw + b*w_ne) + c*w_sw) + d*w_se
                float val = cv[ii][u][0] * wv[ii][0] + cv[ii][u][1] * wv[ii][1] + cv[ii][u][2] * wv[ii][2] + cv[ii][u][3] * wv[ii][3];
                const int kd = sel[ii] ? kind[1][u] : kind[0][u];
                if (kd >= 2) {
                  if (kd == 2) val = tanhf(val);
                  else if (kd == 3) val = cosf(val * 3.14159274101257324f);
                  else val = sinf(val * 3.14159274101257324f);
                }
                val *= sel[ii] ? mulv[1][u] : mulv[0][u];
                uint16_t hi, lo;
                split16<kF16>(val, hi, lo);
                const int c = lane + 32 * u;
                const int off = r * 128 + ((((c >> 3) ^ (r & 7))) << 4) + ((c & 7) << 1);
                *reinterpret_cast<uint16_t*>(a_hi + off) = hi;
                if (S::kAPlanes == 2) *reinterpret_cast<uint16_t*>(a_lo + off) = lo;
              }
            }
          }
          }
          // generic-proxy writes -> visible to the tensor core's async proxy, then one arrive per warp
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      // ---- epilogue of the PREVIOUS tile (complete by now), overlapping the MMAs that drain this tile's last stages
      if (prev_tile >= 0) {
        const int pt = titer - 1;
        const int as = pt & 1;
        gemm_epilogue_tile<kBlockN>(gp, sk, (prev_tile / gp.n_tiles) * GEMM_BLOCK_M, prev_tile % gp.n_tiles, warp, lane, tfull_bar(as), (uint32_t)(pt >> 1) & 1u,
                                    tmem_base + (uint32_t)(as * kBlockN));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      const int pt = titer - 1;
      const int as = pt & 1;
      gemm_epilogue_tile<kBlockN>(gp, sk, (prev_tile / gp.n_tiles) * GEMM_BLOCK_M, prev_tile % gp.n_tiles, warp, lane, tfull_bar(as), (uint32_t)(pt >> 1) & 1u,
                                  tmem_base + (uint32_t)(as * kBlockN));
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

// ------------------------------------------------------------------------------------------------ vectorised producer (v4)
// Same kernel anatomy, but the producers read the REPACKED gather source xg[g][pixel][SPV_LD] (spgan_sphere_concat_repack:
// the flat-concat channels of every group as 16-byte aligned rows) and the K layout of the structure chain: 256 main columns
// per tap (4 K blocks x 9 taps) + ONE 64-wide tail block holding the three trailing channels of all nine taps.  Per stage a
// producer warp makes 4 passes of 4 rows x 8 lanes: two aligned 128-bit loads per corner, packed FFMA2 blend, pairwise 16-bit
// split, one 16-byte swizzled store per plane — ~75 instructions per (4 rows x 64 columns) against ~96 per (1 row x 64
// columns) of the scalar producer above, which made that kernel producer-bound (0.84 ms against 0.48 ms for pack + GEMM).
struct SphereV4In {
  const float* xg;           // (B, H*W, SPV_LD) fp32 repacked gather source
  const float* grid;         // (B / grid_group, 3H, 3W, 2)
  const float* in_mul;       // (B, Ct) or null
  const uint32_t* chan_map;  // (B, cmap_ld): only the coordinate-plane flags are read here
  int32_t B, Ct, H, W, grid_group, cmap_ld;
};

template <int kPasses>
struct SphereV4Smem {
  static constexpr int kBlockN = 256;
  static constexpr int kBTileBytes = kBlockN * GEMM_BLOCK_K * 2;
  static constexpr int kAPlanes = kPasses >= 2 ? 2 : 1;
  static constexpr int kBPlanes = kPasses == 3 ? 2 : 1;
  static constexpr int kStageBytes = kAPlanes * A_TILE_BYTES + kBPlanes * kBTileBytes;
  static constexpr int kStages = (196608 / kStageBytes) < 4 ? (196608 / kStageBytes) : 4;
  static constexpr int kTileBytes = kStageBytes * kStages;
  static constexpr int kTableBytes = SPH_TABLE_BYTES;  // nw corner + 2 flag bits, 4 weights per (row, tap)
  static constexpr int kMulBytes = 2 * SPV_LD * 4;               // modulation rows of the tile's two groups
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kTileBytes + kTableBytes + kMulBytes + kBarrierBytes + 1024;
  static_assert(kTotal <= 232448, "sphere_gemm_v4: shared memory budget exceeded");
};

template <int kPasses>
__global__ void __launch_bounds__(SPH_THREADS, 1)
sphere_gemm_v4_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const GemmParams gp,
                      const GemmSinks sk, const SphereV4In si) {
  using S = SphereV4Smem<kPasses>;
  constexpr int kStages = S::kStages;
  constexpr int kBlockN = S::kBlockN;
  constexpr int kBTile = S::kBTileBytes;
  constexpr bool kF16 = kPasses == 2;
  constexpr int C = SPV_C;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  int* tbl_base = reinterpret_cast<int*>(smem_gen + S::kTileBytes);            // [SPH_TABLE_ENTRIES]: nw pixel * 4 + (dy << 1) + dx
  float* tbl_w = reinterpret_cast<float*>(tbl_base + SPH_TABLE_ENTRIES);       // [4][SPH_TABLE_ENTRIES]
  float* s_mul = reinterpret_cast<float*>(smem_gen + S::kTileBytes + S::kTableBytes);  // [2][SPV_LD]
  const uint32_t bar_base = smem_base + S::kTileBytes + S::kTableBytes + S::kMulBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = gp.m_tiles;  // one 256-wide N tile (Cout <= 256)
  constexpr int kMainBlocks = C / GEMM_BLOCK_K;  // 4
  const int kiters = 9 * kMainBlocks + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1 + SPH_PROD_WARPS);
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
    // ===================================================================== TMA: weight tiles
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int it = 0; it < kiters; ++it) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sb = smem_base + stage * S::kStageBytes + S::kAPlanes * A_TILE_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), S::kBPlanes * kBTile);
          if (it < 9 * kMainBlocks) {
            const int kb = it / 9, t = it - kb * 9;
            const int kcol = t * C + kb * GEMM_BLOCK_K;
            tma_load_4d(sb, &tmB, full_bar(stage), kcol, 0, 0, 0);
            if (S::kBPlanes == 2) tma_load_4d(sb + kBTile, &tmB, full_bar(stage), kcol, 0, 0, 1);
          } else {
            tma_load_4d(sb, &tmB2, full_bar(stage), 0, 0, 0, 0);
            if (S::kBPlanes == 2) tma_load_4d(sb + kBTile, &tmB2, full_bar(stage), 0, 0, 0, 1);
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
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      int n_eff = gp.Cout > kBlockN ? kBlockN : ((gp.Cout + 15) & ~15);
      const uint32_t idesc = umma_idesc_16(n_eff, gp.a_f16 != 0, gp.b_f16 != 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
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
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            const uint32_t koff = k * GEMM_UMMA_K * 2;
            const uint64_t da_hi = umma_desc_sw128(a_hi + koff);
            const uint64_t db_hi = umma_desc_sw128(b_hi + koff);
            tc_mma_f16(d_tmem, da_hi, db_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
            if (kPasses == 3) tc_mma_f16(d_tmem, da_hi, umma_desc_sw128(b_lo + koff), idesc, 1u);
            if (kPasses >= 2) tc_mma_f16(d_tmem, umma_desc_sw128(a_lo + koff), db_hi, idesc, 1u);
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
    // ===================================================================== producers + epilogue (warps 2..9)
    const int tid2 = threadIdx.x - 64;
    const int w2 = tid2 >> 5;
    const int HW = si.H * si.W;
    const int Cx = si.Ct - C;
    const int rsub = lane >> 3, chunk = lane & 7;
    int stage = 0;
    uint32_t phase = 0;
    int titer = 0;
    int prev_tile = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
      const int m0 = tile * GEMM_BLOCK_M;
      const int gA = m0 / HW;
      const int rb = (gA + 1) * HW - m0;  // rows [0, rb) belong to group gA, the rest to gA + 1 (H*W >= 128)
      // ---- corner table of the tile's 128 pixels x 9 taps and the modulation rows of its two groups
      named_bar_sync(1, SPH_PROD_WARPS * 32);
      for (int e = tid2; e < SPH_TABLE_ENTRIES; e += SPH_PROD_WARPS * 32) {
        const int r = e / 9, t = e - r * 9;
        const int p = m0 + r;
        int base = 0;
        float w0 = 0.f, w1 = 0.f, w2f = 0.f, w3 = 0.f;
        if (p < gp.rows) {
          const int g = p / HW;
          const int pix = p - g * HW;
          const int py = pix / si.W, px = pix - py * si.W;
          const TapCorners cn = tap_corners(si.grid, g / si.grid_group, si.H, si.W, py, px, t / 3, t - (t / 3) * 3);
          base = cn.o_nw * 4 + (cn.o_sw != cn.o_nw ? 2 : 0) + (cn.o_ne != cn.o_nw ? 1 : 0);
          w0 = cn.w_nw;
          w1 = cn.w_ne;
          w2f = cn.w_sw;
          w3 = cn.w_se;
        }
        tbl_base[e] = base;
        tbl_w[e] = w0;
        tbl_w[SPH_TABLE_ENTRIES + e] = w1;
        tbl_w[2 * SPH_TABLE_ENTRIES + e] = w2f;
        tbl_w[3 * SPH_TABLE_ENTRIES + e] = w3;
      }
      for (int k = tid2; k < 2 * SPV_LD; k += SPH_PROD_WARPS * 32) {
        const int gs = k / SPV_LD, kk = k - gs * SPV_LD;
        const int g = gA + gs;
        s_mul[k] = (kk < si.Ct && g < si.B) ? (si.in_mul ? __ldg(si.in_mul + (int64_t)g * si.Ct + kk) : 1.f) : 0.f;
      }
      named_bar_sync(1, SPH_PROD_WARPS * 32);
      for (int kb = 0; kb < kMainBlocks; ++kb) {
        const int k0 = kb * GEMM_BLOCK_K + 8 * chunk;
        // coordinate-plane flags of this lane's 8 columns, for both groups of the tile (2 bits per column)
        uint32_t kinds[2] = {0u, 0u};
#pragma unroll
        for (int gs = 0; gs < 2; ++gs) {
          if (gA + gs < si.B) {
            const uint32_t* mrow = si.chan_map + (int64_t)(gA + gs) * si.cmap_ld + k0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t m = __ldg(mrow + j);
              if (m != 0xFFFFFFFFu && (m >> 31)) kinds[gs] |= (1u + (m & 3u)) << (2 * j);
            }
          }
        }
        const bool any_kind = __any_sync(0xffffffffu, (kinds[0] | kinds[1]) != 0u);
        for (int t = 0; t < 9; ++t) {
          if (lane == 0) mbar_wait(empty_bar(stage), phase ^ 1u);
          __syncwarp();
          uint8_t* a_hi = smem_gen + stage * S::kStageBytes;
          uint8_t* a_lo = a_hi + A_TILE_BYTES;
#pragma unroll 2
          for (int pass = 0; pass < 4; ++pass) {
            const int r = w2 * 16 + pass * 4 + rsub;
            const int e = r * 9 + t;
            const int gs = r >= rb ? 1 : 0;
            const float* src = si.xg + (int64_t)min(gA + gs, si.B - 1) * HW * SPV_LD + k0;
            float4 xa[4], xb[4];
            float wt[4];
            int co[4];
            {
              const int base = tbl_base[e];
              co[0] = (base >> 2) * SPV_LD;
              co[1] = co[0] + ((base & 1) ? SPV_LD : 0);
              co[2] = co[0] + ((base & 2) ? si.W * SPV_LD : 0);
              co[3] = co[2] + (co[1] - co[0]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4* q = reinterpret_cast<const float4*>(src + co[c]);
              xa[c] = __ldg(q);
              xb[c] = __ldg(q + 1);
              wt[c] = tbl_w[c * SPH_TABLE_ENTRIES + e];
            }
            const unsigned long long w0 = pack_f2(wt[0], wt[0]), w1 = pack_f2(wt[1], wt[1]), w2p = pack_f2(wt[2], wt[2]),
                                     w3 = pack_f2(wt[3], wt[3]);
            unsigned long long acc[4];
            acc[0] = fma2(pack_f2(xa[3].x, xa[3].y), w3, fma2(pack_f2(xa[2].x, xa[2].y), w2p, fma2(pack_f2(xa[1].x, xa[1].y), w1, mul2(pack_f2(xa[0].x, xa[0].y), w0))));
            acc[1] = fma2(pack_f2(xa[3].z, xa[3].w), w3, fma2(pack_f2(xa[2].z, xa[2].w), w2p, fma2(pack_f2(xa[1].z, xa[1].w), w1, mul2(pack_f2(xa[0].z, xa[0].w), w0))));
            acc[2] = fma2(pack_f2(xb[3].x, xb[3].y), w3, fma2(pack_f2(xb[2].x, xb[2].y), w2p, fma2(pack_f2(xb[1].x, xb[1].y), w1, mul2(pack_f2(xb[0].x, xb[0].y), w0))));
            acc[3] = fma2(pack_f2(xb[3].z, xb[3].w), w3, fma2(pack_f2(xb[2].z, xb[2].w), w2p, fma2(pack_f2(xb[1].z, xb[1].w), w1, mul2(pack_f2(xb[0].z, xb[0].w), w0))));
            float v[8];
            if (any_kind) {
              const uint32_t kd = gs ? kinds[1] : kinds[0];
              if (kd != 0u) {
#pragma unroll
                for (int j = 0; j < 4; ++j) unpack_f2(acc[j], v[2 * j], v[2 * j + 1]);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = encode_coord(v[j], (int)((kd >> (2 * j)) & 3u));
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = pack_f2(v[2 * j], v[2 * j + 1]);
              }
            }
            const float4 m0 = *reinterpret_cast<const float4*>(&s_mul[gs * SPV_LD + k0]);
            const float4 m1 = *reinterpret_cast<const float4*>(&s_mul[gs * SPV_LD + k0 + 4]);
            acc[0] = mul2(acc[0], pack_f2(m0.x, m0.y));
            acc[1] = mul2(acc[1], pack_f2(m0.z, m0.w));
            acc[2] = mul2(acc[2], pack_f2(m1.x, m1.y));
            acc[3] = mul2(acc[3], pack_f2(m1.z, m1.w));
            uint32_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              unpack_f2(acc[j], v[2 * j], v[2 * j + 1]);
              split_pair<kF16>(v[2 * j], v[2 * j + 1], h[j], l[j]);
            }
            const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(chunk ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
            if (S::kAPlanes == 2) *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(full_bar(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      // ---- tail block: columns t*Cx + j of the trailing channels C .. Ct-1, zeros beyond 9*Cx
      {
        if (lane == 0) mbar_wait(empty_bar(stage), phase ^ 1u);
        __syncwarp();
        uint8_t* a_hi = smem_gen + stage * S::kStageBytes;
        uint8_t* a_lo = a_hi + A_TILE_BYTES;
        const int tcol = lane;  // this lane's tail column (lanes >= 9*Cx write zeros); columns 32..63 are zero as well
        const int tt = Cx > 0 ? tcol / Cx : 0, tj = Cx > 0 ? tcol - tt * Cx : 0;
        const bool live = tcol < 9 * Cx;
        int tkind[2] = {0, 0};
        if (live) {
#pragma unroll
          for (int gs = 0; gs < 2; ++gs) {
            if (gA + gs < si.B) {
              const uint32_t m = __ldg(si.chan_map + (int64_t)(gA + gs) * si.cmap_ld + C + tj);
              if (m != 0xFFFFFFFFu && (m >> 31)) tkind[gs] = 1 + (int)(m & 3u);
            }
          }
        }
        for (int i = 0; i < 16; ++i) {
          const int r = w2 * 16 + i;
          uint16_t hh = 0, ll = 0;
          if (live) {
            const int e = r * 9 + tt;
            const int gs = r >= rb ? 1 : 0;
            const float* sp = si.xg + (int64_t)min(gA + gs, si.B - 1) * HW * SPV_LD + C + tj;
            const int base = tbl_base[e];
            const int c0 = (base >> 2) * SPV_LD;
            const int c1 = c0 + ((base & 1) ? SPV_LD : 0);
            const int c2 = c0 + ((base & 2) ? si.W * SPV_LD : 0);
            const int c3 = c2 + (c1 - c0);
            float val = __ldg(sp + c3) * tbl_w[3 * SPH_TABLE_ENTRIES + e] +
                        (__ldg(sp + c2) * tbl_w[2 * SPH_TABLE_ENTRIES + e] +
                         (__ldg(sp + c1) * tbl_w[SPH_TABLE_ENTRIES + e] + __ldg(sp + c0) * tbl_w[e]));
            const int kd = gs ? tkind[1] : tkind[0];
            if (kd) val = encode_coord(val, kd);
            val *= s_mul[gs * SPV_LD + C + tj];
            split16<kF16>(val, hh, ll);
          }
          const uint32_t off0 = (uint32_t)r * 128u + ((uint32_t)((tcol >> 3) ^ (r & 7)) << 4) + ((uint32_t)(tcol & 7) << 1);
          const uint32_t off1 = off0 ^ 64u;  // column tcol + 32
          *reinterpret_cast<uint16_t*>(a_hi + off0) = hh;
          *reinterpret_cast<uint16_t*>(a_hi + off1) = 0;
          if (S::kAPlanes == 2) {
            *reinterpret_cast<uint16_t*>(a_lo + off0) = ll;
            *reinterpret_cast<uint16_t*>(a_lo + off1) = 0;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // ---- epilogue of the PREVIOUS tile (complete by now), overlapping the MMAs that drain this tile's last stages
      if (prev_tile >= 0) {
        const int pt = titer - 1;
        const int as = pt & 1;
        gemm_epilogue_tile<kBlockN>(gp, sk, prev_tile * GEMM_BLOCK_M, 0, warp, lane, tfull_bar(as), (uint32_t)(pt >> 1) & 1u,
                                    tmem_base + (uint32_t)(as * kBlockN));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) {
      const int pt = titer - 1;
      const int as = pt & 1;
      gemm_epilogue_tile<kBlockN>(gp, sk, prev_tile * GEMM_BLOCK_M, 0, warp, lane, tfull_bar(as), (uint32_t)(pt >> 1) & 1u,
                                  tmem_base + (uint32_t)(as * kBlockN));
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

template <int kPasses>
int launch_sphere_v4(const CUtensorMap& tmB, const CUtensorMap& tmB2, const GemmParams& gp, const GemmSinks& sk,
                     const SphereV4In& si, cudaStream_t st) {
  using S = SphereV4Smem<kPasses>;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPGAN_CUDA(cudaGetDevice(&dev), "spgan_sphere_conv_gemm");
  if (dev < 64 && !attr_set[dev]) {
    SPGAN_CUDA(cudaFuncSetAttribute(sphere_gemm_v4_kernel<kPasses>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
               "spgan_sphere_conv_gemm (shared memory opt-in)");
    attr_set[dev] = true;
  }
  const int grid = gp.m_tiles < SPGAN_NUM_SMS ? gp.m_tiles : SPGAN_NUM_SMS;
  sphere_gemm_v4_kernel<kPasses><<<grid, SPH_THREADS, S::kTotal, st>>>(tmB, tmB2, gp, sk, si);
  SPGAN_CHECK_LAUNCH("spgan_sphere_conv_gemm");
  spgan_internal_count_gemm_launch();
  return 0;
}

template <int kPasses, int kBlockN>
int launch_sphere(const CUtensorMap& tmB, const GemmParams& gp, const GemmSinks& sk, const SphereIn& si, cudaStream_t st) {
  using S = SphereSmem<kPasses, kBlockN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  SPGAN_CUDA(cudaGetDevice(&dev), "spgan_sphere_conv_gemm");
  if (dev < 64 && !attr_set[dev]) {
    SPGAN_CUDA(cudaFuncSetAttribute(sphere_gemm_kernel<kPasses, kBlockN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal),
               "spgan_sphere_conv_gemm (shared memory opt-in)");
    attr_set[dev] = true;
  }
  const int tiles = gp.m_tiles * gp.n_tiles;
  const int grid = tiles < SPGAN_NUM_SMS ? tiles : SPGAN_NUM_SMS;
  sphere_gemm_kernel<kPasses, kBlockN><<<grid, SPH_THREADS, S::kTotal, st>>>(tmB, gp, sk, si);
  SPGAN_CHECK_LAUNCH("spgan_sphere_conv_gemm");
  spgan_internal_count_gemm_launch();
  return 0;
}

}  // namespace

namespace {

int sphere_conv_gemm_v4(const SpganConvPass* p, const SpganSphereIn* in, const SpganGemmIO* io, void* stream) {
  const int nc = in->coords ? 3 : 0;
  const int Ct = in->C + nc;
  const int Cx = Ct - SPV_C;
  SPGAN_CHECK_ARG(in->C == SPV_C && in->Cp == SPV_C && Ct <= SPV_LD && Cx >= 0 && 9 * Cx <= 32,
                  "spgan_sphere_conv_gemm (repacked source): needs 256 features + at most 3 trailing channels, got %d + %d", in->C, nc);
  SPGAN_CHECK_ARG(io->kp == 9 * SPV_C, "spgan_sphere_conv_gemm (repacked source): the packed weight must have %d columns, got %d", 9 * SPV_C, io->kp);
  SPGAN_CHECK_ARG(io->kp2 == 64 && io->w2_packed != nullptr, "spgan_sphere_conv_gemm (repacked source): needs the 64-column tail weight (kp2 = 64)");
  const int64_t rows = (int64_t)p->B * p->H * p->W;
  if (rows == 0 || p->Cout == 0) return 0;
  SPGAN_CHECK_ARG(p->H * p->W >= GEMM_BLOCK_M, "spgan_sphere_conv_gemm: images of %dx%d pixels are smaller than one 128-row tile", p->H, p->W);
  SPGAN_CHECK_ARG(p->Cout >= 16 && p->Cout <= 256, "spgan_sphere_conv_gemm (repacked source): Cout=%d must be in [16, 256]", p->Cout);
  SPGAN_CHECK_ARG(p->My == p->H && p->Mx == p->W && p->out_stride == 1 && p->out_off_y == 0 && p->out_off_x == 0 &&
                  p->out_H == p->H && p->out_W == p->W, "spgan_sphere_conv_gemm: the output lattice is the input image");
  SPGAN_CHECK_ARG(in->grid && in->chan_map && io->w_packed, "spgan_sphere_conv_gemm: null pointer");
  SPGAN_CHECK_ARG(in->grid_group >= 1 && p->B % in->grid_group == 0, "spgan_sphere_conv_gemm: grid_group=%d must divide the batch %d", in->grid_group, p->B);
  SPGAN_CHECK_ARG(in->cmap_ld >= Ct, "spgan_sphere_conv_gemm: chan_map row stride %d < %d", in->cmap_ld, Ct);
  SPGAN_CHECK_ARG(io->y || io->y_packed || io->rgb_w, "spgan_sphere_conv_gemm: no output sink");
  SPGAN_CHECK_ARG((int64_t)p->H * p->W * SPV_LD < (1LL << 31) && rows < (1LL << 31) - 65536, "spgan_sphere_conv_gemm: input too large");
  SPGAN_CHECK_ARG((((uintptr_t)in->xg) & 15) == 0 && (((uintptr_t)in->grid) & 7) == 0 && (((uintptr_t)io->w_packed) & 15) == 0 &&
                  (((uintptr_t)io->w2_packed) & 15) == 0, "spgan_sphere_conv_gemm: misaligned pointer");
  const bool general = io->y == nullptr || io->y_layout != 0 || io->y_packed != nullptr || io->rgb_w != nullptr || io->residual_nhwc != nullptr;
  if (general) {
    SPGAN_CHECK_ARG(p->Cout % 32 == 0, "spgan_sphere_conv_gemm: channels-last / packed sinks need Cout %% 32 == 0, got %d", p->Cout);
    SPGAN_CHECK_ARG(io->residual == nullptr, "spgan_sphere_conv_gemm: an NCHW residual is only supported with a plain NCHW output");
  }
  if (io->y_packed) {
    SPGAN_CHECK_ARG(io->y_packed_cols >= p->Cout && io->y_packed_cols % 8 == 0 && (((uintptr_t)io->y_packed) & 15) == 0,
                    "spgan_sphere_conv_gemm: packed sink needs cols %% 8 == 0, cols >= Cout and a 16-byte aligned pointer");
    SPGAN_CHECK_ARG(io->y_packed_rows >= rows, "spgan_sphere_conv_gemm: packed sink has too few rows");
  }
  GemmParams gp = {};
  gp.B = p->B;
  gp.rows = (int32_t)rows;
  gp.Hl = p->H;
  gp.Wl = p->W;
  gp.My = p->H;
  gp.Mx = p->W;
  gp.Cout = p->Cout;
  gp.out_H = p->H;
  gp.out_W = p->W;
  gp.out_stride = 1;
  gp.out_cstride = p->out_cstride ? p->out_cstride : (int64_t)p->H * p->W;
  gp.ntaps = 1;
  gp.m_tiles = (int32_t)((rows + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  gp.n_tiles = 1;
  gp.out_scale = p->out_scale;
  gp.act = p->act;
  gp.act_alpha = p->act_alpha;
  gp.act_gain = p->act_gain;
  gp.a_f16 = io->fmt;
  gp.b_f16 = (int32_t)io->w_fmt;
  gp.y_nhwc = io->y_layout != 0 ? 1 : 0;
  gp.y_bstride = io->y_bstride ? io->y_bstride : (int64_t)p->H * p->W * p->Cout;
  gp.pk_rows = io->y_packed_rows;
  gp.pk_cols = io->y_packed_cols;
  gp.pk_f16 = io->y_packed_fmt;
  gp.rgb_n = io->rgb_w ? io->rgb_n : 0;
  gp.res_bstride = io->res_bstride ? io->res_bstride : (int64_t)p->H * p->W * p->Cout;
  GemmSinks sk = {};
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
  SphereV4In si;
  si.xg = in->xg;
  si.grid = in->grid;
  si.in_mul = in->in_mul;
  si.chan_map = in->chan_map;
  si.B = p->B;
  si.Ct = Ct;
  si.H = p->H;
  si.W = p->W;
  si.grid_group = in->grid_group;
  si.cmap_ld = in->cmap_ld;
  CUtensorMap tmB, tmB2;
  {
    const int kp = io->kp;
    cuuint64_t dims[4] = {(cuuint64_t)kp, (cuuint64_t)p->Cout, 1, 2};
    cuuint64_t strides[3] = {(cuuint64_t)kp * 2, (cuuint64_t)p->Cout * kp * 2, (cuuint64_t)p->Cout * kp * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, 256, 1, 1};
    if (int e = encode_bf16_map(&tmB, io->w_packed, 4, dims, strides, box, "spgan_sphere_conv_gemm (B map)")) return e;
  }
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)p->Cout, 1, 2};
    cuuint64_t strides[3] = {64 * 2, (cuuint64_t)p->Cout * 64 * 2, (cuuint64_t)p->Cout * 64 * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, 256, 1, 1};
    if (int e = encode_bf16_map(&tmB2, io->w2_packed, 4, dims, strides, box, "spgan_sphere_conv_gemm (tail B map)")) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (p->precision == 1) return launch_sphere_v4<3>(tmB, tmB2, gp, sk, si, st);
  if (p->precision == 3) return launch_sphere_v4<2>(tmB, tmB2, gp, sk, si, st);
  return launch_sphere_v4<1>(tmB, tmB2, gp, sk, si, st);
}

}  // namespace

extern "C" int spgan_sphere_conv_gemm(const SpganConvPass* p, const SpganSphereIn* in, const SpganGemmIO* io, void* stream) {
  SPGAN_CHECK_ARG(p != nullptr && in != nullptr && io != nullptr, "spgan_sphere_conv_gemm: null descriptor");
  SPGAN_CHECK_ARG(p->precision >= 1 && p->precision <= 3, "spgan_sphere_conv_gemm: precision must be 1, 2 or 3, got %d", p->precision);
  SPGAN_CHECK_ARG(io->fmt == (p->precision == 3 ? 1 : 0) && io->w_fmt == io->fmt,
                  "spgan_sphere_conv_gemm: operand formats (A %d, W %d) do not match precision %d", io->fmt, (int)io->w_fmt, p->precision);
  if (in->xg != nullptr) return sphere_conv_gemm_v4(p, in, io, stream);
  SPGAN_CHECK_ARG(p->B >= 0 && p->H >= 0 && p->W >= 0 && p->Cout >= 0 && in->C >= 0, "spgan_sphere_conv_gemm: negative size");
  const int nc = in->coords ? 3 : 0;
  SPGAN_CHECK_ARG(in->Cp >= in->C + nc && in->Cp % 64 == 0, "spgan_sphere_conv_gemm: Cp=%d must be a multiple of 64 and >= %d", in->Cp, in->C + nc);
  SPGAN_CHECK_ARG(io->kp == 9 * in->Cp, "spgan_sphere_conv_gemm: the packed weight must have 9*Cp = %d columns, got %d", 9 * in->Cp, io->kp);
  const int64_t rows = (int64_t)p->B * p->H * p->W;
  if (rows == 0 || p->Cout == 0) return 0;
  SPGAN_CHECK_ARG(p->H * p->W >= GEMM_BLOCK_M, "spgan_sphere_conv_gemm: images of %dx%d pixels are smaller than one 128-row tile; "
                  "use spgan_sphere_pack + spgan_conv_gemm", p->H, p->W);
  SPGAN_CHECK_ARG(p->Cout >= 16, "spgan_sphere_conv_gemm: Cout=%d < 16 belongs on the SIMT path", p->Cout);
  SPGAN_CHECK_ARG(p->My == p->H && p->Mx == p->W && p->out_stride == 1 && p->out_off_y == 0 && p->out_off_x == 0 &&
                  p->out_H == p->H && p->out_W == p->W, "spgan_sphere_conv_gemm: the output lattice is the input image");
  SPGAN_CHECK_ARG(in->x_nhwc && in->grid && in->chan_map && io->w_packed, "spgan_sphere_conv_gemm: null pointer");
  SPGAN_CHECK_ARG(io->y || io->y_packed || io->rgb_w, "spgan_sphere_conv_gemm: no output sink");
  SPGAN_CHECK_ARG(p->B <= 65535 && in->C <= 32767, "spgan_sphere_conv_gemm: B=%d / C=%d exceed the channel-map encoding", p->B, in->C);
  SPGAN_CHECK_ARG(rows * (in->C > 3 ? in->C : 3) < (1LL << 31), "spgan_sphere_conv_gemm: input too large for 32-bit plane offsets");
  SPGAN_CHECK_ARG((((uintptr_t)in->grid) & 7) == 0 && (((uintptr_t)io->w_packed) & 15) == 0, "spgan_sphere_conv_gemm: misaligned grid / weight pointer");
  const bool general = io->y == nullptr || io->y_layout != 0 || io->y_packed != nullptr || io->rgb_w != nullptr || io->residual_nhwc != nullptr;
  if (general) {
    SPGAN_CHECK_ARG(p->Cout % 32 == 0, "spgan_sphere_conv_gemm: channels-last / packed sinks need Cout %% 32 == 0, got %d", p->Cout);
    SPGAN_CHECK_ARG(io->residual == nullptr, "spgan_sphere_conv_gemm: an NCHW residual is only supported with a plain NCHW output");
  }
  if (io->y_packed) {
    SPGAN_CHECK_ARG(io->y_packed_cols >= p->Cout && io->y_packed_cols % 8 == 0 && (((uintptr_t)io->y_packed) & 15) == 0,
                    "spgan_sphere_conv_gemm: packed sink needs cols %% 8 == 0, cols >= Cout and a 16-byte aligned pointer");
    SPGAN_CHECK_ARG(io->y_packed_rows >= rows, "spgan_sphere_conv_gemm: packed sink has too few rows");
  }

  GemmParams gp = {};
  gp.B = p->B;
  gp.rows = (int32_t)rows;
  gp.Hl = p->H;
  gp.Wl = p->W;
  gp.My = p->H;
  gp.Mx = p->W;
  gp.Cout = p->Cout;
  gp.out_H = p->H;
  gp.out_W = p->W;
  gp.out_stride = 1;
  gp.out_cstride = p->out_cstride ? p->out_cstride : (int64_t)p->H * p->W;
  gp.ntaps = 1;
  gp.m_tiles = (int32_t)((rows + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M);
  // the A tile is PRODUCED (not fetched) once per N tile: one 256-wide tile whenever Cout allows it
  const int block_n = p->Cout <= 128 ? 128 : 256;
  gp.n_tiles = (p->Cout + block_n - 1) / block_n;
  gp.out_scale = p->out_scale;
  gp.act = p->act;
  gp.act_alpha = p->act_alpha;
  gp.act_gain = p->act_gain;
  gp.a_f16 = io->fmt;
  gp.b_f16 = (int32_t)io->w_fmt;
  gp.y_nhwc = io->y_layout != 0 ? 1 : 0;
  gp.y_bstride = io->y_bstride ? io->y_bstride : (int64_t)p->H * p->W * p->Cout;
  gp.pk_rows = io->y_packed_rows;
  gp.pk_cols = io->y_packed_cols;
  gp.pk_f16 = io->y_packed_fmt;
  gp.rgb_n = io->rgb_w ? io->rgb_n : 0;
  gp.res_bstride = io->res_bstride ? io->res_bstride : (int64_t)p->H * p->W * p->Cout;
  GemmSinks sk = {};
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
  SphereIn si;
  si.xh = in->x_nhwc;
  si.coords = in->coords;
  si.grid = in->grid;
  si.in_mul = in->in_mul;
  si.chan_map = in->chan_map;
  si.B = p->B;
  si.C = in->C;
  si.nc = nc;
  si.H = p->H;
  si.W = p->W;
  si.Cp = in->Cp;

  CUtensorMap tmB;
  {
    const int kp = io->kp;
    cuuint64_t dims[4] = {(cuuint64_t)kp, (cuuint64_t)p->Cout, 1, 2};
    cuuint64_t strides[3] = {(cuuint64_t)kp * 2, (cuuint64_t)p->Cout * kp * 2, (cuuint64_t)p->Cout * kp * 2};
    cuuint32_t box[4] = {GEMM_BLOCK_K, (cuuint32_t)block_n, 1, 1};
    if (int e = encode_bf16_map(&tmB, io->w_packed, 4, dims, strides, box, "spgan_sphere_conv_gemm (B map)")) return e;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (p->precision == 1)
    return block_n == 256 ? launch_sphere<3, 256>(tmB, gp, sk, si, st) : launch_sphere<3, 128>(tmB, gp, sk, si, st);
  if (p->precision == 3)
    return block_n == 256 ? launch_sphere<2, 256>(tmB, gp, sk, si, st) : launch_sphere<2, 128>(tmB, gp, sk, si, st);
  return block_n == 256 ? launch_sphere<1, 256>(tmB, gp, sk, si, st) : launch_sphere<1, 128>(tmB, gp, sk, si, st);
}

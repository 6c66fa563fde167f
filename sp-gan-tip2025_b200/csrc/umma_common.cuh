// tcgen05 / TMEM / TMA building blocks shared by the forward implicit-GEMM (conv_umma.cu) and the weight-gradient GEMM
// (conv_wgrad_umma.cu): mbarrier and TMA wrappers, UMMA shared-memory / instruction descriptors, TMEM loads and the
// host-side tensor-map encoder.  sm_100a only.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_N = 256;
constexpr int GEMM_BLOCK_K = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;  // 16 KiB
constexpr int B_TILE_BYTES = GEMM_BLOCK_N * GEMM_BLOCK_K * 2;  // 32 KiB

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (an error code at the C ABI), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (mbar_try_wait(bar, parity)) return;
    if ((it & 0x3ff) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000ull) {  // 4 s
        printf("spgan conv_gemm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
               threadIdx.x, bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// im2col-mode tile load from a rank-4 (C, W, H, N) tensor map: `pixelsPerColumn` base pixels starting at (w, h, n),
// walking w, then h, then n inside the map's bounding box, each displaced by the filter offset (w_off, h_off).
// Verified on a B200 by tools/probes/tma_im2col.cu (row / image crossing, channel offsets, zero fill past the tensor).
__device__ __forceinline__ void tma_load_im2col(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                                int n, int w_off, int h_off) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], "
      "{%7, %8};" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)w_off), "h"((uint16_t)h_off)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle, 8-row core-matrix groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(int n) {
  return (1u << 4)                  // D format f32
         | (1u << 7)                // A format bf16
         | (1u << 10)               // B format bf16
         | ((uint32_t)(n >> 3) << 17)
         | ((uint32_t)(GEMM_BLOCK_M >> 4) << 24);  // A, B K-major: bits 15, 16 stay 0
}
// kind::f16 descriptor; each operand is fp16 (format code 0) or bf16 (format code 1), independently.
__device__ __forceinline__ uint32_t umma_idesc_16(int n, bool a_f16, bool b_f16) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(GEMM_BLOCK_M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// 16-bit operand split, raw bits: kF16 = false -> bf16 hi/lo (8 + 8 mantissa bits, fp32 range), kF16 = true -> fp16 hi/lo
// (11 + 11 bits; values beyond +-65504 saturate, which no activation or weight of this network approaches).
template <bool kF16>
__device__ __forceinline__ void split16(float v, uint16_t& hi, uint16_t& lo) {
  if (kF16) {
    v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  }
}
__device__ __forceinline__ uint32_t pack2x16(uint16_t a, uint16_t b) { return (uint32_t)a | ((uint32_t)b << 16); }

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = (EncodeTiledFn)sym;
  return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

inline EncodeIm2colFn get_encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = (EncodeIm2colFn)sym;
  return fn;
}

// NHWC bf16 tensor (C, W, H, N) with base pixels restricted to [0, W + upper_w) x [0, H + upper_h) (upper_* <= 0).
inline int encode_bf16_im2col_map(CUtensorMap* map, const void* base, cuuint64_t C, cuuint64_t W, cuuint64_t H, cuuint64_t N,
                                  int upper_w, int upper_h, cuuint32_t channels, cuuint32_t pixels, const char* who) {
  EncodeIm2colFn fn = get_encode_im2col_fn();
  SPGAN_CHECK_ARG(fn != nullptr, "%s: cuTensorMapEncodeIm2col is not available from the CUDA driver", who);
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
  int lower[2] = {0, 0};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, lower, upper, channels,
                  pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPGAN_CHECK_ARG(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeIm2col failed with CUresult %d", who, (int)r);
  return 0;
}

int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box, const char* who) {
  EncodeTiledFn fn = get_encode_fn();
  SPGAN_CHECK_ARG(fn != nullptr, "%s: cuTensorMapEncodeTiled is not available from the CUDA driver", who);
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPGAN_CHECK_ARG(r == CUDA_SUCCESS, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return 0;
}

}  // namespace

// Shared-memory staging for the HBM-bound stencil / gather kernels: a persistent CTA walks its work items with a two-stage
// ring; the contiguous source chunk of item k + 1 is fetched by ONE bulk-async copy (cp.async.bulk, the 1-D TMA path,
// SASS UBLKCP) that completes on an mbarrier while the threads compute item k out of the other stage.  Unlike a
// load -> __syncthreads -> compute -> store CTA, the loads of the next item are in flight during the whole compute phase,
// and they cost one instruction of one thread instead of ~8 LDG + 8 STS per thread.
//
// cp.async.bulk needs 16-byte aligned addresses and sizes; the planes of this path are odd-sized (105 x 105 floats), so a
// chunk starts at an arbitrary float.  The copy therefore starts at the chunk address rounded DOWN to 16 bytes (the staged
// data then begins `shift` floats into the stage) and ends at the chunk end rounded UP — never beyond `limit`, the 16-byte
// floor of the end of the tensor: the at most 3 floats that are left over at the very end of the tensor are copied by the
// issuing thread with plain loads.  Rounding down never leaves the tensor because its base is 16-byte aligned (checked by
// the callers).
#pragma once
#include "common.cuh"

namespace stream_stage {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_bar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool bar_try_wait(uint32_t bar, uint32_t parity) {
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
// Bounded wait: a protocol bug surfaces as a trapped kernel (an error code at the C ABI), never as a hung GPU.
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  if (bar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (uint32_t it = 1;; ++it) {
    if (bar_try_wait(bar, parity)) return;
    if ((it & 0x3ff) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (now - t0 > 2000000000ull) {  // 2 s
        printf("spgan stream stage: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x,
               parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Float offset inside a stage at which a chunk starting at `src` begins.
__device__ __forceinline__ int chunk_shift(const float* src) { return (int)(((uintptr_t)src & 15) >> 2); }

// Issued by ONE thread: fetch `n` floats starting at `src` into `stage` (16-byte aligned shared memory with room for
// n + 8 floats) so that element e lands at stage[chunk_shift(src) + e], and arrive on `bar` (initialised with count 1).
__device__ __forceinline__ void issue_chunk(float* stage, const float* src, int n, uintptr_t limit, uint64_t* bar) {
  const uintptr_t a = (uintptr_t)src;
  const uintptr_t a0 = a & ~(uintptr_t)15;
  uintptr_t a1 = (a + (uintptr_t)n * 4 + 15) & ~(uintptr_t)15;
  if (a1 > limit) a1 = limit;
  const uint32_t b = smem_addr(bar);
  if (a1 > a0) {
    // tail beyond the last whole 16 bytes of the tensor: plain copies, published by the arrive below
    const int done = (int)((a1 - a) >> 2);
    const int shift = (int)((a - a0) >> 2);
    for (int e = done; e < n; ++e) stage[shift + e] = __ldg(src + e);
    const uint32_t bytes = (uint32_t)(a1 - a0);
    bar_arrive_expect_tx(b, bytes);
    bulk_g2s(smem_addr(stage), (const void*)a0, bytes, b);
  } else {
    const int shift = (int)((a - a0) >> 2);
    for (int e = 0; e < n; ++e) stage[shift + e] = __ldg(src + e);
    bar_arrive(b);
  }
}

}  // namespace stream_stage

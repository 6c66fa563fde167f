// Shared helpers for libspgan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/spgan_b200.h"

// SM count of the current device (B200: 2 dies x 74 = 148); grids are sized in multiples of this.  Queried once per device
// (the library is used one process per GPU, but nothing here assumes device 0).
static inline int spgan_num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cache[dev];
}
#define SPGAN_NUM_SMS spgan_num_sms()

void spgan_set_error(const char* fmt, ...);
void spgan_internal_count_gemm_launch();  // conv_umma.cu: bumps the counter behind spgan_gemm_launch_count()

#define SPGAN_CHECK_ARG(cond, ...)          \
  do {                                      \
    if (!(cond)) {                          \
      spgan_set_error(__VA_ARGS__);         \
      return 1;                             \
    }                                       \
  } while (0)

#define SPGAN_CHECK_LAUNCH(name)                                                   \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      spgan_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

#define SPGAN_CUDA(call, name)                                            \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) {                                             \
      spgan_set_error("%s: %s", name, cudaGetErrorString(e__));           \
      return 2;                                                           \
    }                                                                     \
  } while (0)

// Diagnostics: SPGAN_LEGACY_HBM_KERNELS=1 routes the FIR / gather entry points to the pre-streaming kernels (A/B timing in
// tools/microbench.py and cross-checks in the tests).  Read per call so a process can toggle it.
static inline bool spgan_legacy_hbm() {
  const char* e = getenv("SPGAN_LEGACY_HBM_KERNELS");
  return e != nullptr && e[0] == '1';
}

// Opt a kernel into more than 48 KB of dynamic shared memory, once per device (the attribute is per context; the library is
// used one process per GPU, but nothing here assumes it).
template <class Kernel>
static inline bool spgan_allow_smem(Kernel kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  if (done[dev]) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) return false;
  done[dev] = true;
  return true;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid size for a grid-stride kernel: enough CTAs for `work` items at `per_block` each, capped at
// `waves` full waves of the 148 SMs x `ctas_per_sm` resident CTAs.
static inline int grid_for(int64_t work, int64_t per_block, int ctas_per_sm, int waves = 4) {
  int64_t need = ceil_div64(work, per_block);
  int64_t cap = (int64_t)SPGAN_NUM_SMS * ctas_per_sm * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Exact unsigned division by a runtime-constant divisor (Granlund-Montgomery), so that flat-index kernels can decode
// (row, column) without the ~25-instruction hardware-emulated integer division.
struct FastDiv {
  uint32_t d, m, s1, s2;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l > 0 ? l - 1 : 0;
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}
#endif

// Probe: TMA im2col mode on an NHWC bf16 tensor with 128-byte swizzle.  Loads `PIX` consecutive base pixels (traversing
// w, then h, then n inside the bounding box) at filter offset (w_off, h_off) and compares with the expected gather.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int PIX = 128, CH = 64;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int w_off, int h_off, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb0 = (uint32_t)__cvta_generic_to_shared(smem);
  uint32_t sb = (sb0 + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(PIX * CH * 2) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(sb), "l"(&tm), "r"(b), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)w_off), "h"((uint16_t)h_off)
        : "memory");
    uint32_t ok = 0;
    for (int it = 0; it < 50000000 && !ok; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    const __nv_bfloat16* t = reinterpret_cast<const __nv_bfloat16*>(smem + (sb - sb0));
    for (int r = 0; r < PIX; ++r)
      for (int k = 0; k < CH; ++k) {
        int chunk = (k / 8) ^ (r % 8);
        out[r * CH + k] = ok ? __bfloat162float(t[r * CH + chunk * 8 + (k % 8)]) : -1.f;
      }
  }
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int N = 3, H = 11, W = 13, C = 80, K = 7;  // 7x7 filter, no padding: output 5 x 7
  const int OH = H - K + 1, OW = W - K + 1;
  std::vector<__nv_bfloat16> hbuf((size_t)N * H * W * C);
  auto val = [&](int n, int y, int x, int c) { return (float)((n * 131 + y * 31 + x * 7 + c) % 251); };
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < C; ++c) hbuf[(((size_t)n * H + y) * W + x) * C + c] = __float2bfloat16(val(n, y, x, c));
  __nv_bfloat16* d;
  cudaMalloc(&d, hbuf.size() * 2);
  cudaMemcpy(d, hbuf.data(), hbuf.size() * 2, cudaMemcpyHostToDevice);
  float* out;
  cudaMalloc(&out, PIX * CH * 4);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &sym, cudaEnableDefault, &qres);
  if (!sym) { printf("no cuTensorMapEncodeIm2col\n"); return 0; }
  EncodeIm2colFn fn = (EncodeIm2colFn)sym;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lower[2] = {0, 0};
  int upper[2] = {-(K - 1), -(K - 1)};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, lower, upper, CH, PIX, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode im2col -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 0;
  struct Case { int c, m0, wo, ho; };
  Case cases[] = {{0, 0, 0, 0}, {0, 0, 3, 5}, {64, 0, 6, 6}, {0, 20, 2, 1}, {16, 9, 1, 0}};
  for (const Case& cs : cases) {
    // base pixel of output index m0 in (n, oy, ox) order
    int n = cs.m0 / (OH * OW), rem = cs.m0 % (OH * OW), oy = rem / OW, ox = rem % OW;
    cudaMemset(out, 0, PIX * CH * 4);
    probe<<<1, 32, PIX * CH * 2 + 2048>>>(tm, cs.c, ox, oy, n, cs.wo, cs.ho, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  case c=%d m0=%d off=(%d,%d): CUDA error %s\n", cs.c, cs.m0, cs.wo, cs.ho, cudaGetErrorString(e)); return 0; }
    std::vector<float> ho(PIX * CH);
    cudaMemcpy(ho.data(), out, PIX * CH * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first_bad = -1;
    for (int rr = 0; rr < PIX; ++rr) {
      int m = cs.m0 + rr;
      int nn = m / (OH * OW), rm = m % (OH * OW), yy = rm / OW, xx = rm % OW;
      for (int k = 0; k < CH; ++k) {
        int ch = cs.c + k;
        float want = (nn < N && ch < C) ? val(nn, yy + cs.ho, xx + cs.wo, ch) : 0.f;
        if (ho[rr * CH + k] != want) { if (first_bad < 0) first_bad = rr * CH + k; ++bad; }
      }
    }
    printf("  case c=%d m0=%d off=(%d,%d): %d mismatches", cs.c, cs.m0, cs.wo, cs.ho, bad);
    if (bad) printf(" (first at row %d col %d: got %g)", first_bad / CH, first_bad % CH, ho[first_bad]);
    printf("\n");
  }
  return 0;
}

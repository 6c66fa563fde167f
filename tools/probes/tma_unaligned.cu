// Probe: does a 128B-swizzled TMA tile load accept an inner coordinate that is not a multiple of 8 bf16 elements (16 B),
// and an inner extent that is not a multiple of 8?  Prints the first elements of each loaded row (un-swizzled).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
  sb = (sb + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(64 * 8 * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sb),
                 "l"(&tm), "r"(b), "r"(c0), "r"(c1)
                 : "memory");
    uint32_t ok = 0;
    for (int it = 0; it < 100000000 && !ok; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
    const __nv_bfloat16* t = reinterpret_cast<const __nv_bfloat16*>(smem + (sb - (uint32_t)__cvta_generic_to_shared(smem)));
    for (int r = 0; r < 8; ++r)
      for (int k = 0; k < 64; ++k) {
        // 128B swizzle: 16-byte chunk index XOR (row % 8)
        int chunk = (k / 8) ^ (r % 8);
        out[r * 64 + k] = ok ? __bfloat162float(t[r * 64 + chunk * 8 + (k % 8)]) : -1.f;
      }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int Q = 1083, QS = 1088, R = 40;
  std::vector<__nv_bfloat16> h((size_t)R * QS);
  for (int r = 0; r < R; ++r)
    for (int q = 0; q < QS; ++q) h[(size_t)r * QS + q] = __float2bfloat16((float)(r * 2000 + q) / 4.f == 0 ? 0.f : (float)((r * 37 + q) % 251));
  __nv_bfloat16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  float* out;
  cudaMalloc(&out, 8 * 64 * 4);
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
  EncodeTiledFn fn = (EncodeTiledFn)sym;
  for (int extent : {QS, Q}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)extent, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)QS * 2};
    cuuint32_t box[2] = {64, 8};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("extent %d: encode -> %d\n", extent, (int)r);
    if (r != CUDA_SUCCESS) continue;
    for (int c0 : {0, 64, 8, 1, 3, 21, 1056}) {
      cudaMemset(out, 0, 8 * 64 * 4);
      probe<<<1, 32, 8 * 64 * 2 + 2048>>>(tm, c0, 2, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("  c0=%d: CUDA error %s\n", c0, cudaGetErrorString(e));
        return 0;
      }
      float ho[8 * 64];
      cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r2 = 0; r2 < 8; ++r2)
        for (int k = 0; k < 64; ++k) {
          int q = c0 + k;
          float want = q < extent ? (float)(((r2 + 2) * 37 + q) % 251) : 0.f;
          if (ho[r2 * 64 + k] != want) ++bad;
        }
      printf("  c0=%d: %d mismatches (first row: %g %g %g %g)\n", c0, bad, ho[0], ho[1], ho[2], ho[3]);
    }
  }
  return 0;
}

// TMA request-rate microbenchmark (B200): bytes/clk/SM of cp.async.bulk.tensor loads as a function of the box's inner
// extent.  (a) 5-D "pixel group" map of il_stream.cuh: inner dim = 8 px = 16 bytes, (b) 4-D planar map: inner dim = a whole
// row of W px.  One CTA per SM, 4 boxes in flight, 16-bit [N][C][H][W] tensor with C = 32, H = W = 224.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
template <int RANK>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int iters, int box_bytes, int C, int H, long long* cycles) {
  extern __shared__ __align__(128) uint8_t sm[];
  const uint32_t bar0 = smem_u32(sm), buf0 = smem_u32(sm + 128);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0 + 8 * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const long long t0 = clock64();
    const int n = blockIdx.x % 8;
    for (int it = 0; it < iters + 4; ++it) {
      const int s = it & 3;
      if (it >= 4) wait(bar0 + 8 * s, ((it >> 2) - 1) & 1);
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar0 + 8 * s), "r"(box_bytes) : "memory");
        const int y = (it * 4) % (H - 4);
        if (RANK == 5)
          asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
                       ::"r"(buf0 + s * 49152), "l"(&tm), "r"(bar0 + 8 * s), "r"(0), "r"(0), "r"(0), "r"(y), "r"(n) : "memory");
        else
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                       ::"r"(buf0 + s * 49152), "l"(&tm), "r"(bar0 + 8 * s), "r"(0), "r"(y), "r"(0), "r"(n) : "memory");
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn fn = (EncodeTiledFn)fp;
  const int N = 8, C = 32, H = 224, W = 224;
  uint16_t* d; cudaMalloc(&d, (size_t)N * C * H * W * 2); cudaMemset(d, 0, (size_t)N * C * H * W * 2);
  long long* dc; cudaMalloc(&dc, 148 * 8);
  const int iters = 400;
  for (int mode = 0; mode < 3; ++mode) {
    CUtensorMap tm; cuuint32_t es[5] = {1, 1, 1, 1, 1}; int box_bytes;
    if (mode == 0) {          // 5-D pixel-group map, box (8, 18, 28, 4, 1): 2016 inner rows of 16 B
      cuuint64_t dims[5] = {8, (cuuint64_t)C, W / 8, H, N}; cuuint64_t st[4] = {(cuuint64_t)H * W * 2, 16, W * 2, (cuuint64_t)C * H * W * 2};
      cuuint32_t box[5] = {8, 18, 28, 4, 1}; box_bytes = 8 * 18 * 28 * 4 * 2;
      fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {                  // 4-D planar map, box (224 or 64, 4, 18, 1): inner rows of 448 / 128 B
      const int bw = mode == 1 ? 224 : 64;
      cuuint64_t dims[4] = {W, H, (cuuint64_t)C, N}; cuuint64_t st[3] = {W * 2, (cuuint64_t)H * W * 2, (cuuint64_t)C * H * W * 2};
      cuuint32_t box[4] = {(cuuint32_t)bw, 4, 18, 1}; box_bytes = bw * 4 * 18 * 2;
      fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (mode == 0) k<5><<<148, 128, 200 * 1024>>>(tm, iters, box_bytes, C, H, dc); else k<4><<<148, 128, 200 * 1024>>>(tm, iters, box_bytes, C, H, dc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, dc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i] / 148.0;
    const int rows = mode == 0 ? 2016 : 72 * (mode == 1 ? 1 : 1);
    printf("mode %d (%s): %s  %.0f cycles per box of %d bytes -> %.2f bytes/clk/SM, %.1f cycles per inner row (%d rows/box)\n", mode,
           mode == 0 ? "5-D groups, 16-byte inner rows" : (mode == 1 ? "4-D planar, 448-byte rows" : "4-D planar, 128-byte rows"), cudaGetErrorString(e),
           avg / iters, box_bytes, box_bytes / (avg / iters), avg / iters / rows, rows);
  }
  return 0;
}

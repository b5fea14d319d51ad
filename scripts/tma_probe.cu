// Standalone probe: which form of cp.async.bulk.tensor (descriptor location / rank) works on this box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__device__ void probe_body(const CUtensorMap* tm, uint16_t* out, int nbytes, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
  uint16_t* dst = reinterpret_cast<uint16_t*>(smem + 128);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(nbytes) : "memory");
    if (RANK == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n"
                   ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                   ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(mbar)) : "memory");
    if (++spins > (1u << 22)) { if (threadIdx.x == 0) printf("timeout\n"); return; }
  }
  for (int i = threadIdx.x; i < nbytes / 2; i += blockDim.x) out[i] = dst[i];
}

__global__ void probe_param4(const __grid_constant__ CUtensorMap tm, uint16_t* out, int nbytes, int c0, int c1, int c2, int c3) {
  probe_body<4>(&tm, out, nbytes, c0, c1, c2, c3);
}
__global__ void probe_global4(const CUtensorMap* tm, uint16_t* out, int nbytes, int c0, int c1, int c2, int c3) {
  probe_body<4>(tm, out, nbytes, c0, c1, c2, c3);
}
__global__ void probe_param3(const __grid_constant__ CUtensorMap tm, uint16_t* out, int nbytes, int c0, int c1, int c2) {
  probe_body<3>(&tm, out, nbytes, c0, c1, c2, 0);
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "param4";
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn fn = (EncodeTiledFn)p;
  const int N = 2, C = 3, H = 16, W = 32, bw = 16, bh = 9, bc = 3;
  std::vector<uint16_t> h(N * C * H * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint16_t)(i + 1);
  uint16_t *d, *o; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMalloc(&o, 65536));
  CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(o, 0xff, 65536));
  const int nbytes = bw * bh * bc * 2;
  const bool pos = strstr(mode, "pos") != nullptr;
  int cx = pos ? 8 : -4, cy = pos ? 2 : -4;          // or explicit: tma_probe <mode> <cx> <cy>
  const int n = 1;
  if (argc > 3) { cx = atoi(argv[2]); cy = atoi(argv[3]); }
  CUtensorMap tm;
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r;
  if (strstr(mode, "3")) {
    cuuint64_t dims[3] = {W, H, (cuuint64_t)C * N}; cuuint64_t str[2] = {W * 2, H * W * 2};
    cuuint32_t box[3] = {bw, bh, bc};
    r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {W, H, C, N}; cuuint64_t str[3] = {W * 2, H * W * 2, (cuuint64_t)C * H * W * 2};
    cuuint32_t box[4] = {bw, bh, bc, 1};
    r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, strstr(mode, "nol2") ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("[%s] encode -> %d, sizeof(CUtensorMap)=%zu alignof=%zu\n", mode, (int)r, sizeof tm, alignof(CUtensorMap));
  const size_t smem = 128 + nbytes + 128;
  if (strstr(mode, "global")) {
    CUtensorMap* dtm; CK(cudaMalloc(&dtm, sizeof tm)); CK(cudaMemcpy(dtm, &tm, sizeof tm, cudaMemcpyHostToDevice));
    probe_global4<<<1, 128, smem>>>(dtm, o, nbytes, cx, cy, 0, n);
  } else if (strstr(mode, "3")) {
    probe_param3<<<1, 128, smem>>>(tm, o, nbytes, cx, cy, n * C);
  } else {
    probe_param4<<<1, 128, smem>>>(tm, o, nbytes, cx, cy, 0, n);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel: %s\n", mode, cudaGetErrorString(e)); return 2; }
  std::vector<uint16_t> res(nbytes / 2);
  cudaMemcpy(res.data(), o, nbytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int c = 0; c < bc; ++c) for (int y = 0; y < bh; ++y) for (int x = 0; x < bw; ++x) {
    int gy = cy + y, gx = cx + x;
    uint16_t want = (gy < 0 || gy >= H || gx < 0 || gx >= W) ? 0 : h[((n * C + c) * H + gy) * W + gx];
    if (res[(c * bh + y) * bw + x] != want) ++bad;
  }
  printf("[%s] %s (%d mismatches of %d)\n", mode, bad ? "WRONG" : "ok", bad, nbytes / 2);
  return bad ? 3 : 0;
}

// Which feature caps CTAs/SM at 1?  Reports cudaOccupancyMaxActiveBlocksPerMultiprocessor for small variants.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE, int LB>
__global__ void __launch_bounds__(LB, 1) k(float* out, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ uint8_t sm[];
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm);
  if (MODE & 1) {
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(64) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(*slot), "r"(64) : "memory");
  }
  if (MODE & 2) {
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 64);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(256) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                   ::"r"(smem_u32(sm + 128)), "l"(&tm), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    }
  }
  if (MODE & 4) {
    if (threadIdx.x == 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(sm + 64)) : "memory");
  }
  out[threadIdx.x] = (float)sm[threadIdx.x];
}
template <int MODE, int LB>
void rep(const char* name) {
  cudaFuncSetAttribute(k<MODE, LB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k<MODE, LB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  int o1 = -1, o2 = -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k<MODE, LB>, 256, 60000);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k<MODE, LB>, 352, 106496);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<MODE, LB>);
  printf("%-40s regs %d: occupancy %d (256 thr, 60 KB)  %d (352 thr, 104 KB)\n", name, fa.numRegs, o1, o2);
}
int main() {
  rep<0, 768>("plain, launch_bounds(768,1)");
  rep<0, 384>("plain, launch_bounds(384,1)");
  rep<1, 768>("tcgen05.alloc");
  rep<2, 768>("TMA + mbarrier");
  rep<4, 768>("tcgen05.commit");
  rep<7, 768>("all");
  rep<7, 384>("all, launch_bounds(384,1)");
  return 0;
}

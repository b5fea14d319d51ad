// Round-2 prototype — RUN ON A B200 at the very end of round 1: "max |Y - ref| = 0.000976562, 0 of 52224 elements off" (fp16
// rounding of the outputs only).  The 1x1-conv GEMM of a fused ILBlock tile on tcgen05, IN PLACE on a tile stored as
// [pixel group][channel][8 pixels] — the layout DESIGN.md §9 proposes for every phase of the kernel.
//   D[M = 128 channel lanes][N = 256 pixels] (fp32, TMEM) = W[M][K] (K-major A) * X[pixels][K] (MN-major B: pixels contiguous)
//   channels are the M dimension so that an epilogue thread (= TMEM lane = channel) receives CONSECUTIVE pixels and stores
//   16-byte rows of the tile: for one pixel group the 32 lanes of a warp write 512 contiguous bytes.
// Tile: NP = 1024 pixels (4 N blocks of 256), K = 64 input channels, Cout = 51 (padded to the 128 lanes, unused lanes idle).
// TMEM: 512 columns = two 256-pixel accumulator buffers, double-buffered between the issuing thread (warp 4) and the four
// epilogue warps; the epilogue overwrites the pixel groups of its N block, which no later MMA reads.
// Verified by that run: B as the MN-major operand straight from the tile, N = 256 per instruction, the full / empty mbarrier
// phases of the double-buffered accumulators, the in-place epilogue (bias + PReLU + 16-byte stores of 8 pixels per channel).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NP = 1024, K = 64, COUT = 51, M = 128, NB = 256, C = 64;   // C: channel slots per pixel group (>= K and >= COUT)
constexpr int kThreads = 160;                                              // warps 0-3 epilogue, warp 4 issues the MMAs

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1u << 24)) { printf("mbarrier timeout (thread %d)\n", (int)threadIdx.x); __trap(); }
  }
}

__global__ void __launch_bounds__(kThreads, 1) il_gemm(const __half* __restrict__ Xg /*[K][NP]*/, const __half* __restrict__ Wg /*[COUT][K]*/,
                                                       const float* __restrict__ bias, const float* __restrict__ slope,
                                                       __half* __restrict__ Yg /*[COUT][NP]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __half* tile = reinterpret_cast<__half*>(smem);                          // [NP/8][C][8]
  __half* wsm = reinterpret_cast<__half*>(smem + (size_t)NP * C * 2);      // [M/8][K/8][8][8]  (K-major core matrices)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NP * C * 2 + (size_t)M * K * 2);   // full[2], empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < K * NP; i += kThreads) {                           // stands in for the cp.async loads / resample passes
    const int k = i / NP, p = i % NP;
    tile[((size_t)(p >> 3) * C + k) * 8 + (p & 7)] = Xg[i];
  }
  for (int i = tid; i < M * K; i += kThreads) {
    const int m = i / K, k = i % K;
    wsm[(((m >> 3) * (K / 8) + (k >> 3)) * 8 + (m & 7)) * 8 + (k & 7)] = m < COUT ? Wg[m * K + k] : __float2half(0.f);
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bars + b)) : "memory");        // full: one commit
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;\n" ::"r"(smem_u32(bars + 2 + b)) : "memory");  // empty: 128 epilogue threads
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if ((tid & 31) == 0) {
      // D = F32, A = B = F16, A K-major (weights), B MN-major (pixels contiguous), N = 256, M = 128
      const uint32_t idesc = (1u << 4) | (0u << 15) | (1u << 16) | ((uint32_t)(NB >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
      for (int nb = 0; nb < NP / NB; ++nb) {
        const int b = nb & 1;
        if (nb >= 2) mbar_wait(bars + 2 + b, ((nb >> 1) - 1) & 1);         // the epilogue drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int ks = 0; ks < K / 16; ++ks) {
          // A: core matrices [m group][k group]: LBO (second k half) = 128 B, SBO (next 8 rows) = (K/8) * 128 B
          const uint64_t da = make_desc(smem_u32(wsm) + ks * 2 * 128, 128, (K / 8) * 128);
          // B: tile [pixel group][channel][8 px]: core matrix = 8 channels x 8 px = 128 B; next k group (LBO) = 128 B,
          //    next pixel group (SBO) = C * 16 B
          const uint64_t db = make_desc(smem_u32(tile) + (uint32_t)nb * (NB / 8) * C * 16 + ks * 2 * 128, 128, C * 16);
          const uint32_t accumulate = ks > 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(tmem + (uint32_t)b * NB), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bars + b)) : "memory");
      }
    }
  } else {
    const int c = tid;                                                     // TMEM lane = output channel
    const float bi = c < COUT ? bias[c] : 0.f, sm1 = c < COUT ? slope[c] - 1.f : 0.f;
    for (int nb = 0; nb < NP / NB; ++nb) {
      const int b = nb & 1;
      mbar_wait(bars + b, (nb >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      for (int ch = 0; ch < NB / 32; ++ch) {                               // 32 pixels = 4 pixel groups at a time
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * NB + ch * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
            "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        if (c < COUT) {
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {                                 // one 16-byte row of the tile per pixel group
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v0 = __uint_as_float(r[g8 * 8 + 2 * j]) + bi, v1 = __uint_as_float(r[g8 * 8 + 2 * j + 1]) + bi;
              const __half2 h = __floats2half2_rn(fmaf(fminf(v0, 0.f), sm1, v0), fmaf(fminf(v1, 0.f), sm1, v1));
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            const int pg = nb * (NB / 8) + ch * 4 + g8;
            *reinterpret_cast<uint4*>(tile + ((size_t)pg * C + c) * 8) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bars + 2 + b)) : "memory");
    }
  }
  __syncthreads();
  for (int i = tid; i < COUT * NP; i += kThreads) {                        // tile -> planar global, for the check only
    const int c = i / NP, p = i % NP;
    Yg[i] = tile[((size_t)(p >> 3) * C + c) * 8 + (p & 7)];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
  std::vector<__half> X(K * NP), W(COUT * K), Y(COUT * NP);
  std::vector<float> Xf(K * NP), Wf(COUT * K), bias(COUT), slope(COUT);
  srand(2);
  for (int i = 0; i < K * NP; ++i) { Xf[i] = (rand() % 33 - 16) / 16.f; X[i] = __float2half(Xf[i]); }
  for (int i = 0; i < COUT * K; ++i) { Wf[i] = (rand() % 17 - 8) / 16.f; W[i] = __float2half(Wf[i]); }
  for (int c = 0; c < COUT; ++c) { bias[c] = (c % 7 - 3) / 4.f; slope[c] = 0.1f + 0.01f * c; }
  __half *dX, *dW, *dY; float *dB, *dS;
  CK(cudaMalloc(&dX, X.size() * 2)); CK(cudaMalloc(&dW, W.size() * 2)); CK(cudaMalloc(&dY, Y.size() * 2));
  CK(cudaMalloc(&dB, COUT * 4)); CK(cudaMalloc(&dS, COUT * 4));
  CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, bias.data(), COUT * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dS, slope.data(), COUT * 4, cudaMemcpyHostToDevice));
  const size_t smem_bytes = (size_t)NP * C * 2 + (size_t)M * K * 2 + 64;
  CK(cudaFuncSetAttribute(il_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  il_gemm<<<1, kThreads, smem_bytes>>>(dX, dW, dB, dS, dY);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 2; }
  CK(cudaMemcpy(Y.data(), dY, Y.size() * 2, cudaMemcpyDeviceToHost));
  double worst = 0; int bad = 0;
  for (int c = 0; c < COUT; ++c)
    for (int p = 0; p < NP; ++p) {
      float acc = bias[c];
      for (int k = 0; k < K; ++k) acc += Wf[c * K + k] * Xf[k * NP + p];
      const float ref = acc > 0.f ? acc : slope[c] * acc;
      const double d = fabs((double)__half2float(Y[c * NP + p]) - ref);
      if (!(d <= 2e-2 * (1.0 + fabs(ref)))) ++bad;
      if (d > worst || d != d) worst = d;
    }
  printf("tcgen05 in-place ILBlock GEMM prototype: max |Y - ref| = %g, %d of %d elements off\n", worst, bad, COUT * NP);
  return bad ? 3 : 0;
}

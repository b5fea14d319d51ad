// Issue-rate microbenchmark of the FMA flavours the depthwise tail could use (B200, sm_100a):
//   FFMA (fp32), FHFMA (fma.rn.f32.f16: fp16 x fp16 + fp32), HFMA2 (fp16x2), FFMA2 (fma.rn.f32x2), and FFMA / FHFMA mixed
//   with ALU-pipe work (FMNMX) to see which pipes co-issue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fma_rate scripts/fma_rate.cu && ./scripts/fma_rate
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 2048, CH = 8;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, float a0, uint32_t h0) {
  float acc[CH];
  uint32_t hacc[CH];
  float2 acc2[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { acc[i] = a0 + i + threadIdx.x; hacc[i] = h0 + i; acc2[i] = make_float2(a0 + i, a0 - i); }
  uint16_t x = (uint16_t)(h0 & 0xffff), w = (uint16_t)(h0 >> 16);
  float xf = a0 * 0.5f, wf = a0 * 0.25f;
  float2 x2 = make_float2(xf, wf), w2 = make_float2(wf, xf);
  uint16_t xs[CH], ws[CH], xs2[CH];
  float xfs[CH], wfs[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const uint32_t v = h0 + 0x00010001u * (uint32_t)(i + (threadIdx.x & 1)), u = (h0 >> 3) + 0x00030001u * (uint32_t)(i + (threadIdx.x & 3));
    xs[i] = (uint16_t)v; xs2[i] = (uint16_t)(v >> 16); ws[i] = (uint16_t)u;
    xfs[i] = a0 * (i + 1); wfs[i] = a0 / (i + 2 + (threadIdx.x & 1));
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (MODE == 0) acc[i] = fmaf(acc[i], xf, wf);                                                             // FFMA (2 distinct + acc)
      if (MODE == 1) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(x), "h"(w));          // FHFMA
      if (MODE == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(hacc[i]) : "r"(h0), "r"(h0 ^ 0x11u));  // HFMA2
      if (MODE == 3) {                                                                                           // FFMA2
        unsigned long long a, b, c;
        a = *reinterpret_cast<unsigned long long*>(&acc2[i]); b = *reinterpret_cast<unsigned long long*>(&x2); c = *reinterpret_cast<unsigned long long*>(&w2);
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a) : "l"(b), "l"(c));
        *reinterpret_cast<unsigned long long*>(&acc2[i]) = a;
      }
      if (MODE == 4) {                                                                                           // FHFMA + FMNMX 1:1
        asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(x), "h"(w));
        acc2[i].x = fminf(acc2[i].x, acc[(i + 1) % CH]);
      }
      if (MODE == 5) {                                                                                           // FFMA + FMNMX 1:1
        acc[i] = fmaf(acc[i], xf, wf);
        acc2[i].x = fminf(acc2[i].x, acc[(i + 1) % CH]);
      }
      if (MODE == 6) {                                                                                           // FHFMA + FFMA 1:1
        asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(x), "h"(w));
        acc2[i].x = fmaf(acc2[i].x, xf, wf);
      }
      if (MODE == 8) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(xs[i]), "h"(ws[i]));        // FHFMA, 3 distinct regs / instr
      if (MODE == 9) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(xs[i]), "h"(w));            // FHFMA, shared weight operand
      if (MODE == 10) acc[i] = fmaf(xfs[i], wfs[i], acc[i]);                                                          // FFMA, 3 distinct regs / instr
      if (MODE == 11) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(xs[i]), "h"(xs2[i]));      // FHFMA, .H1 of the same regs as 8's pairs
      if (MODE == 7) {                                                                                           // FHFMA + HFMA2 1:1
        asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(x), "h"(w));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(hacc[i]) : "r"(h0), "r"(h0 ^ 0x11u));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i] + __uint_as_float(hacc[i]) + acc2[i].x + acc2[i].y;
  if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_iter) {
  float* d; cudaMalloc(&d, 4096);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, 1024>>>(d, 1.0f, 0x3c003800u);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(d, 1.0f, 0x3c003800u);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double warp_instr = (double)ITERS * CH * per_iter * 32.0;        // per SM (32 warps)
  const double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-28s %.3f ms  -> %.2f warp-instr / clk / SM (%d instr per step)  [%s]\n", name, ms, warp_instr / cycles, per_iter, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0>("FFMA", 1);
  run<1>("FHFMA (f32 += f16*f16)", 1);
  run<2>("HFMA2", 1);
  run<3>("FFMA2 (f32x2)", 1);
  run<4>("FHFMA + FMNMX", 2);
  run<5>("FFMA + FMNMX", 2);
  run<6>("FHFMA + FFMA", 2);
  run<7>("FHFMA + HFMA2", 2);
  run<8>("FHFMA 3 distinct regs", 1);
  run<9>("FHFMA shared weight reg", 1);
  run<10>("FFMA 3 distinct regs", 1);
  run<11>("FHFMA H0/H1 of one reg pair", 1);
  return 0;
}

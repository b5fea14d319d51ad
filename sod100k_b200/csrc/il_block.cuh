// il_block.cuh — one kernel per ILBlock (1x1 kind): gOctaveCBR(1x1, 2 in-branches -> 1|2 out-branches)
// + depthwise 3x3 BN PReLU + depthwise 3x3 BN PReLU, everything between the block's input and output
// tensors stays in shared memory (reference: ILBlock.forward, CSNet/model/csnet.py:72-76, calling
// gOctaveConv.forward :664-726 and SimplifiedGOctConvBR.forward :838-851).
//
// One CTA owns a TH x TW tile of the high-resolution output branch and the co-located (TH/2 x TW/2) tile of
// the low-resolution branch, 16-bit activations (fp16 or bf16), fp32 accumulation:
//
//   load     XH[Chi][hi region, halo 4]  XL[.. Cli][lo region]           global -> smem, zero outside the image
//   pool     XL[0..Chi) = maxpool2x2(XH)                                   (hi -> lo path reads pooled input)
//   lo GEMM  [T1L ; U] = WL . XL     tensor cores (mma.sync m16n8k8), pixels are the N dimension
//            T1L = PReLU(. + b) (lo branch after conv+BN+PReLU),  U = W_lh . x_l (to be upsampled)
//   hi GEMM  T1H = PReLU(WH . XH + bilinear_x2(U) + b)
//   dw1      T2  = PReLU(dw3x3(T1) + b)      CUDA cores, fp32 accumulate, both branches
//   dw2      out = PReLU(dw3x3(T2) + b)  ->  global (tile interior only)
//
// T1/T2 are forced to 0 outside the image so the depthwise convs see the reference's zero padding.
// Region geometry (R = region-local coordinates):
//   hi region origin (hy0-4, hx0-4), size (TH+8) x (TW+8)          [halo 4 = 2 (two dw layers) x 2 (pooling)]
//   lo region origin (ly0-2, lx0-4), size (TH/2+4) x (TW/2+8)      [x origin kept a multiple of 4 for 8-byte I/O]
//   lo R(ry, rx)  <->  hi R(2ry, 2rx-4)
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace csnet {

constexpr int kIlThreads = 512;
constexpr int kIlRowsPerTask = 8;

struct DwParams {
  const float* w;      // [C][9]
  const float* b;      // [C]
  const float* s;      // [C] PReLU slope
};

struct IlArgs {
  const void* xh;
  const void* xl;
  void* yh;
  void* yl;                    // nullptr when Clo == 0
  const uint32_t* wh;          // packed 16-bit [MH16][KH8]
  const uint32_t* wl;          // packed 16-bit [ML16][KL8]
  const float *bias_h, *slope_h, *bias_l, *slope_l;
  DwParams dw1h, dw1l, dw2h, dw2l;
  int32_t H, W;                // hi resolution (lo = H/2 x W/2)
  int32_t Chi, Cli, Cho, Clo;
  int32_t TH, TW, tiles_x;
  int32_t KH8, KL8, MH16, ML16, pool_rows;
  int32_t NPH, NPL;            // padded flat region sizes (multiples of 8, NP/8 odd)
  int32_t rowsAh, rowsAl;      // rows of the X/T2 buffers
};

// ---- 16-bit helpers -----------------------------------------------------------------------------------
template <typename T> struct Pack;
template <> struct Pack<__half> {
  using T2 = __half2;
  static __device__ __forceinline__ float2 to_f2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b));
  }
};
template <> struct Pack<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static __device__ __forceinline__ float2 to_f2(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v)); }
  static __device__ __forceinline__ uint32_t from_f2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  static __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  static __device__ __forceinline__ void mma(float* c, const uint32_t* a, uint32_t b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(b));
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t* r, const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)));
}

__device__ __forceinline__ float prelu(float v, float s) { return v > 0.f ? v : s * v; }

// D[M16 x NP] = Ws[M16 x K8] . X[K8 x NP]; epi(m, p, v0, v1) receives rows m and two adjacent pixels p, p+1.
template <typename T, typename Epi>
__device__ __forceinline__ void gemm_pixels(const uint16_t* Ws, int M16, int K8, const uint16_t* X, int NP, int warp,
                                            int nwarps, int lane, Epi epi) {
  const int ntiles = NP >> 3;
  const int g = lane >> 2, t = lane & 3;
  for (int nt0 = warp * 4; nt0 < ntiles; nt0 += nwarps * 4) {
    int ntl = nt0 + (lane >> 3);
    ntl = ntl < ntiles ? ntl : ntiles - 1;            // clamp: result of a clamped tile is discarded
    for (int mt0 = 0; mt0 < (M16 >> 4); mt0 += 2) {
      const int mts = ((M16 >> 4) - mt0) < 2 ? ((M16 >> 4) - mt0) : 2;
      float acc[2][4][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
      for (int ks = 0; ks < (K8 >> 3); ++ks) {
        uint32_t bf[4];
        ldmatrix_x4_trans(bf, X + (size_t)(ks * 8 + (lane & 7)) * NP + ntl * 8);
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
          if (mi < mts) {
            uint32_t af[2];
            ldmatrix_x2(af, Ws + (size_t)((mt0 + mi) * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * K8 + ks * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) Pack<T>::mma(acc[mi][j], af, bf[j]);
          }
        }
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        if (mi < mts) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int nt = nt0 + j;
            if (nt < ntiles) {
              const int p = nt * 8 + 2 * t;
              epi((mt0 + mi) * 16 + g, p, acc[mi][j][0], acc[mi][j][1]);
              epi((mt0 + mi) * 16 + g + 8, p, acc[mi][j][2], acc[mi][j][3]);
            }
          }
        }
      }
    }
  }
}

// Depthwise 3x3 + bias + PReLU over region rows [r0, r1), 4-pixel groups [g0, g1) of every channel.
// in/out: [C][NP] flat region planes with row stride RW.  Output pixels outside the image are written as 0
// (smem destination) or skipped (global destination).
template <typename T, bool kToGlobal>
__device__ __forceinline__ void dw_pass(const uint16_t* in, uint16_t* out_s, uint16_t* out_g, int C, int RW, int NP,
                                        int r0, int r1, int g0, int g1, DwParams P, int oy0, int ox0, int imgH,
                                        int imgW, int tid, int nthreads) {
  const int G = g1 - g0;
  const int nruns = (r1 - r0 + kIlRowsPerTask - 1) / kIlRowsPerTask;
  const int ntasks = C * nruns * G;
  for (int task = tid; task < ntasks; task += nthreads) {
    const int gi = task % G, run = (task / G) % nruns, c = task / (G * nruns);
    const int x = 4 * (g0 + gi);
    const int ra = r0 + run * kIlRowsPerTask;
    const int rb = (ra + kIlRowsPerTask) < r1 ? (ra + kIlRowsPerTask) : r1;
    float w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = __ldg(P.w + c * 9 + i);
    const float bias = __ldg(P.b + c), slope = __ldg(P.s + c);
    const uint16_t* plane = in + (size_t)c * NP;
    float rows[3][6];
    auto load_row = [&](int r, float* dst) {
      const uint32_t* q = reinterpret_cast<const uint32_t*>(plane + r * RW + x);   // 4-byte aligned (x, RW, NP even)
      const float2 a = Pack<T>::to_f2(q[-1]), b = Pack<T>::to_f2(q[0]), c2 = Pack<T>::to_f2(q[1]), d = Pack<T>::to_f2(q[2]);
      dst[0] = a.y; dst[1] = b.x; dst[2] = b.y; dst[3] = c2.x; dst[4] = c2.y; dst[5] = d.x;
    };
    load_row(ra - 1, rows[0]);
    load_row(ra, rows[1]);
    for (int r = ra; r < rb; ++r) {
      load_row(r + 1, rows[2]);
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = bias;
        v = fmaf(rows[0][i], w[0], v); v = fmaf(rows[0][i + 1], w[1], v); v = fmaf(rows[0][i + 2], w[2], v);
        v = fmaf(rows[1][i], w[3], v); v = fmaf(rows[1][i + 1], w[4], v); v = fmaf(rows[1][i + 2], w[5], v);
        v = fmaf(rows[2][i], w[6], v); v = fmaf(rows[2][i + 1], w[7], v); v = fmaf(rows[2][i + 2], w[8], v);
        o[i] = prelu(v, slope);
      }
      const int gy = oy0 + r, gx = ox0 + x;
      const bool row_in = gy >= 0 && gy < imgH;
      if (kToGlobal) {
        if (row_in && gx >= 0 && gx < imgW) {          // imgW % 4 == 0 and gx % 4 == 0: the group is all in or all out
          uint2 v;
          v.x = Pack<T>::from_f2(o[0], o[1]);
          v.y = Pack<T>::from_f2(o[2], o[3]);
          *reinterpret_cast<uint2*>(out_g + ((size_t)c * imgH + gy) * imgW + gx) = v;
        }
      } else {
        const bool in = row_in && gx >= 0 && gx < imgW;
        uint2 v;
        v.x = in ? Pack<T>::from_f2(o[0], o[1]) : 0u;
        v.y = in ? Pack<T>::from_f2(o[2], o[3]) : 0u;
        *reinterpret_cast<uint2*>(out_s + (size_t)c * NP + r * RW + x) = v;
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) { rows[0][i] = rows[1][i]; rows[1][i] = rows[2][i]; }
    }
  }
}

inline size_t il_smem_bytes(const IlArgs& A) {
  size_t halves = 8 + (size_t)A.rowsAh * A.NPH + (size_t)A.Cho * A.NPH + (size_t)A.rowsAl * A.NPL +
                  (size_t)A.Clo * A.NPL + (size_t)A.Cho * A.NPL + (size_t)A.MH16 * A.KH8 + (size_t)A.ML16 * A.KL8 + 8;
  return halves * 2;
}

template <typename T>
__global__ void __launch_bounds__(kIlThreads, 1) il_block_kernel(const __grid_constant__ IlArgs A) {
  extern __shared__ __align__(16) uint16_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kIlThreads >> 5;
  const int n = blockIdx.z;
  const int tile_y = blockIdx.x / A.tiles_x, tile_x = blockIdx.x % A.tiles_x;
  const int hy0 = tile_y * A.TH, hx0 = tile_x * A.TW, ly0 = hy0 >> 1, lx0 = hx0 >> 1;
  const int H = A.H, W = A.W, Hl = A.H >> 1, Wl = A.W >> 1;
  const int RHh = A.TH + 8, RWh = A.TW + 8, RHl = (A.TH >> 1) + 4, RWl = (A.TW >> 1) + 8;
  const int NPH = A.NPH, NPL = A.NPL;

  uint16_t* bufAh = smem + 8;                          // XH, later T2H
  uint16_t* bufBh = bufAh + (size_t)A.rowsAh * NPH;    // T1H
  uint16_t* bufAl = bufBh + (size_t)A.Cho * NPH;       // XL, later T2L
  uint16_t* bufBl = bufAl + (size_t)A.rowsAl * NPL;    // T1L
  uint16_t* bufU = bufBl + (size_t)A.Clo * NPL;        // U = W_lh . x_l
  uint16_t* wsH = bufU + (size_t)A.Cho * NPL;
  uint16_t* wsL = wsH + A.MH16 * A.KH8;
  uint16_t* guard1 = wsL + A.ML16 * A.KL8;

  // ---- phase 0/1: weights + input regions -> smem ------------------------------------------------------
  if (tid < 4) {
    reinterpret_cast<uint32_t*>(smem)[tid] = 0u;
    reinterpret_cast<uint32_t*>(guard1)[tid] = 0u;
  }
  for (int i = tid; i < (A.MH16 * A.KH8) >> 1; i += kIlThreads) reinterpret_cast<uint32_t*>(wsH)[i] = __ldg(A.wh + i);
  for (int i = tid; i < (A.ML16 * A.KL8) >> 1; i += kIlThreads) reinterpret_cast<uint32_t*>(wsL)[i] = __ldg(A.wl + i);
  {
    const uint16_t* xh = reinterpret_cast<const uint16_t*>(A.xh) + (size_t)n * A.Chi * H * W;
    const int halfW = NPH >> 1, pairs_row = RWh >> 1;
    for (int i = tid; i < A.rowsAh * halfW; i += kIlThreads) {
      const int c = i / halfW, pp = i % halfW;
      const int ry = pp / pairs_row, rx = (pp % pairs_row) * 2;
      const int gy = hy0 - 4 + ry, gx = hx0 - 4 + rx;
      uint32_t v = 0u;
      if (c < A.Chi && ry < RHh && gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const uint32_t*>(xh + ((size_t)c * H + gy) * W + gx));
      reinterpret_cast<uint32_t*>(bufAh + (size_t)c * NPH)[pp] = v;
    }
    const uint16_t* xl = reinterpret_cast<const uint16_t*>(A.xl) + (size_t)n * A.Cli * Hl * Wl;
    const int halfWl = NPL >> 1, pairs_rowl = RWl >> 1;
    const int rows_lo = A.rowsAl - A.pool_rows;
    for (int i = tid; i < rows_lo * halfWl; i += kIlThreads) {
      const int c = i / halfWl, pp = i % halfWl;
      const int ry = pp / pairs_rowl, rx = (pp % pairs_rowl) * 2;
      const int gy = ly0 - 2 + ry, gx = lx0 - 4 + rx;
      uint32_t v = 0u;
      if (c < A.Cli && ry < RHl && gy >= 0 && gy < Hl && gx >= 0 && gx < Wl)
        v = __ldg(reinterpret_cast<const uint32_t*>(xl + ((size_t)c * Hl + gy) * Wl + gx));
      reinterpret_cast<uint32_t*>(bufAl + (size_t)(A.pool_rows + c) * NPL)[pp] = v;
    }
  }
  __syncthreads();

  // ---- phase 2: pooled rows of XL (max_pool2d 2x2 of the hi input, csnet.py:709-712) -------------------
  if (A.pool_rows > 0) {
    const int halfWl = NPL >> 1, pairs_rowl = RWl >> 1;
    for (int i = tid; i < A.pool_rows * halfWl; i += kIlThreads) {
      const int c = i / halfWl, pp = i % halfWl;
      const int ry = pp / pairs_rowl, rx = (pp % pairs_rowl) * 2;          // lo region coords of the pair (rx, rx+1)
      uint32_t v = 0u;
      const int hx = 2 * rx - 4;                                           // hi region col of lo col rx
      if (ry < RHl && hx >= 0 && hx + 3 < RWh) {
        const uint16_t* r0 = bufAh + (size_t)c * NPH + (2 * ry) * RWh + hx;
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(r0), a1 = *reinterpret_cast<const uint32_t*>(r0 + 2);
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(r0 + RWh), b1 = *reinterpret_cast<const uint32_t*>(r0 + RWh + 2);
        const uint32_t m0 = Pack<T>::max2(a0, b0), m1 = Pack<T>::max2(a1, b1);   // vertical max of 2 hi pixel pairs
        const float2 f0 = Pack<T>::to_f2(m0), f1 = Pack<T>::to_f2(m1);
        v = Pack<T>::from_f2(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y));
      }
      reinterpret_cast<uint32_t*>(bufAl + (size_t)c * NPL)[pp] = v;
    }
    __syncthreads();
  }

  // ---- phase 3: lo GEMM -> T1L (rows < Clo) and U (rows Clo .. Clo+Cho) --------------------------------
  {
    const int Clo = A.Clo, Cho = A.Cho;
    const float* bl = A.bias_l;
    const float* sl = A.slope_l;
    auto epi = [&](int m, int p, float v0, float v1) {
      if (m < Clo) {
        const int ry = p / RWl, rx = p % RWl;
        const int gy = ly0 - 2 + ry, gx = lx0 - 4 + rx;
        const bool rin = ry < RHl && gy >= 0 && gy < Hl;
        const float b = __ldg(bl + m), s = __ldg(sl + m);
        const float o0 = (rin && gx >= 0 && gx < Wl) ? prelu(v0 + b, s) : 0.f;
        const float o1 = (rin && gx + 1 >= 0 && gx + 1 < Wl) ? prelu(v1 + b, s) : 0.f;
        *reinterpret_cast<uint32_t*>(bufBl + (size_t)m * NPL + p) = Pack<T>::from_f2(o0, o1);
      } else if (m - Clo < Cho) {
        *reinterpret_cast<uint32_t*>(bufU + (size_t)(m - Clo) * NPL + p) = Pack<T>::from_f2(v0, v1);
      }
    };
    gemm_pixels<T>(wsL, A.ML16, A.KL8, bufAl, NPL, warp, nwarps, lane, epi);
  }
  __syncthreads();

  // ---- phase 4: hi GEMM + bilinear x2 of U + bias + PReLU -> T1H ---------------------------------------
  {
    const int Cho = A.Cho;
    const float* bh = A.bias_h;
    const float* sh = A.slope_h;
    auto epi = [&](int m, int p, float v0, float v1) {
      if (m >= Cho) return;
      const int ry = p / RWh, rx = p % RWh;                 // rx even, pixels rx and rx+1 share the row
      const int gy = hy0 - 4 + ry, gx = hx0 - 4 + rx;
      float o0 = 0.f, o1 = 0.f;
      if (ry >= 2 && ry < RHh - 2 && rx >= 2 && rx < RWh - 2 && gy >= 0 && gy < H && gx >= 0 && gx < W) {
        // F.interpolate(scale_factor=2, bilinear, align_corners=False): src = (dst + .5)/2 - .5 clamped at 0
        float sy = ((float)gy + 0.5f) * 0.5f - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        const int y0 = (int)sy, y1 = y0 + (y0 < Hl - 1 ? 1 : 0);
        const float wy1 = sy - (float)y0, wy0 = 1.f - wy1;
        const uint16_t* u0 = bufU + (size_t)m * NPL + (y0 - (ly0 - 2)) * RWl - (lx0 - 4);
        const uint16_t* u1 = bufU + (size_t)m * NPL + (y1 - (ly0 - 2)) * RWl - (lx0 - 4);
        const float b = __ldg(bh + m), s = __ldg(sh + m);
        auto sample = [&](int x) {
          float sx = ((float)x + 0.5f) * 0.5f - 0.5f;
          sx = sx < 0.f ? 0.f : sx;
          const int x0 = (int)sx, x1 = x0 + (x0 < Wl - 1 ? 1 : 0);
          const float wx1 = sx - (float)x0, wx0 = 1.f - wx1;
          const float v00 = Pack<T>::to_f2((uint32_t)u0[x0]).x, v01 = Pack<T>::to_f2((uint32_t)u0[x1]).x;
          const float v10 = Pack<T>::to_f2((uint32_t)u1[x0]).x, v11 = Pack<T>::to_f2((uint32_t)u1[x1]).x;
          return wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11);
        };
        o0 = prelu(v0 + sample(gx) + b, s);
        o1 = prelu(v1 + sample(gx + 1) + b, s);   // W even and gx even: gx+1 is inside the image too
      }
      *reinterpret_cast<uint32_t*>(bufBh + (size_t)m * NPH + p) = Pack<T>::from_f2(o0, o1);
    };
    gemm_pixels<T>(wsH, A.MH16, A.KH8, bufAh, NPH, warp, nwarps, lane, epi);
  }
  __syncthreads();

  // ---- phase 5: dw1 (T1 -> T2, smem) --------------------------------------------------------------------
  dw_pass<T, false>(bufBh, bufAh, nullptr, A.Cho, RWh, NPH, 3, RHh - 3, 0, RWh >> 2, A.dw1h, hy0 - 4, hx0 - 4, H, W, tid, kIlThreads);
  if (A.Clo > 0)
    dw_pass<T, false>(bufBl, bufAl, nullptr, A.Clo, RWl, NPL, 1, RHl - 1, 0, RWl >> 2, A.dw1l, ly0 - 2, lx0 - 4, Hl, Wl, tid, kIlThreads);
  __syncthreads();

  // ---- phase 6: dw2 (T2 -> global) ----------------------------------------------------------------------
  dw_pass<T, true>(bufAh, nullptr, reinterpret_cast<uint16_t*>(A.yh) + (size_t)n * A.Cho * H * W, A.Cho, RWh, NPH, 4, RHh - 4,
                   1, (RWh >> 2) - 1, A.dw2h, hy0 - 4, hx0 - 4, H, W, tid, kIlThreads);
  if (A.Clo > 0)
    dw_pass<T, true>(bufAl, nullptr, reinterpret_cast<uint16_t*>(A.yl) + (size_t)n * A.Clo * Hl * Wl, A.Clo, RWl, NPL, 2,
                     RHl - 2, 1, (RWl >> 2) - 1, A.dw2l, ly0 - 2, lx0 - 4, Hl, Wl, tid, kIlThreads);
}

}  // namespace csnet

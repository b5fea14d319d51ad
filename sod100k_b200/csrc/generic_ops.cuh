// generic_ops.cuh — shape-generic fused ops (any channel count / resolution / dtype).
//
// These are the always-correct kernels of the engine: every op of the program IR (include/csnet_b200.h)
// can run through them.  The per-thread bodies are plain functions so the very same source is also
// compiled for the host by tests/emu (CSNET_HOST_EMU) and checked against the oracle without a GPU.
// Shape-specialised fast kernels (il_block.cuh, ...) take over the hot shapes; these remain the fallback.
//
// Semantics follow the reference call sites named in include/csnet_b200.h:
//   conv path   = [avg_pool2d 2x2] -> [max_pool2d k] -> conv2d(zero pad, dilation, stride)
//   resample    = bilinear, align_corners=False, src = (dst+0.5)/up-0.5 clamped at 0
//   epilogue    = + bias[c], PReLU slope[c]
#pragma once
#include <stdint.h>

#ifdef CSNET_HOST_EMU
#include <math.h>
#define CSNET_DEV inline
#define CSNET_HD inline
#else
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#define CSNET_DEV __device__ __forceinline__
#define CSNET_HD __host__ __device__ __forceinline__
#endif

namespace csnet {

enum { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };
constexpr int kMaxPaths = 8;
constexpr int kMixCT = 16;    // output channels per thread in the generic MIX kernel
constexpr int kDwRows = 4;    // output rows per thread in the generic DW kernel

struct MixPath {
  const void* src;
  const float* w;             // [cin][ksize*ksize][cout] (cout innermost), BN scale folded in
  int32_t dtype, C, H, W;     // source tensor (full) dims
  int32_t c0, cin;
  int32_t pre_avg, pool;
  int32_t ksize, dil, stride, pad;
  int32_t up;
  int32_t cout0, cout;
};

struct MixArgs {
  void* dst;
  const float* bias;          // nullptr: none
  const float* slope;         // nullptr: none
  int32_t dtype, C, H, W;
  int32_t n_paths;
  const float* proj_w;        // CSNET_OP_MIXPROJ: [C] projection weights (the C-channel result is not stored); else nullptr
  const float* proj_b;        // projection bias (one float) or nullptr
  MixPath p[kMaxPaths];
};

struct DwArgs {
  const void* src;
  void* dst;
  const float* w;             // [C][9], BN scale and the x100 folded in
  const float* bias;
  const float* slope;
  int32_t src_dtype, dst_dtype, C, H, W;
};

CSNET_DEV float ld_elem(const void* p, int dtype, int64_t i) {
#ifndef CSNET_HOST_EMU
  if (dtype == DT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  if (dtype == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
#endif
  return reinterpret_cast<const float*>(p)[i];
}

CSNET_DEV void st_elem(void* p, int dtype, int64_t i, float v) {
#ifndef CSNET_HOST_EMU
  if (dtype == DT_F16) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); return; }
  if (dtype == DT_BF16) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); return; }
#endif
  reinterpret_cast<float*>(p)[i] = v;
}

// pre_avg = f in {0, 1 (legacy: 2), 2, 4, 8}: down-sample by f first.  f = 2 is avg_pool2d(2,2) (csnet.py:679-680) and,
// identically, F.interpolate(bilinear) to half size; f = 4 / 8 is F.interpolate(bilinear, align_corners=False) to a
// quarter / eighth (CSF+Res2Net/networks/gOctConv.py:101-102): source index f*d + f/2 - 0.5, i.e. the mean of the
// 2x2 pixels at offset f/2 - 1 of each f x f cell.
CSNET_HD int pre_factor(int pre_avg) { return pre_avg == 0 ? 1 : (pre_avg == 1 ? 2 : pre_avg); }

// Value of the (down-sampled, max-pooled) source plane at pooled-grid position (y, x); y, x in range.
CSNET_DEV float fetch_pooled(const MixPath& P, int64_t plane, int y, int x) {
  if (!P.pre_avg && P.pool == 1) return ld_elem(P.src, P.dtype, plane + (int64_t)y * P.W + x);
  const int f = pre_factor(P.pre_avg), fo = (f >> 1) - 1;
  float m = -INFINITY;
  for (int py = 0; py < P.pool; ++py) {
    for (int px = 0; px < P.pool; ++px) {
      const int yy = y * P.pool + py, xx = x * P.pool + px;
      float v;
      if (P.pre_avg) {
        const int64_t b = plane + (int64_t)(f * yy + fo) * P.W + f * xx + fo;
        v = ((ld_elem(P.src, P.dtype, b) + ld_elem(P.src, P.dtype, b + 1)) + ld_elem(P.src, P.dtype, b + P.W)) +
            ld_elem(P.src, P.dtype, b + P.W + 1);
        v *= 0.25f;
      } else {
        v = ld_elem(P.src, P.dtype, plane + (int64_t)yy * P.W + xx);
      }
      m = v > m ? v : m;
    }
  }
  return m;
}

// Bilinear sample (align_corners=False) of plane [Hs,Ws] at destination pixel (oy,ox), factor `up`.
CSNET_DEV float bilinear_up(const void* src, int dtype, int64_t plane, int Hs, int Ws, int up, int oy, int ox) {
  const float inv = 1.0f / (float)up;
  float sy = ((float)oy + 0.5f) * inv - 0.5f;
  float sx = ((float)ox + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v00 = ld_elem(src, dtype, plane + (int64_t)y0 * Ws + x0);
  const float v01 = ld_elem(src, dtype, plane + (int64_t)y0 * Ws + x1);
  const float v10 = ld_elem(src, dtype, plane + (int64_t)y1 * Ws + x0);
  const float v11 = ld_elem(src, dtype, plane + (int64_t)y1 * Ws + x1);
  return hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// The generic MIX kernel walks K in chunks: for every conv path that touches this cout tile, for every chunk of
// input channels, the CTA stages that chunk's weights in shared memory ([ci][tap][kMixCT], zero where the output channel
// is outside the path's slice, so the accumulate loop needs no predicates) and every thread accumulates its pixel.
constexpr int kMixStageFloats = 8192;                       // 32 KB of staged weights per chunk

CSNET_HD int mix_chunk_channels(int ksize) {
  const int c = kMixStageFloats / (ksize * ksize * kMixCT);
  return c < 1 ? 1 : c;
}

CSNET_DEV bool mix_path_live(const MixPath& P, int co_base) {
  const int lo = co_base > P.cout0 ? co_base : P.cout0;
  const int hi = (co_base + kMixCT) < (P.cout0 + P.cout) ? (co_base + kMixCT) : (P.cout0 + P.cout);
  return lo < hi;
}

// Cooperative: weights of input channels [ci0, ci1) of path P for this cout tile -> ws[(ci-ci0)][tap][kMixCT].
CSNET_DEV void mix_stage_chunk(const MixPath& P, int co_base, int ci0, int ci1, float* ws, int tid, int nthreads) {
  const int kk = P.ksize * P.ksize, rows = (ci1 - ci0) * kk;
  for (int i = tid; i < rows * kMixCT; i += nthreads) {
    const int r = i / kMixCT, t = i % kMixCT;
    const int co = co_base + t;
    ws[i] = (co >= P.cout0 && co < P.cout0 + P.cout) ? P.w[((int64_t)ci0 * kk + r) * P.cout + (co - P.cout0)] : 0.f;
  }
}

// One output pixel: accumulate input channels [ci0, ci1) of conv path P.
CSNET_DEV void mix_acc_chunk(const MixPath& P, const float* ws, int ci0, int ci1, int n, int oy, int ox, float* acc) {
  const int64_t plane_sz = (int64_t)P.H * P.W;
  const int div = pre_factor(P.pre_avg) * P.pool;
  const int Hc = P.up > 1 ? P.H * P.up : P.H / div, Wc = P.up > 1 ? P.W * P.up : P.W / div;
  const int kk = P.ksize * P.ksize;
  for (int ci = ci0; ci < ci1; ++ci) {
    const int64_t plane = ((int64_t)n * P.C + P.c0 + ci) * plane_sz;
    for (int ky = 0; ky < P.ksize; ++ky) {
      const int y = oy * P.stride - P.pad + ky * P.dil;
      for (int kx = 0; kx < P.ksize; ++kx) {
        const int x = ox * P.stride - P.pad + kx * P.dil;
        const float* wr = ws + ((ci - ci0) * kk + ky * P.ksize + kx) * kMixCT;
        if (y < 0 || y >= Hc || x < 0 || x >= Wc) continue;   // zero padding
        // up > 1 on a (1x1) conv path: the source is bilinearly up-sampled BEFORE the conv (same linear map as
        // the reference's conv-then-interpolate, csnet.py:702-707)
        const float v = P.up > 1 ? bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, y, x) : fetch_pooled(P, plane, y, x);
#pragma unroll
        for (int t = 0; t < kMixCT; ++t) acc[t] += v * wr[t];
      }
    }
  }
}

// One output pixel: resample-add paths, bias, PReLU, store.
CSNET_DEV void mix_finish(const MixArgs& A, int n, int oy, int ox, int co_base, float* acc) {
  for (int p = 0; p < A.n_paths; ++p) {
    const MixPath& P = A.p[p];
    if (P.ksize != 0 || !mix_path_live(P, co_base)) continue;
    const int64_t plane_sz = (int64_t)P.H * P.W;
#pragma unroll
    for (int t = 0; t < kMixCT; ++t) {
      const int co = co_base + t;
      if (co >= P.cout0 && co < P.cout0 + P.cout) {
        const int64_t plane = ((int64_t)n * P.C + P.c0 + (co - P.cout0)) * plane_sz;
        // a resample path adds the source up-sampled (up > 1), or down-sampled by pre_avg / pool (up == 1), or as it is
        acc[t] += (P.pre_avg || P.pool > 1) ? fetch_pooled(P, plane, oy, ox) : bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox);
      }
    }
  }
  const int64_t out_plane = (int64_t)A.H * A.W;
#pragma unroll
  for (int t = 0; t < kMixCT; ++t) {
    const int co = co_base + t;
    if (co < A.C) {
      float y = acc[t];
      if (A.bias) y += A.bias[co];
      if (A.slope) y = y > 0.f ? y : A.slope[co] * y;
      st_elem(A.dst, A.dtype, ((int64_t)n * A.C + co) * out_plane + (int64_t)oy * A.W + ox, y);
    }
  }
}

// GroupNorm(groups) + PReLU (CSF+Res2Net/networks/gOctConv.py:133, csf_res2net.py:220-224): statistics per (image, group)
// over (C/groups, H, W), eps 1e-5, biased variance; then y = prelu(gamma[c] * (x - mean) * rstd + beta[c]).
struct GnArgs {
  const void* src;
  void* dst;
  const float* gamma;
  const float* beta;
  const float* slope;        // nullptr: no PReLU
  float* stats;              // [N][groups][2] = (mean, rstd), written by the stats pass
  int32_t src_dtype, dst_dtype, C, HW, groups;
};

// Depthwise 3x3 (pad 1) + bias + PReLU: one column x, rows [oy0, oy0 + kDwRows) of plane (n, c).
CSNET_DEV void dw_thread(const DwArgs& A, int n, int c, int oy0, int ox) {
  const int64_t plane = ((int64_t)n * A.C + c) * (int64_t)A.H * A.W;
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = A.w[c * 9 + i];
  const float b = A.bias ? A.bias[c] : 0.f;
  const bool has_slope = A.slope != nullptr;
  const float s = has_slope ? A.slope[c] : 1.f;
  const bool xl = ox > 0, xr = ox < A.W - 1;
  float r0[3], r1[3], r2[3];   // rows y-1, y, y+1: (x-1, x, x+1)
  auto load_row = [&](int y, float* r) {
    if (y < 0 || y >= A.H) { r[0] = r[1] = r[2] = 0.f; return; }
    const int64_t base = plane + (int64_t)y * A.W + ox;
    r[0] = xl ? ld_elem(A.src, A.src_dtype, base - 1) : 0.f;
    r[1] = ld_elem(A.src, A.src_dtype, base);
    r[2] = xr ? ld_elem(A.src, A.src_dtype, base + 1) : 0.f;
  };
  load_row(oy0 - 1, r0);
  load_row(oy0, r1);
#pragma unroll
  for (int i = 0; i < kDwRows; ++i) {
    const int oy = oy0 + i;
    if (oy >= A.H) break;
    load_row(oy + 1, r2);
    float y = 0.f;
    y += r0[0] * w[0]; y += r0[1] * w[1]; y += r0[2] * w[2];
    y += r1[0] * w[3]; y += r1[1] * w[4]; y += r1[2] * w[5];
    y += r2[0] * w[6]; y += r2[1] * w[7]; y += r2[2] * w[8];
    y += b;
    if (has_slope) y = y > 0.f ? y : s * y;
    st_elem(A.dst, A.dst_dtype, plane + (int64_t)oy * A.W + ox, y);
#pragma unroll
    for (int k = 0; k < 3; ++k) { r0[k] = r1[k]; r1[k] = r2[k]; }
  }
}

}  // namespace csnet

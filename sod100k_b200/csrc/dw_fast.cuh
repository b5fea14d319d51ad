// dw_fast.cuh — depthwise 3x3 (pad 1) + bias + PReLU on 16-bit planar tensors, W % 4 == 0
// (SimplifiedGOctConvBR.forward, CSNet/model/csnet.py:838-851, for the blocks the fused ILBlock kernel does not cover).
// A thread produces a 4-pixel-wide, RUN-row strip of one plane: all (RUN+2) x 3 loads are issued before any math,
// 8-byte stores; fp32 accumulate.
#pragma once
#include "generic_ops.cuh"
#include "il_block.cuh"

namespace csnet {

constexpr int kDwfThreads = 256;
constexpr int kDwfRun = 4;

template <typename T>
__global__ void __launch_bounds__(kDwfThreads) dw_fast_kernel(const __grid_constant__ DwArgs A) {
  const int G = A.W >> 2, nruns = (A.H + kDwfRun - 1) / kDwfRun;
  const int task = blockIdx.x * kDwfThreads + threadIdx.x;
  if (task >= G * nruns) return;
  const int gi = task % G, run = task / G;
  const int x = 4 * gi, ra = run * kDwfRun, c = blockIdx.y, H = A.H, W = A.W;
  const size_t plane_off = ((size_t)blockIdx.z * A.C + c) * (size_t)H * W;
  const uint16_t* plane = reinterpret_cast<const uint16_t*>(A.src) + plane_off + x;
  uint2 mid[kDwfRun + 2];
  uint32_t lft[kDwfRun + 2], rgt[kDwfRun + 2];
#pragma unroll
  for (int i = 0; i < kDwfRun + 2; ++i) {
    const int r = ra - 1 + i;
    const bool ok = r >= 0 && r < H;
    mid[i] = ok ? *reinterpret_cast<const uint2*>(plane + (size_t)r * W) : make_uint2(0u, 0u);
    lft[i] = (ok && x > 0) ? *reinterpret_cast<const uint32_t*>(plane + (size_t)r * W - 2) : 0u;
    rgt[i] = (ok && x + 4 < W) ? *reinterpret_cast<const uint32_t*>(plane + (size_t)r * W + 4) : 0u;
  }
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = __ldg(A.w + c * 9 + i);
  const float bias = A.bias ? __ldg(A.bias + c) : 0.f;
  const bool has_slope = A.slope != nullptr;
  const float slope = has_slope ? __ldg(A.slope + c) : 1.f;
  float rows[kDwfRun + 2][6];
#pragma unroll
  for (int i = 0; i < kDwfRun + 2; ++i) {
    const float2 a = Pack<T>::to_f2(lft[i]), b = Pack<T>::to_f2(mid[i].x), c2 = Pack<T>::to_f2(mid[i].y), d = Pack<T>::to_f2(rgt[i]);
    rows[i][0] = a.y; rows[i][1] = b.x; rows[i][2] = b.y; rows[i][3] = c2.x; rows[i][4] = c2.y; rows[i][5] = d.x;
  }
  uint16_t* out = reinterpret_cast<uint16_t*>(A.dst) + plane_off + x;
#pragma unroll
  for (int i = 0; i < kDwfRun; ++i) {
    const int r = ra + i;
    if (r < H) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v = bias;
        v = fmaf(rows[i][k], w[0], v); v = fmaf(rows[i][k + 1], w[1], v); v = fmaf(rows[i][k + 2], w[2], v);
        v = fmaf(rows[i + 1][k], w[3], v); v = fmaf(rows[i + 1][k + 1], w[4], v); v = fmaf(rows[i + 1][k + 2], w[5], v);
        v = fmaf(rows[i + 2][k], w[6], v); v = fmaf(rows[i + 2][k + 1], w[7], v); v = fmaf(rows[i + 2][k + 2], w[8], v);
        o[k] = has_slope ? prelu(v, slope) : v;
      }
      uint2 v;
      v.x = Pack<T>::from_f2(o[0], o[1]);
      v.y = Pack<T>::from_f2(o[2], o[3]);
      *reinterpret_cast<uint2*>(out + (size_t)r * W) = v;
    }
  }
}

// A MIX op that is a single resample path (F.interpolate(bilinear, align_corners=False), the final x2 of CSNet.forward,
// csnet.py:385-386, or the x4 of CSF+Res2Net): one thread per output pixel of one plane, the same bilinear_up() arithmetic
// as the generic kernel without its path / channel-tile loops.
__global__ void __launch_bounds__(256) resample_fast_kernel(const __grid_constant__ MixArgs A) {
  const int64_t hw = (int64_t)A.H * A.W;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= hw) return;
  const MixPath& P = A.p[0];
  const int co = blockIdx.y, n = blockIdx.z;
  const int oy = (int)(i / A.W), ox = (int)(i - (int64_t)oy * A.W);
  const int64_t plane = ((int64_t)n * P.C + P.c0 + co) * (int64_t)P.H * P.W;
  float y = (P.pre_avg || P.pool > 1) ? fetch_pooled(P, plane, oy, ox) : bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox);
  if (A.bias) y += __ldg(A.bias + co);
  if (A.slope) y = prelu(y, __ldg(A.slope + co));
  st_elem(A.dst, A.dtype, ((int64_t)n * A.C + co) * hw + i, y);
}

// avg_pool2d(2, 2) (the stride-2 entry of gOctaveConv, csnet.py:679-680) or max_pool2d(2, 2) (its high-to-low paths,
// :708-717) of a 16-bit tensor, materialised once for all the conv paths that consume it: 4 output pixels per thread from
// two 16-byte loads; the average uses the same (((a + b) + c) + d) / 4 order as fetch_pooled(), so the stored value equals
// what the MIX kernels would stage (a maximum is exact anyway).  A.W % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256) pool2_fast_kernel(const __grid_constant__ MixArgs A, const bool is_max) {
  const int G = A.W >> 2;
  const int task = blockIdx.x * 256 + threadIdx.x;
  if (task >= G * A.H) return;
  const int oy = task / G, x = 4 * (task - oy * G), c = blockIdx.y, n = blockIdx.z;
  const MixPath& P = A.p[0];
  const uint16_t* src = reinterpret_cast<const uint16_t*>(P.src) + ((size_t)n * P.C + c) * (size_t)P.H * P.W + (size_t)(2 * oy) * P.W + 2 * x;
  const uint4 r0 = *reinterpret_cast<const uint4*>(src), r1 = *reinterpret_cast<const uint4*>(src + P.W);
  const uint32_t a[4] = {r0.x, r0.y, r0.z, r0.w}, b[4] = {r1.x, r1.y, r1.z, r1.w};
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 u = Pack<T>::to_f2(a[k]), v = Pack<T>::to_f2(b[k]);
    o[k] = is_max ? fmaxf(fmaxf(u.x, u.y), fmaxf(v.x, v.y)) : (((u.x + u.y) + v.x) + v.y) * 0.25f;
  }
  uint2 out;
  out.x = Pack<T>::from_f2(o[0], o[1]);
  out.y = Pack<T>::from_f2(o[2], o[3]);
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(A.dst) + ((size_t)n * A.C + c) * (size_t)A.H * A.W + (size_t)oy * A.W + x) = out;
}

// F.interpolate(scale_factor=up, bilinear, align_corners=False) of a 16-bit tensor into a 16-bit tensor (the materialised
// input of an up-sampling 1x1 path): 4 output pixels of one row per thread, the row pair and its weights computed once, the
// same blend expression as bilinear_up(), one 8-byte store.  A.W % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256) upsample_fast_kernel(const __grid_constant__ MixArgs A) {
  const int G = A.W >> 2;
  const int task = blockIdx.x * 256 + threadIdx.x;
  if (task >= G * A.H) return;
  const int oy = task / G, x = 4 * (task - oy * G), c = blockIdx.y, n = blockIdx.z;
  const MixPath& P = A.p[0];
  const int Hs = P.H, Ws = P.W;
  const float inv = 1.0f / (float)P.up;
  float sy = ((float)oy + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  const int y0 = (int)sy, y1 = y0 + (y0 < Hs - 1 ? 1 : 0);
  const float ly = sy - (float)y0, hy = 1.f - ly;
  const uint16_t* plane = reinterpret_cast<const uint16_t*>(P.src) + ((size_t)n * P.C + P.c0 + c) * (size_t)Hs * Ws;
  const uint16_t *r0 = plane + (size_t)y0 * Ws, *r1 = plane + (size_t)y1 * Ws;
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float sx = ((float)(x + k) + 0.5f) * inv - 0.5f;
    sx = sx < 0.f ? 0.f : sx;
    const int x0 = (int)sx, x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
    const float lx = sx - (float)x0, hx = 1.f - lx;
    o[k] = hy * (hx * Pack<T>::to_f(r0[x0]) + lx * Pack<T>::to_f(r0[x1])) + ly * (hx * Pack<T>::to_f(r1[x0]) + lx * Pack<T>::to_f(r1[x1]));
  }
  uint2 out;
  out.x = Pack<T>::from_f2(o[0], o[1]);
  out.y = Pack<T>::from_f2(o[2], o[3]);
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(A.dst) + ((size_t)n * A.C + c) * (size_t)A.H * A.W + (size_t)oy * A.W + x) = out;
}

}  // namespace csnet

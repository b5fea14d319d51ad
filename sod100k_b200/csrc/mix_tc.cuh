// mix_tc.cuh — tensor-core executor of a CSNET_OP_MIX op with 16-bit operands (same op semantics as the
// generic kernel in generic_ops.cuh; reference call sites: gOctaveConv.forward csnet.py:664-726, MSBlock.forward
// :141-149, cls_layer :381).
//
// Implicit GEMM with the output pixels as the N dimension: D[cout][pixel] += W[cout][(ci,tap)] . X[(ci,tap)][pixel].
// A CTA owns an 8 x 32 tile of output pixels of one image and ALL output channels (MT m16 tiles); each of its 8
// warps owns one tile row (four n8 pixel tiles).  K is walked per conv path in chunks of 8 input channels:
//   stage   Xs[8][XH x XW]  = the chunk's input window at CONV resolution: the reference's avg_pool2d / max_pool2d
//           pre-ops are applied once per staged element (not per tap), zero outside the image (= conv padding);
//           consecutive paths that read the same source slice with the same pre-ops (the five dilations of an
//           MSBlock) share one staged window with the largest halo.
//           Ws[tap][MT*16][8] = the chunk's weights, fp32 blob -> 16-bit, zero rows outside the path's cout slice
//   mma     per tap: B fragment = two 16-bit loads (channels 2t, 2t+1 at pixel g, shifted by the tap offset — any
//           dilation works since no ldmatrix alignment is involved), A fragment = two 32-bit loads, mma.sync m16n8k8
//   epilogue resample-add paths (bilinear from the low-resolution scratch), bias, PReLU, store (16-bit or fp32).
#pragma once
#include "generic_ops.cuh"
#include "il_block.cuh"

namespace csnet {

constexpr int kTcThreads = 256;
constexpr int kTcTH = 8, kTcTW = 32;

struct TcGeom {
  int32_t tiles_x;
  int32_t xs_halves;      // allocated halves per staged channel plane (max over groups)
  int32_t kc;             // input channels per staged chunk (8, 16 or 32)
};

// Staged window of a conv group with halo `pad`: rows [oy0-pad, oy0+8+pad), columns [ox0-padL, ox0-padL+XW) with the
// left halo rounded up to 4 pixels so every row is a whole number of 8-byte chunks (cp.async / vector friendly).
__host__ __device__ inline int tc_pad_left(int pad) { return (pad + 3) & ~3; }
__host__ __device__ inline int tc_xw(int pad) { return (tc_pad_left(pad) + kTcTW + pad + 3) & ~3; }
__host__ __device__ inline int tc_plane_halves(int pad) {
  int n = (kTcTH + 2 * pad) * tc_xw(pad);
  n = (n + 15) / 16 * 16 + 8;          // == 8 (mod 16): the four channel pairs of a B fragment hit distinct banks
  return n;
}
__host__ __device__ inline int tc_wrow(int kc) { return kc + 8; }   // padded Ws row: conflict-free A-fragment loads

// Value a conv path sees at conv-grid position (cy, cx): pooled source, or (1x1 paths with up > 1) the source
// bilinearly up-sampled first — conv1x1(up(x)) == up(conv1x1(x)), the order gOctaveConv uses (csnet.py:702-707).
__device__ __forceinline__ float tc_fetch(const MixPath& P, int64_t plane, int cy, int cx) {
  if (P.up > 1) return bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, cy, cx);
  return fetch_pooled(P, plane, cy, cx);
}

template <typename T, int MT>
__global__ void __launch_bounds__(kTcThreads, 2) mix_tc_kernel(const __grid_constant__ MixArgs A, const TcGeom G) {
  extern __shared__ __align__(16) uint16_t tc_smem[];
  const int KC = G.kc, WR = tc_wrow(KC);
  uint16_t* Xs = tc_smem;                                  // [KC][xs_halves]
  uint16_t* Ws = tc_smem + KC * G.xs_halves;               // [9][MT*16][WR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.z;
  const int oy0 = (blockIdx.x / G.tiles_x) * kTcTH, ox0 = (blockIdx.x % G.tiles_x) * kTcTW;
  constexpr int M16 = MT * 16;

  float acc[MT][4][4];
#pragma unroll
  for (int a = 0; a < MT; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  for (int p0 = 0; p0 < A.n_paths;) {
    const MixPath& P0 = A.p[p0];
    if (P0.ksize == 0) { ++p0; continue; }
    // group = consecutive conv paths reading the same source slice with the same pre-ops
    int p1 = p0 + 1, pad = P0.pad;
    while (p1 < A.n_paths) {
      const MixPath& Q = A.p[p1];
      if (Q.ksize == 0 || Q.src != P0.src || Q.c0 != P0.c0 || Q.cin != P0.cin || Q.pre_avg != P0.pre_avg ||
          Q.pool != P0.pool || Q.up != P0.up)
        break;
      pad = Q.pad > pad ? Q.pad : pad;
      ++p1;
    }
    const int XH = kTcTH + 2 * pad, XW = tc_xw(pad), padL = tc_pad_left(pad), PS = tc_plane_halves(pad);
    const int div = (P0.pre_avg ? 2 : 1) * P0.pool;
    const int Hc = P0.up > 1 ? P0.H * P0.up : P0.H / div, Wc = P0.up > 1 ? P0.W * P0.up : P0.W / div;
    const int64_t plane_sz = (int64_t)P0.H * P0.W;
    const bool plain = !P0.pre_avg && P0.pool == 1 && P0.up == 1 && P0.dtype != DT_F32;   // raw 16-bit copy
    const bool vec = plain && (P0.W & 3) == 0;
    for (int c0 = 0; c0 < P0.cin; c0 += KC) {
      const int kc_live = (P0.cin - c0) < KC ? (P0.cin - c0) : KC;
      const int kc8 = (kc_live + 7) & ~7;                   // channels actually multiplied (multiple of 8, rest zero)
      __syncthreads();                                     // previous chunk's readers are done
      const uint16_t* src16 = reinterpret_cast<const uint16_t*>(P0.src) + ((int64_t)n * P0.C + P0.c0 + c0) * plane_sz;
      if (vec) {
        // ---- (a) 8-byte cp.async chunks, zero fill outside the image / beyond the live channels ----------------
        const int cpr = XW >> 2, total = kc8 * XH * cpr;
        int i = tid;
        int row = i / cpr, col = i - row * cpr;            // one division per thread, then incremental
        const int drow = kTcThreads / cpr, dcol = kTcThreads - drow * cpr;
        for (; i < total; i += kTcThreads) {
          const int ch = row / XH, y = row - ch * XH;
          const int cy = oy0 - pad + y, cx = ox0 - padL + 4 * col;
          const bool ok = ch < kc_live && cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;
          cp_async8(Xs + ch * PS + y * XW + 4 * col, ok ? src16 + (int64_t)ch * plane_sz + (int64_t)cy * P0.W + cx : src16, ok);
          row += drow; col += dcol;
          if (col >= cpr) { col -= cpr; ++row; }
        }
        cp_async_wait_all();
      } else {
        // ---- (b)/(c) one (channel, row) per warp step, 4 steps batched so the loads overlap --------------------
        const int nrows = kc8 * XH;
        for (int rt0 = warp * 4; rt0 < nrows; rt0 += (kTcThreads / 32) * 4) {
          for (int xb = 0; xb < XW; xb += 32) {
            const int x = xb + lane, cx = ox0 - padL + x;
            const bool col_ok = x < XW && cx >= 0 && cx < Wc;
            float v[4];
            uint16_t raw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int rt = rt0 + q, ch = rt / XH, y = rt - ch * XH, cy = oy0 - pad + y;
              const bool ok = col_ok && rt < nrows && ch < kc_live && cy >= 0 && cy < Hc;
              v[q] = 0.f; raw[q] = 0;
              if (ok) {
                if (plain) raw[q] = __ldg(src16 + (int64_t)ch * plane_sz + (int64_t)cy * P0.W + cx);
                else v[q] = tc_fetch(P0, ((int64_t)n * P0.C + P0.c0 + c0 + ch) * plane_sz, cy, cx);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int rt = rt0 + q, ch = rt / XH, y = rt - ch * XH;
              if (rt < nrows && x < XW)
                Xs[ch * PS + y * XW + x] = plain ? raw[q] : (uint16_t)(Pack<T>::from_f2(v[q], 0.f) & 0xffffu);
            }
          }
        }
      }
      for (int p = p0; p < p1; ++p) {
        const MixPath& P = A.p[p];
        const int kk = P.ksize * P.ksize;
        if (p > p0) __syncthreads();                       // Ws of the previous path is no longer read
        // ---- stage this path's weights for the chunk: Ws[tap][m][k] ----------------------------------
        for (int i = tid; i < kk * M16 * kc8; i += kTcThreads) {
          const int k = i % kc8, r = i / kc8;
          const int m = r % M16, tap = r / M16;
          float w = 0.f;
          if (m >= P.cout0 && m < P.cout0 + P.cout && k < kc_live)
            w = __ldg(P.w + ((int64_t)(c0 + k) * kk + tap) * P.cout + (m - P.cout0));
          Ws[(tap * M16 + m) * WR + k] = (uint16_t)(Pack<T>::from_f2(w, 0.f) & 0xffffu);
        }
        __syncthreads();
        // ---- tensor-core accumulate --------------------------------------------------------------------
        const int off = pad - P.pad, offx = padL - P.pad;   // this path's window sits inside the staged one
        for (int ks = 0; ks < kc8; ks += 8) {
          const uint16_t* x0 = Xs + (ks + 2 * t) * PS + (warp + off) * XW + g + offx;
          for (int ky = 0; ky < P.ksize; ++ky) {
            for (int kx = 0; kx < P.ksize; ++kx) {
              const uint16_t* wt = Ws + ((ky * P.ksize + kx) * M16 + g) * WR + ks + 2 * t;
              uint32_t af[MT][2];
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                af[mt][0] = *reinterpret_cast<const uint32_t*>(wt + mt * 16 * WR);
                af[mt][1] = *reinterpret_cast<const uint32_t*>(wt + (mt * 16 + 8) * WR);
              }
              const uint16_t* xt = x0 + (ky * P.dil) * XW + kx * P.dil;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t b = (uint32_t)xt[j * 8] | ((uint32_t)xt[PS + j * 8] << 16);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) Pack<T>::mma(acc[mt][j], af[mt], b);
              }
            }
          }
        }
      }
    }
    p0 = p1;
  }

  // ---- epilogue ------------------------------------------------------------------------------------------
  const int oy = oy0 + warp;
  if (oy >= A.H) return;
  const int64_t out_plane = (int64_t)A.H * A.W;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = mt * 16 + g + 8 * h;
      if (m >= A.C) continue;
      const float bias = A.bias ? __ldg(A.bias + m) : 0.f;
      const bool has_slope = A.slope != nullptr;
      const float slope = has_slope ? __ldg(A.slope + m) : 1.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + j * 8 + 2 * t;
        if (ox >= A.W) continue;
        float v0 = acc[mt][j][2 * h], v1 = acc[mt][j][2 * h + 1];
        for (int p = 0; p < A.n_paths; ++p) {
          const MixPath& P = A.p[p];
          if (P.ksize != 0 || m < P.cout0 || m >= P.cout0 + P.cout) continue;
          const int64_t plane = ((int64_t)n * P.C + P.c0 + (m - P.cout0)) * (int64_t)P.H * P.W;
          v0 += bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox);
          if (ox + 1 < A.W) v1 += bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox + 1);
        }
        v0 += bias; v1 += bias;
        if (has_slope) { v0 = prelu(v0, slope); v1 = prelu(v1, slope); }
        const int64_t o = ((int64_t)n * A.C + m) * out_plane + (int64_t)oy * A.W + ox;
        if (A.dtype != DT_F32 && (A.W & 1) == 0) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(A.dst) + o) = Pack<T>::from_f2(v0, v1);
        } else {
          st_elem(A.dst, A.dtype, o, v0);
          if (ox + 1 < A.W) st_elem(A.dst, A.dtype, o + 1, v1);
        }
      }
    }
  }
}

}  // namespace csnet

// mix_tc.cuh — tensor-core executor of a CSNET_OP_MIX op with 16-bit operands (same op semantics as the
// generic kernel in generic_ops.cuh; reference call sites: gOctaveConv.forward csnet.py:664-726, MSBlock.forward
// :141-149, cls_layer :381).
//
// Implicit GEMM with the output pixels as the N dimension: D[cout][pixel] += W[cout][(ci,tap)] . X[(ci,tap)][pixel].
// A CTA owns an 8 x 32 tile of output pixels of one image and ALL output channels (MT m16 tiles); each of its 8
// warps owns one tile row (four n8 pixel tiles).  K is walked per conv path in chunks of KC (8/16/32) input channels:
//   stage   Xs[KC][XH x XW] = the chunk's input window at CONV resolution.  Plain 16-bit sources are copied with
//           8-byte cp.async (zero fill outside the image = conv padding); pooled / averaged / up-sampled / fp32
//           sources are computed once per staged element (not per tap).  Consecutive paths that read the same
//           source slice with the same pre-ops (the five dilations of an MSBlock) share one window with the
//           largest halo.  All staging loops are division-free: a warp owns whole channels, lanes run along rows.
//           Ws[tap][MT*16][KC+8] = the chunk's weights, fp32 blob -> 16-bit, zero outside the path's cout slice
//   mma     per tap: B fragment = two 16-bit loads (channels 2t, 2t+1 at pixel g, shifted by the tap offset — any
//           dilation works since no ldmatrix alignment is involved), A fragment = two 32-bit loads, mma.sync m16n8k8
//   epilogue resample-add paths (bilinear from the low-resolution scratch), bias, PReLU, store (16-bit or fp32).
#pragma once
#include "generic_ops.cuh"
#include "il_block.cuh"

namespace csnet {

constexpr int kTcThreads = 256;
constexpr int kTcWarps = kTcThreads / 32;
constexpr int kTcTH = 8, kTcTW = 32;

struct TcGeom {
  int32_t tiles_x;
  int32_t rows;           // output rows per warp (R): the tile is 8*R x 32.  R > 1 for wide-halo (dilated) ops: an 8-row
                          // tile with a 16-pixel halo stages 11x the pixels it produces, a 32-row tile 4x
  int32_t xs_halves;      // allocated halves per staged channel plane (max over groups)
  int32_t kc;             // input channels per staged chunk (8, 16 or 32)
  int32_t m16_total;      // output channels rounded up to the CTA slice (MT*16) multiple
  // 16-bit weights of every conv path, packed at csnet_plan_set_blob time as [chunk][tap][m16_total][kc + 8]
  // (zero outside the path's cout slice / beyond cin): a CTA's slice of one (chunk, tap) is contiguous.
  const uint16_t* w16[kMaxPaths];
};

// Staged window of a conv group with halo `pad`: rows [oy0-pad, oy0+8+pad); columns start at ox0-padL with the left
// halo rounded up to 4 pixels so every row is a whole number of 8-byte chunks.  cp.async windows are 32 or 64 wide
// (power of two: shift/mask indexing); computed windows are exactly as wide as needed.
__host__ __device__ inline int tc_pad_left(int pad) { return (pad + 3) & ~3; }
__host__ __device__ inline int tc_xw_vec(int pad) { return pad == 0 ? 32 : 64; }
__host__ __device__ inline int tc_xw_exact(int pad) { return (tc_pad_left(pad) + kTcTW + pad + 3) & ~3; }
__host__ __device__ inline int tc_plane_halves(int pad, int rows = 1) {
  int n = (kTcTH * rows + 2 * pad) * tc_xw_vec(pad);
  n = (n + 15) / 16 * 16 + 8;          // == 8 (mod 16): the four channel pairs of a B fragment hit distinct banks
  return n;
}
__host__ __device__ inline int tc_wrow(int kc) { return kc + 8; }   // padded Ws row: conflict-free A-fragment loads
constexpr int kTcMaxPad = 16;                                          // 3x3 with dilation <= 16

// Value a conv path sees at conv-grid position (cy, cx): pooled source, or (1x1 paths with up > 1) the source
// bilinearly up-sampled first — conv1x1(up(x)) == up(conv1x1(x)), the order gOctaveConv uses (csnet.py:702-707).
__device__ __forceinline__ float tc_fetch(const MixPath& P, int64_t plane, int cy, int cx) {
  if (P.up > 1) return bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, cy, cx);
  return fetch_pooled(P, plane, cy, cx);
}

// F.interpolate(bilinear, align_corners=False) by an integer factor from a 16-bit plane: same arithmetic as bilinear_up()
// without the per-load dtype dispatch.
template <typename T>
__device__ __forceinline__ float bilinear_up16(const uint16_t* plane, int Hs, int Ws, int up, int oy, int ox) {
  const float inv = 1.0f / (float)up;
  float sy = ((float)oy + 0.5f) * inv - 0.5f, sx = ((float)ox + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const uint16_t *r0 = plane + (size_t)y0 * Ws, *r1 = plane + (size_t)y1 * Ws;
  return hy * (hx * Pack<T>::to_f(r0[x0]) + lx * Pack<T>::to_f(r0[x1])) + ly * (hx * Pack<T>::to_f(r1[x0]) + lx * Pack<T>::to_f(r1[x1]));
}

// x2 up-sample of an fp32 plane for the output pixel pair (ox even, ox+1): fixed taps (1/4, 3/4), clamped indices —
// the same values bilinear_up() produces for up == 2.
__device__ __forceinline__ void bilinear_up2_pair_f32(const float* plane, int Hs, int Ws, int oy, int ox, float& v0, float& v1) {
  const int i = oy >> 1, j = ox >> 1;
  const int ya = (oy & 1) ? i : (i > 0 ? i - 1 : 0), yb = (oy & 1) ? (i + 1 < Hs ? i + 1 : Hs - 1) : i;
  const float wa = (oy & 1) ? 0.75f : 0.25f, wb = 1.f - wa;
  const int jm = j > 0 ? j - 1 : 0, jp = j + 1 < Ws ? j + 1 : Ws - 1;
  const float *ra = plane + (size_t)ya * Ws, *rb = plane + (size_t)yb * Ws;
  const float am = ra[jm], a0 = ra[j], ap = ra[jp], bm = rb[jm], b0 = rb[j], bp = rb[jp];
  // even column 2j: (1/4) src[j-1] + (3/4) src[j];  odd column 2j+1: (3/4) src[j] + (1/4) src[j+1]
  v0 += wa * (0.25f * am + 0.75f * a0) + wb * (0.25f * bm + 0.75f * b0);
  v1 += wa * (0.75f * a0 + 0.25f * ap) + wb * (0.75f * b0 + 0.25f * bp);
}

// Up-sampling by a multiple of 4 of an fp32 plane for the output pixel pair (ox even, ox+1): both pixels fall between the same
// two source columns (and the same two rows), so the four source values are loaded once; the blend is bilinear_up()'s
// expression with its own weights, i.e. the same values.
__device__ __forceinline__ void bilinear_up4n_pair_f32(const float* plane, int Hs, int Ws, int up, int oy, int ox, float& v0, float& v1) {
  const float inv = 1.0f / (float)up;
  float sy = ((float)oy + 0.5f) * inv - 0.5f, sa = ((float)ox + 0.5f) * inv - 0.5f, sb = ((float)(ox + 1) + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sa = sa < 0.f ? 0.f : sa;
  sb = sb < 0.f ? 0.f : sb;
  const int y0 = (int)sy, x0 = (int)sa;                    // (int)sb == x0 for even ox and up % 4 == 0
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly = sy - (float)y0, hy = 1.f - ly;
  const float la = sa - (float)x0, ha = 1.f - la, lb = sb - (float)x0, hb = 1.f - lb;
  const float *r0 = plane + (size_t)y0 * Ws, *r1 = plane + (size_t)y1 * Ws;
  const float v00 = r0[x0], v01 = r0[x1], v10 = r1[x0], v11 = r1[x1];
  v0 += hy * (ha * v00 + la * v01) + ly * (ha * v10 + la * v11);
  v1 += hy * (hb * v00 + lb * v01) + ly * (hb * v10 + lb * v11);
}

// Same as fetch_pooled() for a 16-bit source of type T with an even row length: pixel pairs come in as one
// 32-bit load (a 2x2 average = 2 loads, a 2x2 max of plain pixels = 2 loads), no per-load dtype dispatch.
template <typename T>
__device__ __forceinline__ float fetch_pooled16(const MixPath& P, const uint16_t* plane, int y, int x) {
  const int W = P.W, pool = P.pool;
  float m = -INFINITY;
  if (P.pre_avg) {
    const int f = pre_factor(P.pre_avg), fo = (f >> 1) - 1;   // fo is even (0, 0... ) only for f = 2; handled below
    for (int py = 0; py < pool; ++py) {
      const uint16_t* r = plane + (size_t)(f * (y * pool + py) + fo) * W + f * (x * pool) + fo;
      for (int px = 0; px < pool; ++px) {
        const float2 a = Pack<T>::to_f2(*reinterpret_cast<const uint32_t*>(r + f * px));
        const float2 b = Pack<T>::to_f2(*reinterpret_cast<const uint32_t*>(r + W + f * px));
        const float v = (((a.x + a.y) + b.x) + b.y) * 0.25f;
        m = v > m ? v : m;
      }
    }
  } else {                                                  // pool is 2, 4, 8: even, rows of pool pixels are 4-byte aligned
    for (int py = 0; py < pool; ++py) {
      const uint16_t* r = plane + (size_t)(y * pool + py) * W + x * pool;
      for (int px = 0; px < pool; px += 2) {
        const float2 a = Pack<T>::to_f2(*reinterpret_cast<const uint32_t*>(r + px));
        m = fmaxf(m, fmaxf(a.x, a.y));
      }
    }
  }
  return m;
}

template <typename T, int MT, int R = 1>
__global__ void __launch_bounds__(kTcThreads, (MT * R <= 2 ? 3 : 2)) mix_tc_kernel(const __grid_constant__ MixArgs A, const TcGeom G) {
  extern __shared__ __align__(16) uint16_t tc_smem[];
  const int KC = G.kc, WR = tc_wrow(KC);
  uint16_t* Xs = tc_smem;                                  // [KC][xs_halves]
  uint16_t* Ws = tc_smem + KC * G.xs_halves;               // [9][MT*16][WR]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.z;
  constexpr int TH = kTcTH * R;                             // warp w owns output rows w, w + 8, ... of the tile
  const int oy0 = (blockIdx.x / G.tiles_x) * TH, ox0 = (blockIdx.x % G.tiles_x) * kTcTW;
  constexpr int M16 = MT * 16;
  const int m_base = blockIdx.y * M16;                      // this CTA's slice of the output channels

  float acc[R][MT][4][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][a][b][c] = 0.f;

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
    const int cin = P0.cin, srcW = P0.W;
    const int div = pre_factor(P0.pre_avg) * P0.pool;
    const int Hc = P0.up > 1 ? P0.H * P0.up : P0.H / div, Wc = P0.up > 1 ? P0.W * P0.up : P0.W / div;
    const int64_t plane_sz = (int64_t)P0.H * P0.W;
    const bool plain = !P0.pre_avg && P0.pool == 1 && P0.up == 1 && P0.dtype != DT_F32;   // raw 16-bit copy
    const bool vec = plain && (srcW & 3) == 0;
    // typed 32-bit pair loads need 4-byte alignment: true for f = 2 (offset 0); f = 4 / 8 start at odd pixels
    const bool pooled16 = P0.dtype != DT_F32 && P0.up == 1 && (P0.pre_avg || P0.pool > 1) && (srcW & 1) == 0 && pre_factor(P0.pre_avg) <= 2;
    const bool up16 = P0.dtype != DT_F32 && P0.up > 1;
    const int XH = TH + 2 * pad, padL = tc_pad_left(pad), PS = tc_plane_halves(pad, R);
    const int XW = vec ? tc_xw_vec(pad) : tc_xw_exact(pad);
    const int64_t src_base = ((int64_t)n * P0.C + P0.c0) * plane_sz;
    for (int c0 = 0; c0 < cin; c0 += KC) {
      const int kc_live = (cin - c0) < KC ? (cin - c0) : KC;
      const int kc8 = (kc_live + 7) & ~7;                   // channels actually multiplied (multiple of 8, rest zero)
      __syncthreads();                                     // previous chunk's readers are done
      if (vec) {
        // ---- (a) 8-byte cp.async chunks; a warp owns channels warp, warp+8, ...; lanes walk (row, chunk) ---------
        const int cs = pad == 0 ? 3 : 4, cmask = (1 << cs) - 1, items = XH << cs;
        const uint16_t* src16 = reinterpret_cast<const uint16_t*>(P0.src) + src_base + (int64_t)c0 * plane_sz;
        for (int ch = warp; ch < kc8; ch += kTcWarps) {
          const bool ch_ok = ch < kc_live;
          const uint16_t* sp = src16 + (int64_t)ch * plane_sz;
          uint16_t* dp = Xs + ch * PS;
          for (int i = lane; i < items; i += 32) {
            const int y = i >> cs, col = i & cmask;
            const int cy = oy0 - pad + y, cx = ox0 - padL + 4 * col;
            const bool ok = ch_ok && cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;
            cp_async8(dp + y * XW + 4 * col, ok ? sp + (int64_t)cy * srcW + cx : sp, ok);
          }
        }
        cp_async_wait_all();
      } else if (up16 && XH * XW <= kTcThreads) {
        // ---- (b) input-side bilinear up-sampling of a 16-bit source (1x1 paths, pad 0): one staged pixel per thread, its
        //      four source offsets and weights computed once, then a channel loop of 4 loads + the bilinear_up16() blend ----
        const int y = tid / XW, x = tid - y * XW;
        const int cy = oy0 - pad + y, cx = ox0 - padL + x;
        const bool ok = tid < XH * XW && cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;
        const float inv = 1.0f / (float)P0.up;
        float sy = ((float)cy + 0.5f) * inv - 0.5f, sx = ((float)cx + 0.5f) * inv - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int sy0 = ok ? (int)sy : 0, sx0 = ok ? (int)sx : 0;
        const int sy1 = sy0 + (sy0 < P0.H - 1 ? 1 : 0), sx1 = sx0 + (sx0 < P0.W - 1 ? 1 : 0);
        const float ly = sy - (float)sy0, lx = sx - (float)sx0, hy = 1.f - ly, hx = 1.f - lx;
        const int o00 = sy0 * srcW + sx0, o01 = sy0 * srcW + sx1, o10 = sy1 * srcW + sx0, o11 = sy1 * srcW + sx1;
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(P0.src) + src_base + (int64_t)c0 * plane_sz;
        uint16_t* dp = Xs + y * XW + x;
        if (tid < XH * XW) {
#pragma unroll 4
          for (int ch = 0; ch < kc8; ++ch) {
            float v = 0.f;
            if (ok && ch < kc_live) {
              const uint16_t* pl = sp + (int64_t)ch * plane_sz;
              v = hy * (hx * Pack<T>::to_f(pl[o00]) + lx * Pack<T>::to_f(pl[o01])) + ly * (hx * Pack<T>::to_f(pl[o10]) + lx * Pack<T>::to_f(pl[o11]));
            }
            dp[ch * PS] = (uint16_t)(Pack<T>::from_f2(v, 0.f) & 0xffffu);
          }
        }
      } else {
        // ---- (c) scalar copy or computed (pool / avg / up-sample / fp32 source): 4 rows in flight per step ----
        for (int ch = warp; ch < kc8; ch += kTcWarps) {
          const bool ch_ok = ch < kc_live;
          const int64_t plane = src_base + (int64_t)(c0 + ch) * plane_sz;
          uint16_t* dp = Xs + ch * PS;
          for (int y0 = 0; y0 < XH; y0 += 4) {
            for (int xb = 0; xb < XW; xb += 32) {
              const int x = xb + lane, cx = ox0 - padL + x;
              const bool col_ok = ch_ok && x < XW && cx >= 0 && cx < Wc;
              float v[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int cy = oy0 - pad + y0 + q;
                v[q] = 0.f;
                if (col_ok && y0 + q < XH && cy >= 0 && cy < Hc)
                  v[q] = pooled16 ? fetch_pooled16<T>(P0, reinterpret_cast<const uint16_t*>(P0.src) + plane, cy, cx)
                         : up16 ? bilinear_up16<T>(reinterpret_cast<const uint16_t*>(P0.src) + plane, P0.H, P0.W, P0.up, cy, cx)
                                : tc_fetch(P0, plane, cy, cx);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (x < XW && y0 + q < XH) dp[(y0 + q) * XW + x] = (uint16_t)(Pack<T>::from_f2(v[q], 0.f) & 0xffffu);
            }
          }
        }
      }
      for (int p = p0; p < p1; ++p) {
        const MixPath& P = A.p[p];
        const int ksz = P.ksize, kk = ksz * ksz, dil = P.dil;
        if (p > p0) __syncthreads();                       // Ws of the previous path is no longer read
        // ---- stage this path's weights for the chunk: 16-byte cp.async copies of the pre-packed [tap][m][kc+8] slices ----
        {
          const int vec_row = WR >> 3, per_tap = M16 * vec_row;          // 16-byte vectors per weight row / per tap
          const uint16_t* wsrc = G.w16[p] + ((size_t)(c0 / KC) * kk * G.m16_total + m_base) * WR;
          for (int i = tid; i < kk * per_tap; i += kTcThreads) {
            const int tap = i / per_tap, r = i - tap * per_tap;
            cp_async16(Ws + (size_t)tap * M16 * WR + r * 8, wsrc + (size_t)tap * G.m16_total * WR + r * 8);
          }
          cp_async_wait_all();
        }
        __syncthreads();
        // ---- tensor-core accumulate --------------------------------------------------------------------
        const int off = pad - P.pad, offx = padL - P.pad;   // this path's window sits inside the staged one
        // One tap of one k-step: A fragments of every m tile, then per n tile a B register assembled from the two channel
        // planes of this lane's k pair (one PRMT) and MT mma.  All addresses advance by increments (no multiplies in the loop).
        auto tap = [&](const uint16_t* ra, const uint16_t* rb, const uint16_t* wt) {
          uint32_t af[MT][2];
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            af[mt][0] = *reinterpret_cast<const uint32_t*>(wt + mt * 16 * WR);
            af[mt][1] = *reinterpret_cast<const uint32_t*>(wt + (mt * 16 + 8) * WR);
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t b = __byte_perm((uint32_t)ra[r * 8 * XW + j * 8], (uint32_t)rb[r * 8 * XW + j * 8], 0x5410);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) Pack<T>::mma(acc[r][mt][j], af[mt], b);
            }
          }
        };
        const int rstep = dil * XW, wstep = M16 * WR;
        for (int ks = 0; ks < kc8; ks += 8) {
          const uint16_t* xa = Xs + (ks + 2 * t) * PS + (warp + off) * XW + g + offx;   // channel ks + 2t; + PS: ks + 2t + 1
          const uint16_t* wt = Ws + g * WR + ks + 2 * t;
          if (ksz == 1 && pad == 0) {
            // 1x1 group without halo: the staged rows are 16-byte aligned 8-pixel runs, so one ldmatrix.x4.trans delivers the
            // B fragments of all four n tiles of a row (instead of 8 16-bit loads + 4 PRMT)
            uint32_t af[MT][2];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              af[mt][0] = *reinterpret_cast<const uint32_t*>(wt + mt * 16 * WR);
              af[mt][1] = *reinterpret_cast<const uint32_t*>(wt + (mt * 16 + 8) * WR);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
              uint32_t bq[4];
              ldmatrix_x4_trans(bq, Xs + (ks + (lane & 7)) * PS + (warp + 8 * r) * XW + (lane >> 3) * 8);
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) Pack<T>::mma(acc[r][mt][j], af[mt], bq[j]);
            }
          } else if (ksz == 1) {
            tap(xa, xa + PS, wt);
          } else {
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky, xa += rstep) {
              const uint16_t* ra = xa;
#pragma unroll(MT * R <= 1 ? 3 : 1)
              for (int kx = 0; kx < 3; ++kx, ra += dil, wt += wstep) tap(ra, ra + PS, wt);
            }
          }
        }
      }
    }
    p0 = p1;
  }

  // ---- epilogue ------------------------------------------------------------------------------------------
  uint32_t rmask = 0;                                       // resample-add paths of this op
  for (int p = 0; p < A.n_paths; ++p)
    if (A.p[p].ksize == 0) rmask |= 1u << p;
  const int64_t out_plane = (int64_t)A.H * A.W;
  const bool pair_store = A.dtype != DT_F32 && (A.W & 1) == 0;
  const bool proj = A.proj_w != nullptr;                    // CSNET_OP_MIXPROJ: dot the channels with proj_w instead of storing
#pragma unroll
  for (int r = 0; r < R; ++r) {
  const int oy = oy0 + warp + 8 * r;
  if (oy >= A.H) break;                                     // warp-uniform
  float ps[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m_base + mt * 16 + g + 8 * h;
      if (m >= A.C) continue;
      const float bias = A.bias ? __ldg(A.bias + m) : 0.f;
      const bool has_slope = A.slope != nullptr;
      const float slope_m1 = has_slope ? __ldg(A.slope + m) - 1.f : 0.f;
      const float pw = proj ? __ldg(A.proj_w + m) : 0.f;
      const int64_t orow = ((int64_t)n * A.C + m) * out_plane + (int64_t)oy * A.W;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ox = ox0 + j * 8 + 2 * t;
        if (ox >= A.W) continue;
        float v0 = acc[r][mt][j][2 * h] + bias, v1 = acc[r][mt][j][2 * h + 1] + bias;
        for (uint32_t mk = rmask; mk; mk &= mk - 1) {
          const MixPath& P = A.p[__ffs(mk) - 1];
          if (m < P.cout0 || m >= P.cout0 + P.cout) continue;
          const int64_t plane = ((int64_t)n * P.C + P.c0 + (m - P.cout0)) * (int64_t)P.H * P.W;
          if (P.up == 2 && P.dtype == DT_F32 && ox + 1 < A.W) {
            bilinear_up2_pair_f32(reinterpret_cast<const float*>(P.src) + plane, P.H, P.W, oy, ox, v0, v1);
          } else if (P.up >= 4 && (P.up & 3) == 0 && P.dtype == DT_F32 && ox + 1 < A.W) {
            bilinear_up4n_pair_f32(reinterpret_cast<const float*>(P.src) + plane, P.H, P.W, P.up, oy, ox, v0, v1);
          } else {
            v0 += bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox);
            if (ox + 1 < A.W) v1 += bilinear_up(P.src, P.dtype, plane, P.H, P.W, P.up, oy, ox + 1);
          }
        }
        if (has_slope) { v0 = prelu_m1(v0, slope_m1); v1 = prelu_m1(v1, slope_m1); }
        if (proj) {
          ps[j][0] = fmaf(pw, v0, ps[j][0]);
          ps[j][1] = fmaf(pw, v1, ps[j][1]);
        } else if (pair_store) {
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(A.dst) + orow + ox) = Pack<T>::from_f2(v0, v1);
        } else {
          st_elem(A.dst, A.dtype, orow + ox, v0);
          if (ox + 1 < A.W) st_elem(A.dst, A.dtype, orow + ox + 1, v1);
        }
      }
    }
  }
  if (proj) {
    // the channel rows of a pixel live in the 8 lanes that share t (lane = 4 g + t): butterfly over g, lane g == 0 stores
    const float pb = A.proj_b ? __ldg(A.proj_b) : 0.f;
    const int64_t orow = (int64_t)n * out_plane + (int64_t)oy * A.W;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = ps[j][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int ox = ox0 + j * 8 + 2 * t + e;
        if (g == 0 && ox < A.W) st_elem(A.dst, A.dtype, orow + ox, v + pb);
      }
    }
  }
  }
}

}  // namespace csnet

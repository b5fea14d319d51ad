// TEST INFRASTRUCTURE — host emulation of the generic CUDA kernels.
//
// Compiles sod100k_b200/csrc/generic_ops.cuh (the exact per-thread bodies the GPU kernels run) for the CPU
// and walks a program the way plan.cu does: one "block" stages weights, then every (pixel, cout tile)
// thread runs.  fp32 tensors only.  Lets `-m "not gpu"` tests check kernel indexing (pooling, dilation,
// bilinear resample, channel slices, arena planning) against the oracle without a GPU.
#define CSNET_HOST_EMU 1
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/csnet_b200.h"
#include "../../sod100k_b200/csrc/generic_ops.cuh"

extern "C" int csnet_emu_run(const csnet_tensor_desc* tensors, int n_tensors, const csnet_op_desc* ops, int n_ops,
                             const float* blob, int N, void* const* ext, char* arena) {
  auto ptr = [&](int t) -> void* {
    const csnet_tensor_desc& d = tensors[t];
    return d.external >= 0 ? ext[d.external] : (void*)(arena + (int64_t)N * d.arena_offset);
  };
  for (int i = 0; i < n_tensors; ++i)
    if (tensors[i].dtype != CSNET_F32) return -4;
  for (int k = 0; k < n_ops; ++k) {
    const csnet_op_desc& op = ops[k];
    const csnet_tensor_desc& D = tensors[op.dst];
    if (op.kind == CSNET_OP_MIX) {
      csnet::MixArgs A{};
      A.dst = ptr(op.dst);
      A.bias = op.bias_off >= 0 ? blob + op.bias_off : nullptr;
      A.slope = op.slope_off >= 0 ? blob + op.slope_off : nullptr;
      A.dtype = D.dtype; A.C = D.C; A.H = D.H; A.W = D.W; A.n_paths = op.n_paths;
      for (int p = 0; p < op.n_paths; ++p) {
        const csnet_path_desc& q = op.paths[p];
        const csnet_tensor_desc& S = tensors[q.src];
        csnet::MixPath& m = A.p[p];
        m.src = ptr(q.src); m.w = q.ksize > 0 ? blob + q.w_off : nullptr;
        m.dtype = S.dtype; m.C = S.C; m.H = S.H; m.W = S.W; m.c0 = q.c0; m.cin = q.cin;
        m.pre_avg = q.pre_avg; m.pool = q.pool; m.ksize = q.ksize; m.dil = q.dil; m.stride = q.stride;
        m.pad = q.pad; m.up = q.up; m.cout0 = q.cout0; m.cout = q.cout;
      }
      const int64_t npix = (int64_t)N * D.H * D.W;
      for (int co = 0; co < D.C; co += csnet::kMixCT) {
        std::vector<float> acc((size_t)npix * csnet::kMixCT, 0.f), ws(csnet::kMixStageFloats);
        for (int p = 0; p < A.n_paths; ++p) {
          const csnet::MixPath& P = A.p[p];
          if (P.ksize == 0 || !csnet::mix_path_live(P, co)) continue;
          const int chunk = csnet::mix_chunk_channels(P.ksize);
          for (int ci0 = 0; ci0 < P.cin; ci0 += chunk) {
            const int ci1 = ci0 + chunk < P.cin ? ci0 + chunk : P.cin;
            csnet::mix_stage_chunk(P, co, ci0, ci1, ws.data(), 0, 1);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < npix; ++i) {
              const int n = (int)(i / (D.H * D.W)), r = (int)(i % (D.H * D.W));
              csnet::mix_acc_chunk(P, ws.data(), ci0, ci1, n, r / D.W, r % D.W, &acc[(size_t)i * csnet::kMixCT]);
            }
          }
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < npix; ++i) {
          const int n = (int)(i / (D.H * D.W)), r = (int)(i % (D.H * D.W));
          csnet::mix_finish(A, n, r / D.W, r % D.W, co, &acc[(size_t)i * csnet::kMixCT]);
        }
      }
    } else if (op.kind == CSNET_OP_DW) {
      const csnet_path_desc& q = op.paths[0];
      csnet::DwArgs A{};
      A.src = ptr(q.src); A.dst = ptr(op.dst); A.w = blob + q.w_off;
      A.bias = op.bias_off >= 0 ? blob + op.bias_off : nullptr;
      A.slope = op.slope_off >= 0 ? blob + op.slope_off : nullptr;
      A.src_dtype = tensors[q.src].dtype; A.dst_dtype = D.dtype; A.C = D.C; A.H = D.H; A.W = D.W;
#pragma omp parallel for collapse(2) schedule(static)
      for (int n = 0; n < N; ++n)
        for (int c = 0; c < D.C; ++c)
          for (int oy = 0; oy < D.H; oy += csnet::kDwRows)
            for (int ox = 0; ox < D.W; ++ox) csnet::dw_thread(A, n, c, oy, ox);
    } else if (op.kind == CSNET_OP_GN) {
      // GroupNorm + PReLU in plain C++ (the GPU kernels gn_stats_kernel / gn_apply_kernel are checked on the GPU)
      const float* x = (const float*)ptr(op.paths[0].src);
      float* y = (float*)ptr(op.dst);
      const int groups = op.paths[0].up, cpg = D.C / groups, HW = D.H * D.W;
      const float *ga = blob + op.ext_off[0], *be = blob + op.ext_off[1], *sl = op.slope_off >= 0 ? blob + op.slope_off : nullptr;
      for (int n = 0; n < N; ++n)
        for (int g = 0; g < groups; ++g) {
          const int64_t base = ((int64_t)n * D.C + (int64_t)g * cpg) * HW, cnt = (int64_t)cpg * HW;
          double s = 0, q = 0;
          for (int64_t i = 0; i < cnt; ++i) s += x[base + i];
          const double mu = s / cnt;
          for (int64_t i = 0; i < cnt; ++i) { const double d = x[base + i] - mu; q += d * d; }
          const float r = 1.0f / sqrtf((float)(q / cnt) + 1e-5f);
          for (int c = g * cpg; c < (g + 1) * cpg; ++c)
            for (int i = 0; i < HW; ++i) {
              float v = (x[((int64_t)n * D.C + c) * HW + i] - (float)mu) * r * ga[c] + be[c];
              if (sl) v = v > 0.f ? v : sl[c] * v;
              y[((int64_t)n * D.C + c) * HW + i] = v;
            }
        }
    } else {
      return -1;
    }
  }
  return 0;
}

// kr_gemm_epi.cuh — the fused GEMM epilogue shared by the data-parallel (kr_gemm.cu) and the stream-K
// (kr_gemm_sk.cu) single-CTA kernels: 32 consecutive output columns of ONE output row, fp32 accumulators in
// registers -> bias -> GELU-tanh / per-frame gate / residual with the reference's 16-bit rounding points
// (nn.Linear output rounded, then each following elementwise op rounded: causal_model.py:433-435, :466-488)
// -> 16-bit (or fp32) store, columns >= n_split redirected to the second destination.
#pragma once
#include "kr_common.cuh"
#include "kr_ops.h"

namespace kr {

template <bool kBf16, int kEpi>
KR_DEVICE void gemm_epilogue_row32(float (&v)[32], const int row, const int col0, const GemmParams& p,
                                   const uint16_t* gate_row) {
    if (p.bias != nullptr) {
      const uint4* b4 = reinterpret_cast<const uint4*>(
          reinterpret_cast<const uint16_t*>(p.bias) + col0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 bb = __ldg(b4 + q);
        uint32_t w[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float2 f = kBf16 ? unpack_bf16x2(w[h]) : unpack_f16x2(w[h]);
          v[q * 8 + h * 2] += f.x;
          v[q * 8 + h * 2 + 1] += f.y;
        }
      }
    }
    if constexpr (kEpi == EPI_F32) {
      float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldc + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 f4 = make_float4(v[q * 4] * p.alpha, v[q * 4 + 1] * p.alpha,
                                v[q * 4 + 2] * p.alpha, v[q * 4 + 3] * p.alpha);
        reinterpret_cast<float4*>(o)[q] = f4;
      }
    } else {
      auto rnd = [](float x) -> float {
        return kBf16 ? __bfloat162float(__float2bfloat16_rn(x))
                     : __half2float(__float2half_rn(x));
      };
      if constexpr (kEpi == EPI_BIAS_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(rnd(v[j]));
      }
      if constexpr (kEpi == EPI_BIAS_GATE_RES) {
        const uint4* g4 = reinterpret_cast<const uint4*>(gate_row + col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 gg = __ldg(g4 + q);
          uint32_t w[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float2 f = kBf16 ? unpack_bf16x2(w[h]) : unpack_f16x2(w[h]);
            v[q * 8 + h * 2] = rnd(rnd(v[q * 8 + h * 2]) * f.x);
            v[q * 8 + h * 2 + 1] = rnd(rnd(v[q * 8 + h * 2 + 1]) * f.y);
          }
        }
      }
      if constexpr (kEpi == EPI_BIAS_GATE_RES || kEpi == EPI_BIAS_RES || kEpi == EPI_MUL) {
        const uint4* r4 = reinterpret_cast<const uint4*>(
            reinterpret_cast<const uint16_t*>(p.residual) + static_cast<size_t>(row) * p.ldr +
            col0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 rr = r4[q];
          uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float2 f = kBf16 ? unpack_bf16x2(w[h]) : unpack_f16x2(w[h]);
            if constexpr (kEpi == EPI_MUL) {
            v[q * 8 + h * 2] = f.x * rnd(v[q * 8 + h * 2]);
            v[q * 8 + h * 2 + 1] = f.y * rnd(v[q * 8 + h * 2 + 1]);
          } else if constexpr (kEpi == EPI_BIAS_RES) {
              v[q * 8 + h * 2] = f.x + rnd(v[q * 8 + h * 2]);
              v[q * 8 + h * 2 + 1] = f.y + rnd(v[q * 8 + h * 2 + 1]);
            } else {
              v[q * 8 + h * 2] = f.x + v[q * 8 + h * 2];
              v[q * 8 + h * 2 + 1] = f.y + v[q * 8 + h * 2 + 1];
            }
          }
        }
      }
      uint16_t* o = (p.out2 != nullptr && col0 >= p.n_split)
          ? reinterpret_cast<uint16_t*>(p.out2) + static_cast<size_t>(row) * p.ldc2 + (col0 - p.n_split)
          : reinterpret_cast<uint16_t*>(p.out) + static_cast<size_t>(row) * p.ldc + col0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        if (kBf16) {
          w.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
          w.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
          w.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
          w.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
        } else {
          w.x = pack_f16x2(v[q * 8 + 0], v[q * 8 + 1]);
          w.y = pack_f16x2(v[q * 8 + 2], v[q * 8 + 3]);
          w.z = pack_f16x2(v[q * 8 + 4], v[q * 8 + 5]);
          w.w = pack_f16x2(v[q * 8 + 6], v[q * 8 + 7]);
        }
        reinterpret_cast<uint4*>(o)[q] = w;
      }
    }
}

}  // namespace kr

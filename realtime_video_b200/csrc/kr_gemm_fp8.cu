// kr_gemm_fp8.cu — FP8 (e4m3) tcgen05 GEMM with per-tensor scales + the dynamic activation quantisation kernels.
//
//   out[M,N] = epilogue( (Aq[M,K] @ Wq[N,K]^T) * scale_a * scale_w + bias[N] )      Aq, Wq e4m3, out bf16
//
// Replaces, on the reference path, what `quantize_(transformer, Float8DynamicActivationFloat8WeightConfig(
// granularity=PerTensor()))` (release_server.py:179-182, `enable_fp8: true`) turns every nn.Linear into: torchao's
// dynamic per-tensor activation cast (amax over the whole activation -> scale = 448 / amax -> saturating e4m3 cast)
// followed by torch._scaled_mm (cuBLASLt FP8, fp32 accumulation, bf16 output) — here a hand-written
// tcgen05.mma.kind::f8f6f4 kernel with the same fused epilogues as the bf16 GEMM (bias, GELU, gate + residual, ...).
// torchao is absent from the image: the algorithm is restated from its published definition (oracle/fp8_oracle.py).
//
// Kernel structure = kr_gemm.cu (persistent 1 CTA/SM, TMA producer warp, one elected MMA thread, 4 epilogue warps,
// TMEM accumulators double-buffered), with byte-sized elements: a 128-byte swizzle row holds 128 k-elements, one
// MMA consumes K = 32 (32 bytes), so a 48 KB stage (A 128x128 B, W 256x128 B) feeds 4 MMAs of 128 clk — twice the
// FLOPs of the bf16 kernel per byte moved.
#include "kr_common.cuh"
#include "kr_gemm_epi.cuh"
#include "kr_ops.h"

#include <cuda_fp8.h>

namespace kr {

namespace {
constexpr int F8_BM = 128, F8_BN = 256, F8_BK = 128 /* elements = bytes */, F8_UK = 32;
constexpr int F8_THREADS = 192, F8_STAGES = 4;
constexpr int F8_A_BYTES = F8_BM * F8_BK, F8_B_BYTES = F8_BN * F8_BK, F8_STAGE_BYTES = F8_A_BYTES + F8_B_BYTES;
constexpr int F8_SMEM = F8_STAGES * F8_STAGE_BYTES + 1024 + 256;

KR_DEVICE void umma_ss_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
}  // namespace

template <int kEpi>
__global__ void __launch_bounds__(F8_THREADS, 1)
gemm_fp8_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p, const float* __restrict__ scale_a, const float scale_w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + F8_STAGES * F8_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F8_STAGES * F8_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + F8_STAGES;
  uint64_t* tmem_full = bars + 2 * F8_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + F8_BM - 1) / F8_BM;
  const int num_tiles = num_m * (p.N / F8_BN);
  const int num_k = (p.K + F8_BK - 1) / F8_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < F8_STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_base_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % num_m, n_blk = tile / num_m;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], F8_STAGE_BYTES);
          tma_load_2d(smem_a + stage * F8_A_BYTES, &tmap_a, &full_bar[stage], kb * F8_BK, m_blk * F8_BM);
          tma_load_2d(smem_b + stage * F8_B_BYTES, &tmap_b, &full_bar[stage], kb * F8_BK, n_blk * F8_BN);
          if (++stage == F8_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // kind::f8f6f4 instruction descriptor: D f32, A / B e4m3 (format 0), both K-major
      constexpr uint32_t idesc = make_idesc_ab(F8_BM, F8_BN, 0u, 0u, 0, 0);
      const uint64_t desc_hi = make_smem_desc(0, 16, 1024) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo_c = static_cast<uint32_t>(make_smem_desc(0, 16, 1024));
      const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * F8_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = desc_lo_c | (((a0 + stage * F8_A_BYTES) & 0x3FFFF) >> 4);
          const uint32_t b_lo = desc_lo_c | (((b0 + stage * F8_B_BYTES) & 0x3FFFF) >> 4);
#pragma unroll
          for (int k = 0; k < F8_BK / F8_UK; ++k)      // 32 bytes per k-step -> start address + 2 (x16 B)
            umma_ss_f8(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);
          if (++stage == F8_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const float s = __ldg(scale_a) * scale_w;       // dequantisation: per-tensor activation x weight scale
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int row = m_blk * F8_BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * F8_BN;
      const uint16_t* gate_row = nullptr;
      if constexpr (kEpi == EPI_BIAS_GATE_RES) {
        const int g = row_ok ? (row + p.row_offset) / p.rows_per_gate : 0;
        gate_row = reinterpret_cast<const uint16_t*>(p.gate) + static_cast<size_t>(g) * p.gate_stride;
      }
#pragma unroll 1
      for (int c = 0; c < F8_BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (c == F8_BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * s;
          gemm_epilogue_row32<true, kEpi>(v, row, n_blk * F8_BN + c * 32, p, gate_row);
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// dynamic per-tensor activation quantisation: amax -> scale -> saturating e4m3 cast
// ---------------------------------------------------------------------------
__global__ void fp8_amax_kernel(const uint16_t* __restrict__ x, long ld, int rows, int cols, float* amax) {
  float m = 0.f;
  const int vec_per_row = cols >> 3;
  const long total = static_cast<long>(rows) * vec_per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row);
    const uint4 u = *reinterpret_cast<const uint4*>(x + r * ld + c * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));   // m >= 0: int order = float order
}

__global__ void fp8_quant_kernel(const uint16_t* __restrict__ x, long ld, int rows, int cols, uint8_t* __restrict__ q,
                                 long ldq, float* state) {
  // state[0] = amax (input), state[1] = dequantisation scale amax / 448 (output, read by the GEMM epilogue)
  const float amax = fmaxf(state[0], 1e-12f);
  const float mul = 448.0f / amax;
  if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = amax / 448.0f;
  const int vec_per_row = cols >> 3;
  const long total = static_cast<long>(rows) * vec_per_row;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / vec_per_row;
    const int c = static_cast<int>(i - r * vec_per_row);
    const uint4 u = *reinterpret_cast<const uint4*>(x + r * ld + c * 8);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t out[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      const __nv_fp8x2_storage_t h = __nv_cvt_float2_to_fp8x2(make_float2(f.x * mul, f.y * mul), __NV_SATFINITE, __NV_E4M3);
      if (j & 1) out[j >> 1] |= static_cast<uint32_t>(h) << 16;
      else out[j >> 1] = static_cast<uint32_t>(h);
    }
    *reinterpret_cast<uint2*>(q + r * ldq + c * 8) = make_uint2(out[0], out[1]);
  }
}

int fp8_quantize(const void* x, long ld, int rows, int cols, void* q, long ldq, float* state, cudaStream_t stream) {
  if (cols % 8 != 0 || ld % 8 != 0 || ldq % 8 != 0 || rows <= 0) {
    set_last_error("fp8_quantize: unsupported rows=%d cols=%d ld=%ld ldq=%ld", rows, cols, ld, ldq);
    return KR_ERR_UNSUPPORTED_SHAPE;
  }
  cudaError_t e = cudaMemsetAsync(state, 0, 2 * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_last_error("fp8_quantize: memset failed: %s", cudaGetErrorString(e));
    return KR_ERR_CUDA;
  }
  const long total = static_cast<long>(rows) * (cols / 8);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  fp8_amax_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), ld, rows, cols, state);
  fp8_quant_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint16_t*>(x), ld, rows, cols,
                                            static_cast<uint8_t*>(q), ldq, state);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("fp8_quantize: launch failed: %s", cudaGetErrorString(e));
    return KR_ERR_CUDA;
  }
  return KR_OK;
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
template <int kEpi>
static int launch_fp8(const void* a, int lda, const void* w, int ldw, const GemmParams& p, const float* scale_a,
                      float scale_w, cudaStream_t stream) {
  CUtensorMap ta, tb;
  // byte tensors: [rows, K] uint8, box [128 | 256 rows, 128 bytes], SWIZZLE_128B
  const uint64_t da[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.M)};
  const uint64_t sa[1] = {static_cast<uint64_t>(lda)};
  const uint32_t ba[2] = {F8_BK, F8_BM};
  int rc = make_tmap_u8(&ta, a, 2, da, sa, ba, 128);
  if (rc != KR_OK) return rc;
  const uint64_t db[2] = {static_cast<uint64_t>(p.K), static_cast<uint64_t>(p.N)};
  const uint64_t sb[1] = {static_cast<uint64_t>(ldw)};
  const uint32_t bb[2] = {F8_BK, F8_BN};
  rc = make_tmap_u8(&tb, w, 2, db, sb, bb, 128);
  if (rc != KR_OK) return rc;
  auto kern = gemm_fp8_kernel<kEpi>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F8_SMEM);
    if (e != cudaSuccess) {
      set_last_error("gemm_fp8: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return KR_ERR_CUDA;
    }
    attr_set = true;
  }
  const int num_tiles = ((p.M + F8_BM - 1) / F8_BM) * (p.N / F8_BN);
  int grid = sm_count();
  if (grid > num_tiles) grid = num_tiles;
  kern<<<grid, F8_THREADS, F8_SMEM, stream>>>(ta, tb, p, scale_a, scale_w);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("gemm_fp8: launch failed: %s", cudaGetErrorString(e));
    return KR_ERR_CUDA;
  }
  return KR_OK;
}

int gemm_fp8_tn(int epi, const void* a, int lda, const void* w, int ldw, const GemmParams& p, const float* scale_a,
                float scale_w, cudaStream_t stream) {
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.N % F8_BN != 0 || p.K % 16 != 0 || lda % 16 != 0 || ldw % 16 != 0 ||
      p.ldc % 8 != 0) {
    set_last_error("gemm_fp8: unsupported shape M=%d N=%d K=%d (need N %% 256 == 0, K / ld %% 16 == 0)", p.M, p.N, p.K);
    return KR_ERR_UNSUPPORTED_SHAPE;
  }
  if ((epi == EPI_BIAS_GATE_RES || epi == EPI_BIAS_RES || epi == EPI_MUL) && p.residual == nullptr) {
    set_last_error("gemm_fp8: residual epilogue without residual pointer");
    return KR_ERR_INVALID_ARG;
  }
  if (scale_a == nullptr) {
    set_last_error("gemm_fp8: null activation scale");
    return KR_ERR_INVALID_ARG;
  }
  switch (epi) {
    case EPI_BIAS: return launch_fp8<EPI_BIAS>(a, lda, w, ldw, p, scale_a, scale_w, stream);
    case EPI_BIAS_GELU: return launch_fp8<EPI_BIAS_GELU>(a, lda, w, ldw, p, scale_a, scale_w, stream);
    case EPI_BIAS_GATE_RES: return launch_fp8<EPI_BIAS_GATE_RES>(a, lda, w, ldw, p, scale_a, scale_w, stream);
    case EPI_BIAS_RES: return launch_fp8<EPI_BIAS_RES>(a, lda, w, ldw, p, scale_a, scale_w, stream);
    default: set_last_error("gemm_fp8: unsupported epilogue %d", epi); return KR_ERR_INVALID_ARG;
  }
}

}  // namespace kr

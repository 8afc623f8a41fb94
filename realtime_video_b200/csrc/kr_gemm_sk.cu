// kr_gemm_sk.cu — stream-K variant of the persistent tcgen05 GEMM (kr_gemm.cu) for shapes whose output-tile
// count does not fill the 148 SMs evenly.
//
//   C[M,N] = epilogue( A[M,K] @ W[N,K]^T + bias[N] )        same operands / epilogues as kr_gemm.cu
//
// Why: the data-parallel kernels give every CTA whole 128x256 output tiles.  A sequence-parallel rank of the
// multi-GPU mode owns M = 4680/N token rows (585 at 8 GPUs): to_qkv is then 5 x 60 = 300 tiles = 2.03 waves of
// 148 CTAs (a third wave for 4 tiles), the o / cross-attention / ffn.2 projections are 100 tiles (0.68 of a wave).
// Here the unit of work is one (tile, 64-wide k-block) MMA iteration: the num_tiles * num_k iterations are cut
// into 148 equal contiguous ranges, so every SM runs the same number of tcgen05.mma k-blocks (+-1) whatever the
// tile count.  A range that starts or ends inside a tile yields a PARTIAL accumulator:
//   * every contributor of a split tile writes its fp32 partial (raw accumulator, 128 x 256) to its own slot of a
//     workspace in HBM/L2 (slot 2*cta + {0 first segment, 1 last segment}), fences, and bumps the tile's counter;
//   * the contributor that arrives last (counter == contributors - 1) re-reads all partials of the tile with
//     L1-bypassing loads, sums them in ascending CTA order (deterministic) and runs the fused epilogue.
// Tiles that one CTA covers completely take the data-parallel path (TMEM -> registers -> epilogue) untouched.
// Roles / pipelines are those of kr_gemm.cu: warp 0 TMA producer (128x64 A + 256x64 W per stage, SWIZZLE_128B,
// 4-stage ring), warp 1 one elected tcgen05.mma issuer, warps 2-5 epilogue, TMEM accumulators double-buffered.
#include "kr_common.cuh"
#include "kr_gemm_epi.cuh"
#include "kr_ops.h"

namespace kr {

namespace {
constexpr int SK_BM = 128, SK_BN = 256, SK_BK = 64, SK_UK = 16;
constexpr int SK_THREADS = 192;
constexpr int SK_STAGES = 4;
constexpr int SK_A_BYTES = SK_BM * SK_BK * 2, SK_B_BYTES = SK_BN * SK_BK * 2;
constexpr int SK_STAGE_BYTES = SK_A_BYTES + SK_B_BYTES;
constexpr int SK_SMEM = SK_STAGES * SK_STAGE_BYTES + 1024 + 256;
constexpr int SK_TMEM_COLS = 2 * SK_BN;
constexpr size_t SK_SLOT_FLOATS = static_cast<size_t>(SK_BM) * SK_BN;
constexpr size_t SK_COUNTER_BYTES = 64 * 1024;      // 16384 tile counters in front of the partial slots

struct SkPlan {
  int num_m, num_n, num_k, num_tiles;
  long long q;       // iterations per CTA (the first r CTAs run q + 1)
  int r;
  float* slots;      // [2 * grid][128][256] fp32
  int* counters;     // [num_tiles], zero between launches (the last arriver resets its tile)
};

KR_DEVICE long long sk_begin(const SkPlan& s, int cta) {
  return static_cast<long long>(cta) * s.q + (cta < s.r ? cta : s.r);
}
KR_DEVICE int sk_cta_of(const SkPlan& s, long long it) {
  const long long split = static_cast<long long>(s.r) * (s.q + 1);
  return it < split ? static_cast<int>(it / (s.q + 1)) : s.r + static_cast<int>((it - split) / s.q);
}
}  // namespace

template <bool kBf16, int kEpi>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p, const SkPlan s) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + SK_STAGES * SK_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SK_STAGES * SK_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + SK_STAGES;
  uint64_t* tmem_full = bars + 2 * SK_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_base_smem + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const long long it_begin = sk_begin(s, blockIdx.x), it_end = sk_begin(s, blockIdx.x + 1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < SK_STAGES; ++i) {
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
    tmem_alloc<SK_TMEM_COLS>(tmem_base_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long it = it_begin; it < it_end;) {
        const int tile = static_cast<int>(it / s.num_k);
        const int k0 = static_cast<int>(it - static_cast<long long>(tile) * s.num_k);
        const long long left = it_end - it;
        const int k1 = (s.num_k - k0) < left ? s.num_k : k0 + static_cast<int>(left);
        const int m_blk = tile % s.num_m, n_blk = tile / s.num_m;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], SK_STAGE_BYTES);
          tma_load_2d(smem_a + stage * SK_A_BYTES, &tmap_a, &full_bar[stage], kb * SK_BK, m_blk * SK_BM);
          tma_load_2d(smem_b + stage * SK_B_BYTES, &tmap_b, &full_bar[stage], kb * SK_BK, n_blk * SK_BN);
          if (++stage == SK_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        it += k1 - k0;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc<kBf16>(SK_BM, SK_BN, 0, 0);
      const uint64_t desc_hi = make_smem_desc(0, 16, 1024) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo_c = static_cast<uint32_t>(make_smem_desc(0, 16, 1024));
      const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long it = it_begin; it < it_end;) {
        const int tile = static_cast<int>(it / s.num_k);
        const int k0 = static_cast<int>(it - static_cast<long long>(tile) * s.num_k);
        const long long left = it_end - it;
        const int k1 = (s.num_k - k0) < left ? s.num_k : k0 + static_cast<int>(left);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * SK_BN;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = desc_lo_c | (((a0 + stage * SK_A_BYTES) & 0x3FFFF) >> 4);
          const uint32_t b_lo = desc_lo_c | (((b0 + stage * SK_B_BYTES) & 0x3FFFF) >> 4);
#pragma unroll
          for (int k = 0; k < SK_BK / SK_UK; ++k) {
            umma_ss(d_tmem, desc_hi | (a_lo + k * 2), desc_hi | (b_lo + k * 2), idesc,
                    (kb != k0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == k1 - 1) umma_commit(&tmem_full[acc]);
          if (++stage == SK_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        it += k1 - k0;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;
    const int etid = threadIdx.x - 64;          // 0..127 among the epilogue threads
    int acc = 0, seg = 0;
    uint32_t acc_phase = 0;
    for (long long it = it_begin; it < it_end; ++seg) {
      const int tile = static_cast<int>(it / s.num_k);
      const int k0 = static_cast<int>(it - static_cast<long long>(tile) * s.num_k);
      const long long left = it_end - it;
      const int k1 = (s.num_k - k0) < left ? s.num_k : k0 + static_cast<int>(left);
      it += k1 - k0;
      const int m_blk = tile % s.num_m, n_blk = tile / s.num_m;
      const int lrow = quarter * 32 + lane;
      const int row = m_blk * SK_BM + lrow;
      const bool row_ok = row < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * SK_BN;
      const uint16_t* gate_row = nullptr;
      if constexpr (kEpi == EPI_BIAS_GATE_RES) {
        const int g = row_ok ? (row + p.row_offset) / p.rows_per_gate : 0;
        gate_row = reinterpret_cast<const uint16_t*>(p.gate) + static_cast<size_t>(g) * p.gate_stride;
      }
      const bool whole = (k0 == 0 && k1 == s.num_k);
      if (whole) {
#pragma unroll 1
        for (int c = 0; c < SK_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (c == SK_BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          if (row_ok) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            gemm_epilogue_row32<kBf16, kEpi>(v, row, n_blk * SK_BN + c * 32, p, gate_row);
          }
        }
      } else {
        // ---- partial accumulator -> my workspace slot ----
        float* mine = s.slots + (static_cast<size_t>(2 * blockIdx.x + (seg == 0 ? 0 : 1)) * SK_SLOT_FLOATS) +
                      static_cast<size_t>(lrow) * SK_BN;
#pragma unroll 1
        for (int c = 0; c < SK_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(t_row + c * 32, r);
          tmem_ld_wait();
          if (c == SK_BN / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          float4* dst = reinterpret_cast<float4*>(mine + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(dst + j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
        __threadfence();
        // contributors of this tile: the CTAs whose ranges intersect [tile*num_k, (tile+1)*num_k)
        const long long t0 = static_cast<long long>(tile) * s.num_k;
        const int c_first = sk_cta_of(s, t0), c_last = sk_cta_of(s, t0 + s.num_k - 1);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (etid == 0) {
          const int old = atomicAdd(s.counters + tile, 1);
          const int last = (old == c_last - c_first) ? 1 : 0;
          if (last) s.counters[tile] = 0;            // ready for the next launch
          *last_flag = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const bool is_last = *last_flag != 0;
        asm volatile("bar.sync 1, 128;" ::: "memory");   // flag consumed before the next segment rewrites it
        if (is_last) {
          __threadfence();
          if (row_ok) {
#pragma unroll 1
            for (int c = 0; c < SK_BN / 32; ++c) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
              for (int cc = c_first; cc <= c_last; ++cc) {
                // tile is cc's first segment iff cc's range starts inside (or at the start of) this tile
                const int sl = 2 * cc + (sk_begin(s, cc) >= t0 ? 0 : 1);
                const float4* src = reinterpret_cast<const float4*>(
                    s.slots + static_cast<size_t>(sl) * SK_SLOT_FLOATS + static_cast<size_t>(lrow) * SK_BN + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 f = __ldcg(src + j);
                  v[4 * j] += f.x; v[4 * j + 1] += f.y; v[4 * j + 2] += f.z; v[4 * j + 3] += f.w;
                }
              }
              gemm_epilogue_row32<kBf16, kEpi>(v, row, n_blk * SK_BN + c * 32, p, gate_row);
            }
          }
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
    tmem_dealloc<SK_TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
size_t gemm_sk_workspace_bytes() {
  return SK_COUNTER_BYTES + static_cast<size_t>(2) * sm_count() * SK_SLOT_FLOATS * sizeof(float);
}

// Is stream-K worth it for this shape?  Compared on the data-parallel kernel's wave efficiency.
bool gemm_sk_preferred(int epi, int M, int N, int K) {
  if (epi == EPI_F32 || N % SK_BN != 0 || K < 8 * SK_BK) return false;
  const int sms = sm_count();
  const long tiles = static_cast<long>((M + SK_BM - 1) / SK_BM) * (N / SK_BN);
  if (tiles > SK_COUNTER_BYTES / 4) return false;
  const long waves = (tiles + sms - 1) / sms;
  const double eff = static_cast<double>(tiles) / static_cast<double>(waves * sms);
  return eff < 0.86;
}

template <bool kBf16, int kEpi>
static int launch_sk(const void* a, int lda, const void* w, int ldw, const GemmParams& p, void* ws,
                     cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, a, p.M, p.K, lda, SK_BM, SK_BK, kBf16);
  if (rc != KR_OK) return rc;
  rc = make_tmap_2d(&tb, w, p.N, p.K, ldw, SK_BN, SK_BK, kBf16);
  if (rc != KR_OK) return rc;
  auto kern = gemm_sk_kernel<kBf16, kEpi>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    if (e != cudaSuccess) {
      set_last_error("gemm_sk: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return KR_ERR_CUDA;
    }
    attr_set = true;
  }
  SkPlan s;
  s.num_m = (p.M + SK_BM - 1) / SK_BM;
  s.num_n = p.N / SK_BN;
  s.num_k = (p.K + SK_BK - 1) / SK_BK;
  s.num_tiles = s.num_m * s.num_n;
  const long long total = static_cast<long long>(s.num_tiles) * s.num_k;
  int grid = sm_count();
  if (grid > total) grid = static_cast<int>(total);
  s.q = total / grid;
  s.r = static_cast<int>(total % grid);
  s.counters = reinterpret_cast<int*>(ws);
  s.slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + SK_COUNTER_BYTES);
  kern<<<grid, SK_THREADS, SK_SMEM, stream>>>(ta, tb, p, s);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("gemm_sk: launch failed: %s", cudaGetErrorString(e));
    return KR_ERR_CUDA;
  }
  return KR_OK;
}

// workspace: gemm_sk_workspace_bytes() bytes, zero-filled once by the caller, private to one stream
int gemm_sk_tn(int dtype, int epi, const void* a, int lda, const void* w, int ldw, const GemmParams& p,
               void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (workspace == nullptr || workspace_bytes < gemm_sk_workspace_bytes()) {
    set_last_error("gemm_sk: workspace of %zu bytes required", gemm_sk_workspace_bytes());
    return KR_ERR_INVALID_ARG;
  }
  const bool bf = dtype == 0;
#define KR_SK(E) (bf ? launch_sk<true, E>(a, lda, w, ldw, p, workspace, stream) \
                     : launch_sk<false, E>(a, lda, w, ldw, p, workspace, stream))
  switch (epi) {
    case EPI_BIAS: return KR_SK(EPI_BIAS);
    case EPI_BIAS_GELU: return KR_SK(EPI_BIAS_GELU);
    case EPI_BIAS_GATE_RES: return KR_SK(EPI_BIAS_GATE_RES);
    case EPI_BIAS_RES: return KR_SK(EPI_BIAS_RES);
    case EPI_MUL: return KR_SK(EPI_MUL);
    default: set_last_error("gemm_sk: unsupported epilogue %d", epi); return KR_ERR_INVALID_ARG;
  }
#undef KR_SK
}

}  // namespace kr

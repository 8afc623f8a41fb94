// Experiment (B200): may a K-major SWIZZLE_128B UMMA operand start at a row that is NOT a multiple of 8
// (start address not 1024 B aligned), and which "matrix base offset" (descriptor bits 49..51) does the
// tensor core then expect?  This decides whether the VAE conv can issue its 3x3 spatial taps as shifted
// windows over ONE halo tile in shared memory instead of re-fetching the input per tap.
//   smem A: 160 rows x 64 bf16 (128 B rows), written with the swizzle TMA would apply (16-byte chunk
//   index ^= address bits 7..9), base 1024-aligned.  D[128 x 64] = A[s .. s+128) . B^T for row shifts s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I../../realtime_video_b200/csrc
//        -I../../include umma_shift_test.cu -o umma_shift_test
#include <cstdio>
#include <cuda_bf16.h>
#include "kr_common.cuh"

using namespace kr;

__device__ float a_val(int r, int c) { return static_cast<float>((r * 7 + c * 3) % 13 - 6); }
__device__ float b_val(int n, int c) { return static_cast<float>((n * 5 + c) % 7 - 3); }

__global__ void __launch_bounds__(128, 1) shift_kernel(int shift, int base_off, int sbo_rows, int* mismatches) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 512 rows x 128 B (enough for sbo_rows up to 24 with 16 groups)
  uint8_t* sb = smem + 512 * 128;     // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // physical row p of the A buffer holds logical matrix row p (value a_val(p, .))
  for (int i = tid; i < 512 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(a_val(row, ch * 8 + e));
    const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(sa + off) = *reinterpret_cast<uint4*>(v);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    __nv_bfloat16 v[8];
    for (int e = 0; e < 8; ++e) v[e] = __float2bfloat16(b_val(row, ch * 8 + e));
    const uint32_t off = row * 128 + ((ch ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(sb + off) = *reinterpret_cast<uint4*>(v);
  }
  if (warp == 0) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<64>(tmem_slot);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 && elect_one()) {
    constexpr uint32_t idesc = make_idesc<true>(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_smem_desc(smem_u32(sa) + shift * 128 + k * 32, 16, sbo_rows * 128);
      ad |= static_cast<uint64_t>(base_off & 7) << 49;
      const uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024);
      umma_ss(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t d[32];
  int bad = 0;
  const int m = tid;   // output row: group g = m / 8 lives sbo_rows physical rows after the previous one
  const int prow = shift + (m >> 3) * sbo_rows + (m & 7);
  for (int half = 0; half < 2; ++half) {
    tmem_ld_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * 32, d);
    tmem_ld_wait();
    for (int n = 0; n < 32; ++n) {
      float ref = 0.f;
      for (int c = 0; c < 64; ++c) ref += a_val(prow, c) * b_val(half * 32 + n, c);
      if (__uint_as_float(d[n]) != ref) ++bad;
    }
  }
  if (bad) atomicAdd(mismatches, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  int* dev;
  cudaMalloc(&dev, sizeof(int));
  const int smem = 512 * 128 + 64 * 128 + 1024 + 64;
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int sbos[3] = {8, 16, 24};
  for (int si = 0; si < 3; ++si)
    for (int shift = 0; shift < 4; ++shift)
      for (int bo = 0; bo < 4; ++bo) {
        if (bo != 0 && bo != shift) continue;
        int zero = 0, host = -1;
        cudaMemcpy(dev, &zero, sizeof(int), cudaMemcpyHostToDevice);
        shift_kernel<<<1, 128, smem>>>(shift, bo, sbos[si], dev);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&host, dev, sizeof(int), cudaMemcpyDeviceToHost);
        printf("sbo_rows=%2d shift=%d base_offset=%d : %s (%d mismatching of 8192) %s\n", sbos[si], shift, bo,
               host == 0 ? "EXACT" : "WRONG", host, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}

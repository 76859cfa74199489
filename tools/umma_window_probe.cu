// Hardware probe (developer tool, not part of the library): can a tcgen05.mma SW128 K-major A descriptor start at
// an arbitrary 128-byte row of a swizzled shared-memory region, and with a group stride (SBO) that is not a
// multiple of 1024 B?  The answer decides how the halo-strip convolution addresses its shifted tap windows
// (DESIGN.md "fused operand load").
//
// A region: 320 logical rows x 64 bf16, stored with the absolute-address 128B swizzle TMA uses
// (16-byte chunk c of row R lives at chunk c ^ (R & 7) when the region base is 1024-byte aligned).
// For every (start row r0, SBO, base_offset field) the kernel runs M=128 N=32 K=64 and the host compares with
//   D[m][n] = sum_k A[r0 + (m/8)*(SBO/128) + m%8][k] * B[n][k].
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_window_probe tools/umma_window_probe.cu
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../deepliif_b200/csrc/ptx.cuh"

using namespace dlb;

constexpr int kRows = 320;

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       int r0, int sbo, int base_off, float* __restrict__ D) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // kRows x 128 B
  uint8_t* sB = smem + kRows * 128;         // 32 x 128 B (kRows*128 is a multiple of 1024)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kRows * 8; i += 128) {
    const int R = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + R * 128 + ((c ^ (R & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + R * 64 + c * 8);
  }
  for (int i = tid; i < 32 * 8; i += 128) {
    const int R = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + R * 128 + ((c ^ (R & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + R * 64 + c * 8);
  }
  fence_proxy_async();
  if (tid == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_smem, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, 32, 1);
    const uint32_t a0 = smem_u32(sA) + r0 * 128, b0 = smem_u32(sB);
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem, make_desc(a0 + k * 32, sbo, base_off), make_desc(b0 + k * 32, 1024, 0), idesc, k > 0);
    umma_commit(&done_bar);
  }
  mbar_wait(&done_bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) D[tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

int main() {
  std::vector<__nv_bfloat16> hA(kRows * 64), hB(32 * 64);
  std::vector<float> fA(kRows * 64), fB(32 * 64);
  srand(7);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 255 - 127) / 64.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 255 - 127) / 64.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = kRows * 128 + 32 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int sbos[4] = {1024, 1280, 2048, 2304};
  printf("r0 sbo base_off max_abs_err verdict\n");
  for (int si = 0; si < 4; ++si)
    for (int r0 = 0; r0 <= 10; ++r0)
      for (int mode = 0; mode < 2; ++mode) {
        const int sbo = sbos[si], base_off = mode ? (r0 & 7) : 0;
        if (mode == 1 && (r0 & 7) == 0) continue;
        if (r0 + 15 * (sbo / 128) + 8 > kRows) continue;
        cudaMemset(dD, 0, 128 * 32 * 4);
        probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, r0, sbo, base_off, dD);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%d %d %d launch failed: %s\n", r0, sbo, base_off, cudaGetErrorString(e)); return 1; }
        std::vector<float> hD(128 * 32);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int R = r0 + (m / 8) * (sbo / 128) + m % 8;
          for (int n = 0; n < 32; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += static_cast<double>(fA[R * 64 + k]) * fB[n * 64 + k];
            const double d = fabs(ref - hD[m * 32 + n]);
            if (d > maxerr) maxerr = d;
          }
        }
        printf("%d %d %d %.4g %s\n", r0, sbo, base_off, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
      }
  return 0;
}

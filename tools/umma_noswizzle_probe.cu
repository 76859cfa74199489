// Hardware probe (developer tool, not part of the library): does a tcgen05.mma K-major A descriptor WITHOUT swizzle
// accept overlapping core matrices — LBO = 16 B (the next 8-element K chunk starts one 16-byte slot later) and
// SBO = 128 B (the next 8 rows start 8 slots later)?  With one 16-byte slot per pixel this makes the A row of pixel m
// the window slot[m], slot[m+1], ... : a horizontal convolution window read straight from a packed image row, no
// im2col copy.  The stem convolution (3 -> 64, 7 x 7) is built on the answer.
//
// A: kSlots x 8 bf16, linear.  For every (first slot p0, variant) the kernel runs M=128 N=32 K=64 (4 MMAs, MMA j
// starts at slot p0 + 2j) and the host compares with D[m][n] = sum_k slot[p0 + m + k/8][k%8] * B[n][k].
//   variant 0: LBO = 16, SBO = 128      variant 1: LBO = 128, SBO = 16 (roles swapped, expected WRONG)
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_noswizzle_probe tools/umma_noswizzle_probe.cu
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../deepliif_b200/csrc/ptx.cuh"

using namespace dlb;

constexpr int kSlots = 160;

__device__ __forceinline__ uint64_t make_desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;                                  // layout type 0: no swizzle
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       int p0, int variant, float* __restrict__ D) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                       // 32 x 128 B, SW128
  uint8_t* sA = smem + 32 * 128;            // kSlots x 16 B, linear
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < kSlots; i += 128)
    *reinterpret_cast<uint4*>(sA + i * 16) = *reinterpret_cast<const uint4*>(A + i * 8);
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
    const uint32_t a0 = smem_u32(sA) + p0 * 16, b0 = smem_u32(sB);
    const uint32_t lbo = variant == 0 ? 16 : 128, sbo = variant == 0 ? 128 : 16;
    for (int j = 0; j < 4; ++j)
      umma_f16(tmem, make_desc_none(a0 + j * 32, lbo, sbo), make_sw128_kmajor_desc(b0 + j * 32), idesc, j > 0);
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
  std::vector<__nv_bfloat16> hA(kSlots * 8), hB(32 * 64);
  std::vector<float> fA(kSlots * 8), fB(32 * 64);
  srand(11);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 255 - 127) / 64.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 255 - 127) / 64.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem_bytes = 32 * 128 + kSlots * 16 + 1024;
  printf("p0 variant max_abs_err verdict\n");
  for (int variant = 0; variant < 2; ++variant)
    for (int p0 = 0; p0 <= 9; ++p0) {
      cudaMemset(dD, 0, 128 * 32 * 4);
      probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, p0, variant, dD);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%d %d launch failed: %s\n", p0, variant, cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(128 * 32);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += static_cast<double>(fA[(p0 + m + k / 8) * 8 + k % 8]) * fB[n * 64 + k];
          const double d = fabs(ref - hD[m * 32 + n]);
          if (d > maxerr) maxerr = d;
        }
      printf("%d %d %.4g %s\n", p0, variant, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}

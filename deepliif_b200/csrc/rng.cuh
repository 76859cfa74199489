// Counter-based dropout mask shared by the forward (norm.cu) and backward (norm_bwd.cu) passes: the keep decision of
// element `idx` is a pure function of (seed, idx), so nothing is stored between forward and backward.
// (nn.Dropout(0.5) in ResnetBlock / UnetSkipConnectionBlock, networks.py:493-494, 604-605.  The reference's masks come
// from ATen's Philox stream and cannot be reproduced bit-for-bit by any other implementation; tests share this mask
// with the oracle instead — tests/test_training_gpu.py::dropout_mask_numpy is the same function in numpy.)
#pragma once
#include <stdint.h>

namespace dlb {

__host__ __device__ __forceinline__ uint32_t dropout_hash(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = idx + seed * 0x9E3779B97F4A7C15ULL + 0x632BE59BD9B4E019ULL;   // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}

// Seed of one layer invocation: the host-drawn seed, advanced by an optional device-resident step counter so that a
// captured CUDA graph draws a fresh mask on every replay (the forward and the backward of one replay read the same value).
__device__ __forceinline__ unsigned long long effective_seed(unsigned long long seed, const unsigned long long* epoch) {
  return epoch ? seed + __ldg(epoch) * 0xD1B54A32D192ED03ULL : seed;
}

// keep with probability 1 - p; returns the multiplier (0 or 1/(1-p))
__host__ __device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx, float p) {
  const float u = static_cast<float>(dropout_hash(seed, idx) >> 8) * (1.0f / 16777216.0f);   // [0,1)
  return u >= p ? 1.0f / (1.0f - p) : 0.0f;
}

}  // namespace dlb

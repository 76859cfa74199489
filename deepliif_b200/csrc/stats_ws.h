// Layout of the normalisation-statistics workspace shared by the producers of partial statistics (the conv_tc
// epilogue and the standalone stats_partial kernel) and the finalize kernel.
//   header   int S            number of slices the last producer wrote per image
//   counters int[...]         "last block done" tickets of the finalize kernel (self-resetting; the workspace
//                             must be zero-initialised ONCE when it is allocated)
//   cnt      float [N][S_cap]           valid pixels per slice
//   partial  float2[N][S_cap][C]        (sum, M2 about the slice mean) per slice and channel
//   group    float4[N][G_cap][C]        (count, mean, M2, -) merged over groups of 64 slices
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace dlb {

constexpr int kSlicesPerGroup = 64;

struct StatsLayout {
  int S_cap, G_cap, cchunks;
  size_t off_counters, off_cnt, off_partial, off_group, total;
};

inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

inline StatsLayout stats_layout(int N, int HW, int C) {
  StatsLayout L;
  L.S_cap = 2 * ((HW + 31) / 32) + 256;
  L.G_cap = (L.S_cap + kSlicesPerGroup - 1) / kSlicesPerGroup;
  L.cchunks = (C + 31) / 32;
  L.off_counters = 256;
  L.off_cnt = align256(L.off_counters + sizeof(int) * static_cast<size_t>(N + 1) * L.cchunks);
  L.off_partial = align256(L.off_cnt + sizeof(float) * static_cast<size_t>(N) * L.S_cap);
  L.off_group = align256(L.off_partial + sizeof(float2) * static_cast<size_t>(N) * L.S_cap * C);
  L.total = align256(L.off_group + sizeof(float4) * static_cast<size_t>(N) * L.G_cap * C);
  return L;
}

struct StatsPtrs {
  int* S;
  int* counters;
  float* cnt;
  float2* partial;
  float4* group;
  int S_cap, G_cap, cchunks;
};

inline StatsPtrs stats_ptrs(void* ws, const StatsLayout& L) {
  char* b = static_cast<char*>(ws);
  StatsPtrs p;
  p.S = reinterpret_cast<int*>(b);
  p.counters = reinterpret_cast<int*>(b + L.off_counters);
  p.cnt = reinterpret_cast<float*>(b + L.off_cnt);
  p.partial = reinterpret_cast<float2*>(b + L.off_partial);
  p.group = reinterpret_cast<float4*>(b + L.off_group);
  p.S_cap = L.S_cap; p.G_cap = L.G_cap; p.cchunks = L.cchunks;
  return p;
}

}  // namespace dlb

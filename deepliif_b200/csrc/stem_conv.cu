// Row-streaming tensor-core convolution for the generator stem: ReflectionPad2d(3) / ZeroPad2d(3) + Conv2d(C <= 4, 64, 7)
// (+ bias) from the fp32 NCHW network input (reference networks.py:386-397), fp32 NHWC raw output + the partial
// normalisation statistics of the following norm layer (stats_ws.h), in one kernel and WITHOUT an im2col operand.
//
// Operand.  Every input pixel is one 16-byte slot of 8 bf16 lanes [hi(c0..c3), lo(c0..c3)] (hi = rn(x), lo = rn(x - hi)).
// A padded input row is a linear run of slots in shared memory.  tcgen05.mma reads a K-major A tile WITHOUT swizzle as
// 8-row x 16-byte core matrices at byte offsets  row*16 (inside a group of 8 rows), group*SBO, kchunk*LBO;  with
// SBO = 128 B and LBO = 16 B the element (pixel m, K chunk q) is slot[m + q]: the A row of output pixel m is the
// window slot[m], ..., slot[m+7] of the packed row, i.e. the horizontal taps kw = q of a 7-wide filter, read in place
// (overlapping core matrices; measured exact on B200 with tools/umma_noswizzle_probe.cu).  One K = 16 MMA covers two
// taps; a filter row kh is 4 MMAs (tap 7 has zero weights), the 7 filter rows read 7 different ring rows.
// Split precision in ONE pass: B = [w_hi, w_hi | w_lo, 0] over the lanes [hi, lo] as 128 output columns, so
//   D[:, 0:64]  = x_hi*w_hi + x_lo*w_hi      D[:, 64:128] = x_hi*w_lo      out = D[:, 0:64] + D[:, 64:128]  (epilogue).
// Per output row: 28 x tcgen05.mma M128 N128 K16.
//
// A CTA walks down a 128-pixel-wide column strip: 2 converter warps keep a 16-row ring of packed rows ahead of the MMA
// warp (3 x 134 floats per row: plain coalesced loads, zero / reflected border resolved here), the MMA warp issues a row
// as soon as the rows r-3..r+3 are packed, 2 x 4 epilogue warps (even / odd rows) add the halves (+ bias), hand the TMEM
// accumulator back, stage the 128 px x 64 ch fp32 tile in swizzled shared memory for two TMA box stores and write one
// statistics slice (count, sum, M2 about the slice mean per channel) per row tile.
// Work split: all N x strips x H strip rows form one sequence cut into gridDim.x contiguous ranges.
// Algorithmic bytes: 4*C read per input pixel + 256 written per output pixel.
#include <cuda_bf16.h>

#include "internal.h"
#include "ptx.cuh"
#include "stats_ws.h"

namespace dlb {
namespace {

constexpr int kScThreads = 384;                      // warp 0 weights, 1 MMA, 2-5 / 6-9 epilogue (even / odd rows), 10-11 converters
                                                     // (384 threads: ptxas then allows 168 registers, the epilogue keeps 64 sums live)
constexpr int kScTW = 128;
constexpr int kScS = 7, kScHalo = 3;
constexpr int kScSlots = 136;                        // slots per packed row: 128 + 7 window + 1
constexpr uint32_t kScRowBytes = kScSlots * 16;      // 2176
constexpr int kScNR = 16;                            // packed-row ring
constexpr int kScConvWarps = 2;
constexpr uint32_t kScWTile = 128 * 128;             // one kh weight tile: 128 rows (64 co x {main, w_lo}) x K = 64
constexpr uint32_t kScWBytes = kScS * kScWTile;      // 114688
constexpr uint32_t kScTileBytes = 2 * 128 * 128;      // one output row tile: 2 halves of 32 channels x 128 px x 128 B (SW128)
constexpr uint32_t kScOffTile = kScWBytes;           // [2 epilogue groups], 1024-byte aligned
constexpr uint32_t kScOffRing = kScOffTile + 2 * kScTileBytes;
constexpr uint32_t kScOffBar = kScOffRing + kScNR * kScRowBytes;
constexpr uint32_t kScOffStat = kScOffBar + 512;
constexpr uint32_t kScSmem = kScOffStat + 2 * (4 * 32 * 8 + 64) + 1024;

struct StemParams {
  CUtensorMap ymap;                // y as [N*H][W][64] fp32, box {32, 128, 1}, 128-byte swizzle (output tile stores)
  const float* x; const uint8_t* wpk; const float* bias; float* y;
  int N, C, H, W, border_mode, strips;
  long long rows_total;
  float2* st_partial; float* st_cnt; int* st_S; int st_S_cap;
};

struct Piece { int n, s, c0, ra, rb, ia, ib; };

__device__ __forceinline__ bool next_piece(const StemParams& p, long long& cur, long long end, Piece& pc) {
  if (cur >= end) return false;
  const long long col = cur / p.H;
  const int r0 = static_cast<int>(cur - col * p.H);
  const long long left = end - cur;
  const int len = left < (p.H - r0) ? static_cast<int>(left) : (p.H - r0);
  pc.n = static_cast<int>(col / p.strips);
  pc.s = static_cast<int>(col % p.strips);
  pc.c0 = pc.s * kScTW;
  pc.ra = r0; pc.rb = r0 + len;
  pc.ia = r0 - kScHalo < 0 ? 0 : r0 - kScHalo;
  pc.ib = pc.rb + kScHalo > p.H ? p.H : pc.rb + kScHalo;
  cur += len;
  return true;
}

// K-major, no swizzle: leading (K-chunk) byte offset 16, stride (8-row group) byte offset 128 -> overlapping windows
__device__ __forceinline__ uint64_t make_window_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(16 >> 4) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// Column sums of a 32 x 32 tile held one row per lane: see conv_tc.cu (same fixed butterfly order).
__device__ __forceinline__ float sc_colsum32(float (&a)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float recv = __shfl_xor_sync(0xffffffffu, send, s);
      a[i] = (up ? a[i + s] : a[i]) + recv;
    }
  }
  return a[0];
}

__global__ void __launch_bounds__(kScThreads, 1) stem_conv_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ uint8_t sc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;
  uint8_t* sR = smem + kScOffRing;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kScOffBar);
  uint64_t* wbar = bars;                 // weights landed
  uint64_t* aready = bars + 1;           // [kScNR] packed row written (one converter warp)
  uint64_t* afree = bars + 1 + kScNR;    // [kScNR] the MMAs that read the row retired
  uint64_t* tfull = bars + 1 + 2 * kScNR;  // [2]
  uint64_t* tempty = tfull + 2;            // [2] (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float2* st_x_all = reinterpret_cast<float2*>(smem + kScOffStat);  // [2 groups][4][32]
  float* st_n_all = reinterpret_cast<float*>(smem + kScOffStat + 2 * 4 * 32 * 8);   // [2][16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = p.rows_total;
  const long long beg = T * blockIdx.x / gridDim.x, end = T * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < kScNR; ++i) { mbar_init(&aready[i], 1); mbar_init(&afree[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
    if (blockIdx.x == 0 && p.st_S != nullptr) *p.st_S = p.H * p.strips;
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, kScWBytes);
      bulk_copy_g2s(sW, p.wpk, kScWBytes, wbar);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-convergent, one elected lane issues) =====================
    mbar_wait(wbar, 0);
    const uint32_t idesc = make_idesc_f16(128, 128, 1);
    const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(sW));
    const uint64_t da0 = make_window_desc(smem_u32(sR));
    const bool refl = p.border_mode == DLB_PAD_REFLECT;
    uint32_t seq0 = 0, ot = 0;                      // packed rows before this piece; output row-tile counter
    long long cur = beg; Piece pc;
    while (next_piece(p, cur, end, pc)) {
      int waited = pc.ia;                           // rows [ia, waited) are known to be packed
      for (int r = pc.ra; r < pc.rb; ++r, ++ot) {
        const int need = r + kScHalo + 1 < pc.ib ? r + kScHalo + 1 : pc.ib;
        for (; waited < need; ++waited) {
          const uint32_t sq = seq0 + static_cast<uint32_t>(waited - pc.ia);
          mbar_wait(&aready[sq % kScNR], (sq / kScNR) & 1u);
        }
        const uint32_t t = ot & 1u;
        mbar_wait(&tempty[t], ((ot >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + t * 128u;
        uint32_t accumulate = 0;
#pragma unroll 1
        for (int kh = 0; kh < kScS; ++kh) {
          int ri = r + kh - kScHalo;
          if (ri < 0) { if (!refl) continue; ri = -ri; }
          if (ri >= p.H) { if (!refl) continue; ri = 2 * p.H - 2 - ri; }
          const uint32_t sq = seq0 + static_cast<uint32_t>(ri - pc.ia);
          const uint64_t da = da0 + (((sq % kScNR) * kScRowBytes) >> 4);
          const uint64_t db = db0 + ((kh * kScWTile) >> 4);
          umma_f16_k64_elect(d_tmem, da, db, idesc, accumulate);   // the four K = 16 steps (two taps each) from one asm block
          accumulate = 1;
        }
        umma_commit_elect(&tfull[t]);
        // rows above r - 3 are not read again: release row r - 3 (all remaining rows after the last output row)
        if (r - kScHalo >= pc.ia) umma_commit_elect(&afree[(seq0 + static_cast<uint32_t>(r - kScHalo - pc.ia)) % kScNR]);
        if (r == pc.rb - 1) {
          int f = pc.rb - kScHalo; if (f < pc.ia) f = pc.ia;
          for (; f < pc.ib; ++f) umma_commit_elect(&afree[(seq0 + static_cast<uint32_t>(f - pc.ia)) % kScNR]);
        }
      }
      seq0 += static_cast<uint32_t>(pc.ib - pc.ia);
    }
  } else if (warp < 10) {
    // ===================== epilogue: halves added, + bias, fp32 NHWC store, statistics slice =====================
    // two groups of four warps: group g drains accumulator g (even / odd output rows), so a row's statistics butterflies
    // overlap the next row's
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;
    const int px = q * 32 + lane;
    float2* st_x = st_x_all + grp * 128;
    float* st_n = st_n_all + grp * 16;
    const bool leader = q == 0 && lane == 0;
    uint8_t* const tile_ptr = smem + kScOffTile + grp * kScTileBytes;
    const uint32_t tile = smem_u32(tile_ptr);
    auto group_sync = [&]() {
      if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
    };
    uint32_t ot = 0;
    long long cur = beg; Piece pc;
    while (next_piece(p, cur, end, pc)) {
      const int col = pc.c0 + px;
      const bool valid = col < p.W;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const float cntf = static_cast<float>(__popc(vmask));
      for (int r = pc.ra; r < pc.rb; ++r, ++ot) {
        const uint32_t t = ot & 1u;
        if (t != static_cast<uint32_t>(grp)) continue;
        mbar_wait_sleep(&tfull[t], (ot >> 1) & 1u);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 128u;
        const long long st_row = static_cast<long long>(pc.n) * p.st_S_cap + static_cast<long long>(r) * p.strips + pc.s;
        // both 32-channel groups are summed into registers first and the accumulator is handed back before the stores and
        // the statistics butterflies, so the MMAs of row r + 2 start ~2k cycles earlier
        float a0[32], a1[32];
        {
          uint32_t v[32], v2[32];
          tmem_ld_32x32(taddr, v);
          tmem_ld_32x32(taddr + 64, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) a0[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
          tmem_ld_32x32(taddr + 32, v);
          tmem_ld_32x32(taddr + 96, v2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) a1[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[t]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + j));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + 32 + j));
            a0[j] += b0.x; a0[j + 1] += b0.y; a0[j + 2] += b0.z; a0[j + 3] += b0.w;
            a1[j] += b1.x; a1[j + 1] += b1.y; a1[j + 2] += b1.z; a1[j + 3] += b1.w;
          }
        }
        // output tile -> shared memory in the layout of a 128-byte-swizzled TMA box (row = pixel, 16-byte chunk j of the row at
        // j ^ (px & 7): four wavefronts per warp store), then ONE elected thread stores both halves with TMA: 128-byte lines
        // instead of 2048 scattered 16-byte sectors per row tile through the LSU (measured: the LSU was the stem's bound)
        if (leader) bulk_wait_group_read0();                  // the previous tile of this group has left shared memory
        group_sync();
        {
          const uint32_t rowa = tile + static_cast<uint32_t>(px) * 128u, sw = static_cast<uint32_t>(px & 7);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t off = ((static_cast<uint32_t>(j) ^ sw) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + off), "f"(a0[4 * j]), "f"(a0[4 * j + 1]),
                         "f"(a0[4 * j + 2]), "f"(a0[4 * j + 3]) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + 16384u + off), "f"(a1[4 * j]), "f"(a1[4 * j + 1]),
                         "f"(a1[4 * j + 2]), "f"(a1[4 * j + 3]) : "memory");
          }
        }
        fence_proxy_async();
        auto stats = [&](float (&a)[32], const int c, const bool first) {
          float w[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = valid ? a[j] : 0.f;
          const float sum = sc_colsum32(w, lane);
          const float mean = cntf > 0.f ? sum / cntf : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = a[j] - __shfl_sync(0xffffffffu, mean, j);
            w[j] = valid ? d * d : 0.f;
          }
          const float m2 = sc_colsum32(w, lane);
          st_x[q * 32 + lane] = make_float2(sum, m2);
          if (lane == 0) st_n[q] = cntf;
          group_sync();
          if (first && leader) {                               // every thread's tile writes are fenced and behind the barrier
            tma_store_3d(&p.ymap, reinterpret_cast<const void*>(tile_ptr), 0, pc.c0, pc.n * p.H + r);
            tma_store_3d(&p.ymap, reinterpret_cast<const void*>(tile_ptr + 16384), 32, pc.c0, pc.n * p.H + r);
            bulk_commit_group();
          }
          if (q == 0) {
            float nt = 0.f, st = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { nt += st_n[k]; st += st_x[k * 32 + lane].x; }
            const float mt = nt > 0.f ? st / nt : 0.f;
            float m2t = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (st_n[k] > 0.f) { const float d = st_x[k * 32 + lane].x / st_n[k] - mt; m2t += st_x[k * 32 + lane].y + st_n[k] * d * d; }
            if (p.st_partial != nullptr) {
              p.st_partial[st_row * 64 + c + lane] = make_float2(st, m2t);
              if (c == 0 && lane == 0) p.st_cnt[st_row] = nt;
            }
          }
          group_sync();
        };
        stats(a0, 0, true);
        stats(a1, 32, false);
      }
    }
    if (leader) bulk_wait_group0();                           // the last tiles are in global memory before the CTA exits
  } else {
    // ===================== converters: warp cw packs the rows seq = cw (mod 2) =====================
    const int cw = warp - 10;
    const bool refl = p.border_mode == DLB_PAD_REFLECT;
    const long long plane = static_cast<long long>(p.H) * p.W;
    uint32_t seq0 = 0;
    long long cur = beg; Piece pc;
    while (next_piece(p, cur, end, pc)) {
      const float* xn = p.x + static_cast<long long>(pc.n) * p.C * plane;
      const int cbase = pc.c0 - kScHalo;
      // this lane's slots j = lane + 32*k: source column (or -1 = zero), fixed for the whole piece
      int src[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int j = lane + 32 * k;
        int cj = cbase + j;
        int cr = -1;
        if (j < kScTW + 2 * kScHalo) {
          if (cj >= 0 && cj < p.W) cr = cj;
          else if (refl && cj >= -kScHalo && cj < p.W + kScHalo) cr = cj < 0 ? -cj : 2 * p.W - 2 - cj;
        }
        src[k] = cr;
      }
      const int rows = pc.ib - pc.ia;
      for (int k0 = 0; k0 < rows; ++k0) {
        const uint32_t sq = seq0 + static_cast<uint32_t>(k0);
        if ((sq & (kScConvWarps - 1)) != static_cast<uint32_t>(cw)) continue;
        const int ri = pc.ia + k0;
        const float* xr = xn + static_cast<long long>(ri) * p.W;
        float v[5][4];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            v[k][ch] = (src[k] >= 0 && ch < p.C && lane + 32 * k < kScSlots) ? __ldg(xr + ch * plane + src[k]) : 0.f;
        }
        const uint32_t slot = sq % kScNR;
        mbar_wait_sleep(&afree[slot], ((sq / kScNR) & 1u) ^ 1u);     // far ahead of the MMAs most of the time: sleep, not poll
        const uint32_t rowa = smem_u32(sR + slot * kScRowBytes);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const int j = lane + 32 * k;
          if (j < kScSlots) {
            __nv_bfloat16 h[8];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              h[ch] = __float2bfloat16_rn(v[k][ch]);
              h[4 + ch] = __float2bfloat16_rn(v[k][ch] - __bfloat162float(h[ch]));
            }
            const uint4 pk = *reinterpret_cast<const uint4*>(h);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + j * 16), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w)
                         : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&aready[slot]);
      }
      seq0 += static_cast<uint32_t>(rows);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// w fp32 [64][C][7][7] -> [kh][row (128)][k = kw*8 + lane (64)] bf16, SW128 rows of 128 B; row = co (0..63): lanes
// [w_hi(c0..c3), w_hi(c0..c3)] (x_hi*w_hi + x_lo*w_hi), row = 64 + co: lanes [w_lo(c0..c3), 0 x 4] (x_hi*w_lo); kw = 7 zero.
__global__ void stem_pack_kernel(const float* __restrict__ w, int C, __nv_bfloat16* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kScS * 128 * 64) return;
  const int k = e & 63, row = (e >> 6) & 127, kh = e >> 13;
  const int kw = k >> 3, ln = k & 7, ch = ln & 3;
  const int co = row & 63, half = row >> 6;
  float v = 0.f;
  if (kw < kScS && ch < C) v = w[((co * C + ch) * kScS + kh) * kScS + kw];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16 val;
  if (half == 0) val = hi;
  else val = ln < 4 ? __float2bfloat16_rn(v - __bfloat162float(hi)) : __float2bfloat16_rn(0.f);
  const int chunk = (k >> 3) ^ (row & 7);
  out[(static_cast<size_t>(kh) * 128 + row) * 64 + chunk * 8 + (k & 7)] = val;
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" size_t dlb_stem_conv_weight_bytes(void) { return kScWBytes; }

extern "C" int dlb_stem_conv_pack_weights(const float* w, int Cout, int C, int R, int S, void* out, dlb_stream_t stream) {
  if (Cout != 64 || C < 1 || C > 4 || R != kScS || S != kScS) return set_error("dlb_stem_conv_pack_weights: needs a [64][C<=4][7][7] filter");
  stem_pack_kernel<<<(kScS * 128 * 64 + 255) / 256, 256, 0, stream>>>(w, C, reinterpret_cast<__nv_bfloat16*>(out));
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stem_pack_kernel launch");
  return 0;
}

extern "C" int dlb_stem_conv_fwd(const float* x_nchw, int N, int C, int H, int W, const void* w_packed, const float* bias,
                                 int Cout, int border_mode, float* y, void* stats_ws, size_t stats_ws_bytes, dlb_stream_t stream) {
  if (Cout != 64 || C < 1 || C > 4) return set_error("dlb_stem_conv_fwd: needs C <= 4 and Cout == 64");
  if (H < 8 || W < 8 || N < 1) return set_error("dlb_stem_conv_fwd: needs H, W >= 8");
  if (border_mode != DLB_PAD_ZERO && border_mode != DLB_PAD_REFLECT) return set_error("dlb_stem_conv_fwd: bad border mode");
  if ((reinterpret_cast<uintptr_t>(w_packed) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return set_error("dlb_stem_conv_fwd: w_packed and y must be 16-byte aligned");
  int sms = 0;
  if (int rc = device_num_sms(&sms)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(stem_conv_kernel), static_cast<int>(kScSmem), kSlotStemConv)) return rc;
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.x = x_nchw; p.wpk = static_cast<const uint8_t*>(w_packed); p.bias = bias; p.y = y;
  p.N = N; p.C = C; p.H = H; p.W = W; p.border_mode = border_mode;
  p.strips = (W + kScTW - 1) / kScTW;
  p.rows_total = static_cast<long long>(N) * p.strips * H;
  {
    const uint64_t dims[3] = {64, static_cast<uint64_t>(W), static_cast<uint64_t>(N) * H};
    const uint64_t strides[2] = {256, static_cast<uint64_t>(W) * 256};
    const uint32_t box[3] = {32, 128, 1};
    if (!encode_f32_map(&p.ymap, y, 3, dims, strides, box, 1)) return DLB_ERR_INVALID;
  }
  if (stats_ws != nullptr) {
    const StatsLayout L = stats_layout(N, H * W, Cout);
    if (stats_ws_bytes < L.total) return set_error("dlb_stem_conv_fwd: statistics workspace too small");
    if (H * p.strips > L.S_cap) return set_error("dlb_stem_conv_fwd: statistics workspace slice capacity exceeded");
    const StatsPtrs sp = stats_ptrs(stats_ws, L);
    p.st_partial = sp.partial; p.st_cnt = sp.cnt; p.st_S = sp.S; p.st_S_cap = sp.S_cap;
  }
  long long grid = p.rows_total / 4;
  if (grid < 1) grid = 1;
  if (grid > sms) grid = sms;
  stem_conv_kernel<<<static_cast<unsigned>(grid), kScThreads, kScSmem, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stem_conv_kernel launch");
  return 0;
}

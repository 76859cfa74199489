// Weight gradient of Conv2d / ConvTranspose2d on the tcgen05 tensor cores (training path; autograd of the layers
// conv_tc.cu runs forward — reference: loss.backward() through networks.py:399-404, 425-430, 490, 505, 638-659).
//
// Per filter tap t the gradient is a GEMM whose reduction runs over pixels:
//     G_t[m][n] = sum_px P[px][m] * Q[px*st + tap_t][n]
// P = "anchor" tensor (dy for Conv2d, x for ConvTranspose2d), Q = the other one, both NHWC 16-bit planes (hi [, lo]).
// Both operands are therefore *MN-major* for the MMA (channels contiguous, pixels = K strided): a TMA box of
// [64 channels x 64 pixels] lands as 64 K-rows of 128 B, which is exactly the canonical SW128 MN-major layout
// (8-row atoms 1024 B apart = SBO; the next 64-channel block 8192 B further = LBO).  No transposes are materialised.
// Q is read through the same shifted / stride-2 5-D views as the forward A operand (zero padding = TMA OOB fill).
//
// Work decomposition: one CTA = (tap, 128-row M tile, N tile, K split); K = all anchor pixels in 64-pixel boxes.
// Split-K partial tiles are stored in fp32 to a workspace and reduced in fixed order by wgrad_reduce_kernel, which
// also writes the PyTorch weight layout (deterministic, no atomics).  x3 precision as in the forward kernel.
#include "internal.h"
#include "ptx.cuh"

namespace dlb {
namespace {

constexpr int kMaxTaps = 64;
constexpr int kPx = 64;                 // pixels (K) per pipeline stage
constexpr int kBlkBytes = kPx * 128;    // one [64 ch x 64 px] block of one plane: 8 KB
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 232448;

struct alignas(64) WgParams {
  CUtensorMap p_hi, p_lo, q_hi, q_lo;
  int ntaps, planes, n_tile, nb;        // nb = n_tile / 64 Q blocks
  int q_dim_sel[5];                     // Q-map dim i takes: 0 channel block, 1 w0, 2 h0, 3 n0, 4 nothing
  int tap_off[kMaxTaps][5];
  int tile_w, tile_h, tile_n;           // anchor pixel box (tile_w*tile_h*tile_n = 64)
  int tiles_w, tiles_h, tiles_n;
  int m_tiles, n_tiles, splits;
  int CP, CQ;                           // channels of P (GEMM M) and Q (GEMM N)
  float* ws;                            // [splits][ntaps][CP][CQ]
  uint32_t idesc;
  int stages;
};

// MN-major SW128 descriptor: LBO = 8192 B (next 64-channel block), SBO = 1024 B (next 8 K-rows)
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes = kBlkBytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_kernel(const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  int item = blockIdx.x;
  const int split = item % p.splits; item /= p.splits;
  const int nt = item % p.n_tiles; item /= p.n_tiles;
  const int mt = item % p.m_tiles; item /= p.m_tiles;
  const int tap = item;
  const int total_pt = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (total_pt + p.splits - 1) / p.splits;
  const int pt0 = split * per, pt1 = min(pt0 + per, total_pt);
  const int p_bytes = 2 * kBlkBytes;                 // M = 128 = two 64-channel blocks
  const int q_bytes = p.nb * kBlkBytes;
  const int stage_bytes = p.planes * (p_bytes + q_bytes);
  uint32_t tmem_cols = 32; while (tmem_cols < static_cast<uint32_t>(p.n_tile)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.p_hi); prefetch_tensormap(&p.q_hi);
    if (p.planes == 2) { prefetch_tensormap(&p.p_lo); prefetch_tensormap(&p.q_lo); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_smem, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0; uint32_t ph = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        int t = pt;
        const int w0 = (t % p.tiles_w) * p.tile_w; t /= p.tiles_w;
        const int h0 = (t % p.tiles_h) * p.tile_h; t /= p.tiles_h;
        const int n0 = t * p.tile_n;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(stage_bytes));
        uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
        // P: anchor tensor, dims (c, w, h, n, 1), two 64-channel blocks of the M tile
        for (int pl = 0; pl < p.planes; ++pl) {
          const CUtensorMap* pm = pl == 0 ? &p.p_hi : &p.p_lo;
          for (int b = 0; b < 2; ++b)
            tma_load_5d(st + pl * p_bytes + b * kBlkBytes, pm, &full_bar[s], mt * 128 + b * 64, w0, h0, n0, 0);
        }
        // Q: shifted / strided view, nb 64-channel blocks of the N tile
        uint8_t* sq = st + p.planes * p_bytes;
        for (int pl = 0; pl < p.planes; ++pl) {
          const CUtensorMap* qm = pl == 0 ? &p.q_hi : &p.q_lo;
          for (int b = 0; b < p.nb; ++b) {
            int c[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int sel = p.q_dim_sel[i];
              c[i] = p.tap_off[tap][i] + (sel == 0 ? nt * p.n_tile + b * 64 : sel == 1 ? w0 : sel == 2 ? h0 : sel == 3 ? n0 : 0);
            }
            tma_load_5d(sq + pl * q_bytes + b * kBlkBytes, qm, &full_bar[s], c[0], c[1], c[2], c[3], c[4]);
          }
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int s = 0; uint32_t ph = 0;
      uint32_t accumulate = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t ph_ = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t pl_ = ph_ + p_bytes;
        const uint32_t qh_ = ph_ + p.planes * p_bytes;
        const uint32_t ql_ = qh_ + q_bytes;
#pragma unroll
        for (int k = 0; k < kPx / 16; ++k) {               // 16 pixels (K) per MMA = two 8-row atoms = 2048 B
          const uint64_t da_hi = make_sw128_mnmajor_desc(ph_ + k * 2048);
          const uint64_t db_hi = make_sw128_mnmajor_desc(qh_ + k * 2048);
          if (p.planes == 2) {
            const uint64_t da_lo = make_sw128_mnmajor_desc(pl_ + k * 2048);
            const uint64_t db_lo = make_sw128_mnmajor_desc(ql_ + k * 2048);
            umma_f16(tmem_base, da_lo, db_hi, p.idesc, accumulate);
            umma_f16(tmem_base, da_hi, db_lo, p.idesc, 1);
            umma_f16(tmem_base, da_hi, db_hi, p.idesc, 1);
          } else {
            umma_f16(tmem_base, da_hi, db_hi, p.idesc, accumulate);
          }
          accumulate = 1;
        }
        umma_commit(&empty_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 partial tile =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // row of the M tile = P channel
    const int mrow = mt * 128 + m;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    float* out = p.ws + ((static_cast<long long>(split) * p.ntaps + tap) * p.CP + mrow) * p.CQ + nt * p.n_tile;
    const bool any = pt1 > pt0;
    for (int c = 0; c < p.n_tile; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      tmem_ld_wait();
      if (mrow < p.CP && (nt * p.n_tile + c) < p.CQ) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = any ? __uint_as_float(v[j + 0]) : 0.f; o.y = any ? __uint_as_float(v[j + 1]) : 0.f;
          o.z = any ? __uint_as_float(v[j + 2]) : 0.f; o.w = any ? __uint_as_float(v[j + 3]) : 0.f;
          *reinterpret_cast<float4*>(out + c + j) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}


// ---------------------------------------------------------------------------------------------------------------------
// Multi-tap variant (every map at least 8 x 8 pixels).  A CTA owns a *group* of filter taps that differ only by a whole
// number of rows of the other operand Q (stride 1: the R vertical taps of one horizontal offset; stride 2: the taps of
// one (horizontal offset, row parity)).  The anchor pixel tile is 8 wide x 8 tall, so in the MN-major smem image one
// 8-pixel row = one 1024-byte swizzle atom: the Q strip [8 wide x (8 + extra) rows] is loaded ONCE and tap r's operand
// is the same strip starting r atoms further down (descriptor start address + r * 1024 keeps the swizzle phase).
// One accumulator per tap lives side by side in TMEM (taps * n_tile <= 512 columns).  Against the one-tap kernel this
// loads P once per group instead of once per tap and Q ~(8 + extra) / (8 * taps) as often — the one-tap kernel is
// bound by that L2 -> shared-memory traffic, not by the MMAs.
constexpr int kMaxGroups = 16;
constexpr int kMaxGT = 8;               // taps per group

struct alignas(64) WgMtParams {
  CUtensorMap p_hi, p_lo, q_hi, q_lo;
  int planes, n_tile, nb, CP, CQ, ntaps;
  int ngroups, rows;                    // rows of the Q strip box (same for every group)
  int grp_ntaps[kMaxGroups];
  int grp_off[kMaxGroups][5];           // Q-map coordinate offsets of the strip
  int grp_tap_row[kMaxGroups][kMaxGT];  // strip row where the tap's window starts
  int grp_tap_id[kMaxGroups][kMaxGT];   // r * S + s
  int q_dim_sel[5];
  int tiles_w, tiles_h, tiles_n, m_tiles, n_tiles, splits;
  float* ws;
  uint32_t idesc;
  int stages;
};

__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_mt_kernel(const __grid_constant__ WgMtParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int item = blockIdx.x;
  const int split = item % p.splits; item /= p.splits;
  const int nt = item % p.n_tiles; item /= p.n_tiles;
  const int mt = item % p.m_tiles; item /= p.m_tiles;
  const int grp = item;
  const int gt = p.grp_ntaps[grp];
  const int total_pt = p.tiles_w * p.tiles_h * p.tiles_n;
  const int per = (total_pt + p.splits - 1) / p.splits;
  const int pt0 = split * per, pt1 = min(pt0 + per, total_pt);
  const int p_bytes = 2 * kBlkBytes;                       // M = 128: two 64-channel blocks of an 8 x 8 pixel tile
  const int strip_bytes = p.rows * 1024;                   // one 64-channel block of the Q strip
  const int q_bytes = p.nb * strip_bytes;
  const int stage_bytes = p.planes * (p_bytes + q_bytes);
  uint32_t tmem_cols = 32; while (tmem_cols < static_cast<uint32_t>(gt * p.n_tile)) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&p.p_hi); prefetch_tensormap(&p.q_hi);
    if (p.planes == 2) { prefetch_tensormap(&p.p_lo); prefetch_tensormap(&p.q_lo); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tmem_base_smem, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0; uint32_t ph = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        int t = pt;
        const int w0 = (t % p.tiles_w) * 8; t /= p.tiles_w;
        const int h0 = (t % p.tiles_h) * 8; t /= p.tiles_h;
        const int n0 = t;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(stage_bytes));
        uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
        for (int pl = 0; pl < p.planes; ++pl) {
          const CUtensorMap* pm = pl == 0 ? &p.p_hi : &p.p_lo;
          for (int b = 0; b < 2; ++b)
            tma_load_5d(st + pl * p_bytes + b * kBlkBytes, pm, &full_bar[s], mt * 128 + b * 64, w0, h0, n0, 0);
        }
        uint8_t* sq = st + p.planes * p_bytes;
        for (int pl = 0; pl < p.planes; ++pl) {
          const CUtensorMap* qm = pl == 0 ? &p.q_hi : &p.q_lo;
          for (int b = 0; b < p.nb; ++b) {
            int c[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int sel = p.q_dim_sel[i];
              c[i] = p.grp_off[grp][i] + (sel == 0 ? nt * p.n_tile + b * 64 : sel == 1 ? w0 : sel == 2 ? h0 : sel == 3 ? n0 : 0);
            }
            tma_load_5d(sq + pl * q_bytes + b * strip_bytes, qm, &full_bar[s], c[0], c[1], c[2], c[3], c[4]);
          }
        }
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int s = 0; uint32_t ph = 0;
      uint32_t accumulate = 0;
      for (int pt = pt0; pt < pt1; ++pt) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t ph_ = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t pl_ = ph_ + p_bytes;
        const uint32_t qh_ = ph_ + p.planes * p_bytes;
        const uint32_t ql_ = qh_ + q_bytes;
        for (int g = 0; g < gt; ++g) {
          const uint32_t roff = static_cast<uint32_t>(p.grp_tap_row[grp][g]) * 1024u;
          const uint32_t acc = tmem_base + static_cast<uint32_t>(g * p.n_tile);
#pragma unroll
          for (int k = 0; k < kPx / 16; ++k) {             // 16 pixels (K) per MMA = two 8-pixel rows = 2048 B
            const uint64_t da_hi = make_sw128_mnmajor_desc(ph_ + k * 2048);
            const uint64_t db_hi = make_sw128_mnmajor_desc(qh_ + roff + k * 2048, strip_bytes);
            if (p.planes == 2) {
              const uint64_t da_lo = make_sw128_mnmajor_desc(pl_ + k * 2048);
              const uint64_t db_lo = make_sw128_mnmajor_desc(ql_ + roff + k * 2048, strip_bytes);
              umma_f16(acc, da_lo, db_hi, p.idesc, accumulate | (k > 0));
              umma_f16(acc, da_hi, db_lo, p.idesc, 1);
              umma_f16(acc, da_hi, db_hi, p.idesc, 1);
            } else {
              umma_f16(acc, da_hi, db_hi, p.idesc, accumulate | (k > 0));
            }
          }
        }
        accumulate = 1;
        umma_commit(&empty_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      umma_commit(&done_bar);
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 partial tiles (one per tap) =====================
    const int q = warp & 3;
    const int mrow = mt * 128 + q * 32 + lane;
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const bool any = pt1 > pt0;
    for (int g = 0; g < gt; ++g) {
      const int tap = p.grp_tap_id[grp][g];
      float* out = p.ws + ((static_cast<long long>(split) * p.ntaps + tap) * p.CP + mrow) * p.CQ + nt * p.n_tile;
      for (int c = 0; c < p.n_tile; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * p.n_tile + c), v);
        tmem_ld_wait();
        if (mrow < p.CP && (nt * p.n_tile + c) < p.CQ) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = any ? __uint_as_float(v[j + 0]) : 0.f; o.y = any ? __uint_as_float(v[j + 1]) : 0.f;
            o.z = any ? __uint_as_float(v[j + 2]) : 0.f; o.w = any ? __uint_as_float(v[j + 3]) : 0.f;
            *reinterpret_cast<float4*>(out + c + j) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dW (PyTorch layout [CP][CQ][taps]) (+)= sum_splits ws[split][tap][m][n]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int taps, int CP, int CQ,
                                    float* __restrict__ dw, int accumulate) {
  const long long total = static_cast<long long>(taps) * CP * CQ;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx % CQ);
    const int m = static_cast<int>((idx / CQ) % CP);
    const int t = static_cast<int>(idx / (static_cast<long long>(CQ) * CP));
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[static_cast<long long>(s) * total + idx];
    const long long dst = (static_cast<long long>(m) * CQ + n) * taps + t;
    dw[dst] = accumulate ? dw[dst] + acc : acc;
  }
}

bool encode5(CUtensorMap* map, const void* base, int is_bf16, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box) {
  return encode_tiled_map(map, base, is_bf16, 5, dims, strides, box);
}

int pow2c(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct WgGeom {
  int PN, PH, PW, CP;      // anchor tensor
  int QH, QW, CQ;          // other tensor
  int stride, taps, R, S, pad;
  int n_tile, nb, m_tiles, n_tiles, splits, tile_w, tile_h, tile_n, tiles_w, tiles_h, tiles_n;
  // multi-tap kernel
  int mt;                  // 1 = use conv_wgrad_mt_kernel
  int ngroups, rows, max_gt;
  int grp_ntaps[kMaxGroups], grp_min[kMaxGroups], grp_s[kMaxGroups], grp_hp[kMaxGroups];
  int grp_tap_row[kMaxGroups][kMaxGT], grp_tap_id[kMaxGroups][kMaxGT];
};

// Fill two waves of 148 SMs without spilling into a third one.
int pick_splits(int items, int total_pt, int cap) {
  int splits = (2 * 148) / items;
  if (splits < 1) splits = 1;
  if (splits > total_pt) splits = total_pt;
  if (splits > cap) splits = cap;
  return splits;
}

int wg_geometry(const dlb_conv_desc* d, WgGeom* g) {
  int OH, OW;
  if (d->transposed) { OH = (d->H - 1) * d->stride - 2 * d->pad + d->R + d->output_padding; OW = (d->W - 1) * d->stride - 2 * d->pad + d->S + d->output_padding; }
  else { OH = (d->H + 2 * d->pad - d->R) / d->stride + 1; OW = (d->W + 2 * d->pad - d->S) / d->stride + 1; }
  g->stride = d->stride; g->R = d->R; g->S = d->S; g->pad = d->pad; g->taps = d->R * d->S; g->PN = d->N;
  if (!d->transposed) { g->PH = OH; g->PW = OW; g->CP = d->Cout; g->QH = d->H; g->QW = d->W; g->CQ = d->Cin[0]; }
  else { g->PH = d->H; g->PW = d->W; g->CP = d->Cin[0]; g->QH = OH; g->QW = OW; g->CQ = d->Cout; }
  if (g->taps > kMaxTaps) return set_error("wgrad: more than 64 taps");
  if (g->CP % 64 != 0 || g->CQ % 64 != 0) return set_error("wgrad: channel counts must be multiples of 64 (pad small sides)");
  if (g->stride == 2 && ((g->QH & 1) || (g->QW & 1))) return set_error("wgrad: stride-2 needs even extents");
  g->n_tile = g->CQ >= 256 ? 256 : g->CQ;
  if (g->n_tile != 64 && g->n_tile != 128 && g->n_tile != 256) { g->n_tile = 64; }
  g->nb = g->n_tile / 64;
  g->m_tiles = (g->CP + 127) / 128;
  g->n_tiles = (g->CQ + g->n_tile - 1) / g->n_tile;
  g->tile_w = g->PW >= kPx ? kPx : pow2c(g->PW);
  g->tile_h = kPx / g->tile_w; { const int hp = pow2c(g->PH); if (g->tile_h > hp) g->tile_h = hp; }
  g->tile_n = kPx / (g->tile_w * g->tile_h);
  g->tiles_w = (g->PW + g->tile_w - 1) / g->tile_w;
  g->tiles_h = (g->PH + g->tile_h - 1) / g->tile_h;
  g->tiles_n = (g->PN + g->tile_n - 1) / g->tile_n;
  g->mt = (g->PW >= 8 && g->PH >= 8) ? 1 : 0;
  if (g->mt) {
    // groups of taps that share one Q strip: same horizontal offset (and, for stride 2, the same row parity)
    g->ngroups = 0; g->rows = 8; g->max_gt = 1;
    for (int s = 0; s < g->S; ++s)
      for (int hp = 0; hp < g->stride; ++hp) {
        int cnt = 0, lo = 0, hi = 0;
        int gi = g->ngroups;
        for (int r = 0; r < g->R; ++r) {
          const int dh = r - g->pad;
          const int par = ((dh % g->stride) + g->stride) % g->stride;
          if (par != hp) continue;
          const int row = (dh - par) / g->stride;          // Q row (in strip units) = anchor row + row
          if (cnt == 0) { lo = hi = row; } else { lo = row < lo ? row : lo; hi = row > hi ? row : hi; }
          if (cnt >= kMaxGT || gi >= kMaxGroups) return set_error("wgrad: too many taps per group");
          g->grp_tap_row[gi][cnt] = row; g->grp_tap_id[gi][cnt] = r * g->S + s;
          ++cnt;
        }
        if (cnt == 0) continue;
        for (int k = 0; k < cnt; ++k) g->grp_tap_row[gi][k] -= lo;
        g->grp_ntaps[gi] = cnt; g->grp_min[gi] = lo; g->grp_s[gi] = s; g->grp_hp[gi] = hp;
        if (8 + hi - lo > g->rows) g->rows = 8 + hi - lo;
        if (cnt > g->max_gt) g->max_gt = cnt;
        ++g->ngroups;
      }
    int nt = 256;
    while (nt > 64 && (nt * g->max_gt > 512 || nt > g->CQ)) nt >>= 1;
    g->n_tile = nt; g->nb = nt / 64;
    g->n_tiles = (g->CQ + nt - 1) / nt;
    g->tile_w = 8; g->tile_h = 8; g->tile_n = 1;
    g->tiles_w = (g->PW + 7) / 8; g->tiles_h = (g->PH + 7) / 8; g->tiles_n = g->PN;
    g->splits = pick_splits(g->ngroups * g->m_tiles * g->n_tiles, g->tiles_w * g->tiles_h * g->tiles_n, 296);
    return 0;
  }
  g->splits = pick_splits(g->taps * g->m_tiles * g->n_tiles, g->tiles_w * g->tiles_h * g->tiles_n, 64);
  return 0;
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" size_t dlb_conv_wgrad_workspace(const dlb_conv_desc* d) {
  WgGeom g;
  if (wg_geometry(d, &g) != 0) return 0;
  return static_cast<size_t>(g.splits) * g.taps * g.CP * g.CQ * sizeof(float);
}

extern "C" int dlb_conv_wgrad(const dlb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi,
                              const void* dy_lo, float* dw, int accumulate, int fmt, int split, void* workspace,
                              size_t workspace_bytes, dlb_stream_t stream) {
  if (d->nsrc != 1) return set_error("dlb_conv_wgrad: single source only");
  if (fmt != DLB_FMT_BF16 && fmt != DLB_FMT_FP16) return set_error("dlb_conv_wgrad: bad fmt");
  WgGeom g;
  if (wg_geometry(d, &g) != 0) return DLB_ERR_INVALID;
  const size_t need = static_cast<size_t>(g.splits) * g.taps * g.CP * g.CQ * sizeof(float);
  if (workspace_bytes < need) return set_error("dlb_conv_wgrad: workspace too small");
  const int is_bf16 = fmt == DLB_FMT_BF16;
  if (g.mt) {
    WgMtParams q;
    memset(&q, 0, sizeof(q));
    q.planes = split ? 2 : 1; q.ntaps = g.taps; q.n_tile = g.n_tile; q.nb = g.nb; q.CP = g.CP; q.CQ = g.CQ;
    q.ngroups = g.ngroups; q.rows = g.rows;
    q.tiles_w = g.tiles_w; q.tiles_h = g.tiles_h; q.tiles_n = g.tiles_n;
    q.m_tiles = g.m_tiles; q.n_tiles = g.n_tiles; q.splits = g.splits;
    q.ws = reinterpret_cast<float*>(workspace);
    q.idesc = make_idesc_f16(128, g.n_tile, is_bf16) | (1u << 15) | (1u << 16);
    const void* P_hi = d->transposed ? x_hi : dy_hi; const void* P_lo = d->transposed ? x_lo : dy_lo;
    const void* Q_hi = d->transposed ? dy_hi : x_hi; const void* Q_lo = d->transposed ? dy_lo : x_lo;
    {
      const uint64_t C = g.CP, W = g.PW, H = g.PH, N = g.PN;
      uint64_t dims[5] = {C, W, H, N, 1};
      uint64_t str[4] = {C * 2, W * C * 2, H * W * C * 2, N * H * W * C * 2};
      uint32_t box[5] = {64, 8, 8, 1, 1};
      if (!encode5(&q.p_hi, P_hi, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
      if (split && !encode5(&q.p_lo, P_lo, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
    }
    {
      const uint64_t C = g.CQ, W = g.QW, H = g.QH, N = g.PN;
      uint64_t dims[5], str[4]; uint32_t box[5];
      if (g.stride == 1) {
        dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 1;
        str[0] = C * 2; str[1] = W * C * 2; str[2] = H * W * C * 2; str[3] = N * H * W * C * 2;
        box[0] = 64; box[1] = 8; box[2] = g.rows; box[3] = 1; box[4] = 1;
        q.q_dim_sel[0] = 0; q.q_dim_sel[1] = 1; q.q_dim_sel[2] = 2; q.q_dim_sel[3] = 3; q.q_dim_sel[4] = 4;
      } else {
        dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
        str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
        box[0] = 64; box[1] = 8; box[2] = 1; box[3] = g.rows; box[4] = 1;
        q.q_dim_sel[0] = 0; q.q_dim_sel[1] = 1; q.q_dim_sel[2] = 4; q.q_dim_sel[3] = 2; q.q_dim_sel[4] = 3;
      }
      if (!encode5(&q.q_hi, Q_hi, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
      if (split && !encode5(&q.q_lo, Q_lo, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
      for (int gi = 0; gi < g.ngroups; ++gi) {
        q.grp_ntaps[gi] = g.grp_ntaps[gi];
        for (int k = 0; k < g.grp_ntaps[gi]; ++k) { q.grp_tap_row[gi][k] = g.grp_tap_row[gi][k]; q.grp_tap_id[gi][k] = g.grp_tap_id[gi][k]; }
        const int dw = g.grp_s[gi] - g.pad;
        if (g.stride == 1) {
          q.grp_off[gi][0] = 0; q.grp_off[gi][1] = dw; q.grp_off[gi][2] = g.grp_min[gi]; q.grp_off[gi][3] = 0; q.grp_off[gi][4] = 0;
        } else {
          const int wp = ((dw % 2) + 2) % 2;
          q.grp_off[gi][0] = wp * static_cast<int>(C); q.grp_off[gi][1] = (dw - wp) / 2; q.grp_off[gi][2] = g.grp_hp[gi];
          q.grp_off[gi][3] = g.grp_min[gi]; q.grp_off[gi][4] = 0;
        }
      }
    }
    const int stage_bytes = q.planes * (2 * kBlkBytes + g.nb * g.rows * 1024);
    int stages = (kSmemLimit - 2048 - 1024) / stage_bytes;
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return set_error("dlb_conv_wgrad: not enough shared memory (multi-tap)");
    q.stages = stages;
    if (ensure_dyn_smem(reinterpret_cast<const void*>(conv_wgrad_mt_kernel), kSmemLimit - 2048, kSlotWgradMt) != 0) return DLB_ERR_CUDA;
    const int grid = g.ngroups * g.m_tiles * g.n_tiles * g.splits;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    conv_wgrad_mt_kernel<<<grid, kThreads, stages * stage_bytes + 1024, st>>>(q);
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("conv_wgrad_mt_kernel launch");
    const long long total = static_cast<long long>(g.taps) * g.CP * g.CQ;
    long long rg = (total + 255) / 256; if (rg > 148 * 8) rg = 148 * 8;
    wgrad_reduce_kernel<<<static_cast<int>(rg), 256, 0, st>>>(q.ws, g.splits, g.taps, g.CP, g.CQ, dw, accumulate);
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("wgrad_reduce_kernel launch");
    return 0;
  }
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.planes = split ? 2 : 1; p.ntaps = g.taps; p.n_tile = g.n_tile; p.nb = g.nb;
  p.tile_w = g.tile_w; p.tile_h = g.tile_h; p.tile_n = g.tile_n;
  p.tiles_w = g.tiles_w; p.tiles_h = g.tiles_h; p.tiles_n = g.tiles_n;
  p.m_tiles = g.m_tiles; p.n_tiles = g.n_tiles; p.splits = g.splits; p.CP = g.CP; p.CQ = g.CQ;
  p.ws = reinterpret_cast<float*>(workspace);
  // instruction descriptor: both operands MN-major (bits 15, 16)
  p.idesc = make_idesc_f16(128, g.n_tile, is_bf16) | (1u << 15) | (1u << 16);
  const void* P_hi = d->transposed ? x_hi : dy_hi; const void* P_lo = d->transposed ? x_lo : dy_lo;
  const void* Q_hi = d->transposed ? dy_hi : x_hi; const void* Q_lo = d->transposed ? dy_lo : x_lo;
  {  // anchor map: (c, w, h, n, 1)
    const uint64_t C = g.CP, W = g.PW, H = g.PH, N = g.PN;
    uint64_t dims[5] = {C, W, H, N, 1};
    uint64_t str[4] = {C * 2, W * C * 2, H * W * C * 2, N * H * W * C * 2};
    uint32_t box[5] = {64, (uint32_t)g.tile_w, (uint32_t)g.tile_h, (uint32_t)g.tile_n, 1};
    if (!encode5(&p.p_hi, P_hi, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
    if (split && !encode5(&p.p_lo, P_lo, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
  }
  {  // Q map: stride 1 -> (c, w, h, n, 1); stride 2 -> (wp*C + c, ww, hp, hh, n)
    const uint64_t C = g.CQ, W = g.QW, H = g.QH, N = g.PN;
    uint64_t dims[5], str[4]; uint32_t box[5];
    if (g.stride == 1) {
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 1;
      str[0] = C * 2; str[1] = W * C * 2; str[2] = H * W * C * 2; str[3] = N * H * W * C * 2;
      box[0] = 64; box[1] = g.tile_w; box[2] = g.tile_h; box[3] = g.tile_n; box[4] = 1;
      p.q_dim_sel[0] = 0; p.q_dim_sel[1] = 1; p.q_dim_sel[2] = 2; p.q_dim_sel[3] = 3; p.q_dim_sel[4] = 4;
    } else {
      dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
      str[0] = 2 * C * 2; str[1] = W * C * 2; str[2] = 2 * W * C * 2; str[3] = H * W * C * 2;
      box[0] = 64; box[1] = g.tile_w; box[2] = 1; box[3] = g.tile_h; box[4] = g.tile_n;
      p.q_dim_sel[0] = 0; p.q_dim_sel[1] = 1; p.q_dim_sel[2] = 4; p.q_dim_sel[3] = 2; p.q_dim_sel[4] = 3;
    }
    if (!encode5(&p.q_hi, Q_hi, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
    if (split && !encode5(&p.q_lo, Q_lo, is_bf16, dims, str, box)) return DLB_ERR_INVALID;
    for (int r = 0; r < g.R; ++r)
      for (int s = 0; s < g.S; ++s) {
        const int t = r * g.S + s;
        const int dh = r - g.pad, dw = s - g.pad;      // Q index = anchor*stride + (r - pad)
        if (g.stride == 1) {
          p.tap_off[t][0] = 0; p.tap_off[t][1] = dw; p.tap_off[t][2] = dh; p.tap_off[t][3] = 0; p.tap_off[t][4] = 0;
        } else {
          const int hp = ((dh % 2) + 2) % 2, wp = ((dw % 2) + 2) % 2;
          p.tap_off[t][0] = wp * static_cast<int>(C); p.tap_off[t][1] = (dw - wp) / 2; p.tap_off[t][2] = hp;
          p.tap_off[t][3] = (dh - hp) / 2; p.tap_off[t][4] = 0;
        }
      }
  }
  const int stage_bytes = p.planes * (2 * kBlkBytes + g.nb * kBlkBytes);
  int stages = (kSmemLimit - 2048 - 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error("dlb_conv_wgrad: not enough shared memory");
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(conv_wgrad_kernel), kSmemLimit - 2048, kSlotWgrad) != 0) return DLB_ERR_CUDA;
  const int grid = g.taps * g.m_tiles * g.n_tiles * g.splits;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  conv_wgrad_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("conv_wgrad_kernel launch");
  const long long total = static_cast<long long>(g.taps) * g.CP * g.CQ;
  long long rg = (total + 255) / 256; if (rg > 148 * 8) rg = 148 * 8;
  wgrad_reduce_kernel<<<static_cast<int>(rg), 256, 0, st>>>(p.ws, g.splits, g.taps, g.CP, g.CQ, dw, accumulate);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("wgrad_reduce_kernel launch");
  return 0;
}

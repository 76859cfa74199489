// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05) for sm_100a.
//
// Replaces, for every layer whose input-channel count is a multiple of 64, the cuDNN call behind
// nn.Conv2d / nn.ConvTranspose2d in the reference generators / discriminators
// (/root/reference/deepliif/models/networks.py:399-404 down convs, :490/:505 ResNet-block convs,
//  :425-430 ConvTranspose upsampling, :576-600 UNet convs, :638-659 PatchGAN convs).
//
// Formulation.  One "phase" of a convolution is
//     y[n, i, j, co] = sum_t sum_ci  x[n, i*st + dh_t, j*st + dw_t, ci] * w[t][co][ci]
// over a tap list t (a stride-1/2 conv is one phase; a stride-2 ConvTranspose is four output-parity
// phases with 1/2/2/4 (3x3) or 4 (4x4) taps each, so no zero-insertion MACs are executed).
// GEMM view per CTA tile: M = 128 output pixels (a tile_n x tile_h x tile_w box), N = n_tile output
// channels, K = taps x Cin, walked in 64-channel chunks.
//
// Data path.  Activations live in HBM as NHWC 16-bit planes (hi [, lo]); weights as [tap][Cout][Cin]
// 16-bit planes (hi [, lo]).  One elected producer thread issues TMA tiled loads (5-D tensor map over
// the activation: the halo / zero padding is the TMA out-of-bounds fill, the stride-2 case is a
// (2C, W/2, 2, H/2, N) view of the same memory) into 128B-swizzled shared-memory stages guarded by
// full/empty mbarriers.  The MMA warp runs convergent and one elected lane issues tcgen05.mma (M=128,
// N=n_tile, K=16, fp32 accumulate in TMEM), one asm block per 64-channel chunk so that the
// descriptors live in uniform registers (ptx.cuh).  In split precision ("x3") every K step issues
// three MMAs hi*hi + hi*lo + lo*hi, which restores ~fp32 products from 16-bit operands (SURVEY.md
// §8d: the 1e-3 parity gate cannot be met by single-pass TF32/BF16).  The accumulator is
// double-buffered in TMEM (2 x n_tile columns) so the epilogue of tile i overlaps the MMAs of tile
// i+1.  Four epilogue warps read TMEM with tcgen05.ld (one output pixel per thread), add the bias
// and write fp32 NHWC through a shared-memory transpose buffer (TcParams::ets: coalesced 128-byte
// lines, or TMA box stores on the fused-operand launches; direct stores where the buffer does not
// fit), taking the normalisation statistics of the tile on the way.
// The kernel is persistent: grid = min(#tiles, #SMs), static round-robin tile schedule (tiles have
// identical cost).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <type_traits>

#include "internal.h"
#include "ptx.cuh"

namespace dlb {

namespace {

constexpr int kMaxTaps = 16;
constexpr int kMaxStages = 8;
constexpr int kKC = 64;                // channels per pipeline stage: 64 x 2 B = one 128 B swizzle row
constexpr int kABytes = 128 * 128;     // one A plane of a stage: 128 pixels x 128 B
constexpr int kThreads = 192;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int kConvWarps = 8;          // fused-operand mode: warps 6-13 build the A planes from fp32 producer outputs
constexpr int kThreadsFa = kThreads + 32 * kConvWarps + 32;   // + warp 14: fp32 source staging producer (ht mode)
constexpr int kStemPatchBytes = 4 * 24 * 16 * 4;   // fused stem: fp32 input patch [4 ch][<= 24 rows][16 cols] per strip stage
constexpr int kHsMaxPx = 192;          // halo-strip mode: at most this many strip pixels (3x3: 18 x 10 = 180; 2x2 taps: 17 x 9)
constexpr int kSmemLimit = 232448;     // 227 KB per CTA (static + dynamic)
constexpr int kStaticSmem = 2560;      // barriers + the epilogue groups' statistics scratch (ptxas: 2352 B)
constexpr int kMaxDynSmem = kSmemLimit - kStaticSmem - 1024;

struct alignas(64) TcParams {
  CUtensorMap a_hi[2];
  CUtensorMap a_lo[2];
  CUtensorMap b_hi;
  CUtensorMap b_lo;
  int ntaps, nsrc, planes, n_tile;
  int kchunks[2];
  int src_koff[2];
  int dim_sel[5];                      // A-map dim i takes: 0 channel chunk, 1 w0, 2 h0, 3 n0, 4 nothing
  int tap_off[kMaxTaps][5];
  int tap_w[kMaxTaps];
  int tile_w, tile_h, tile_n;
  int tiles_w, tiles_h, tiles_n, tiles_c;
  int N, OH, OW, cout_total;
  long long ys_n, ys_h, ys_w, y_base;  // output addressing in elements
  float* y;
  const float* bias;
  uint32_t idesc;
  int stages;
  // fused normalisation statistics (nullable): per 32-row slice (sum, M2) of y (bias included)
  float2* st_partial;
  float* st_cnt;
  int* st_S;
  int st_S_cap, st_slice_base, st_S_total;
  // "vertical strip" mode (R x 1 filters with small weights: stem / head): the weights of all taps stay resident in
  // shared memory, one A strip of tile_h + span rows is loaded per tile and channel chunk, and every tap is an MMA on
  // a 1024-byte-aligned window of that strip (tile_w = 8 pixels = one 8-row swizzle atom per image row).
  int vs, vs_rows, vs_dh_min;
  int vs_row_off[kMaxTaps];
  // "fused operand" mode (fa): there are no operand planes in HBM.  Converter warps read the PRODUCER's raw fp32 output,
  // apply its normalisation (per-(n,c) scale/shift), activation and residual add, split the result into hi/lo 16-bit
  // values and write them straight into the 128B-swizzled A stage the MMA reads — the nn.BatchNorm2d/InstanceNorm2d +
  // nn.ReLU (+ ResnetBlock skip add) between two convolutions (networks.py:490-513) costs no HBM round trip.
  int fa, fa_is_bf16, fa_wb_tap;
  const float* fa_x[2];                // fp32 NHWC [N, Hs, Ws, cin[s]] per K-source
  const float* fa_scale[2];            // [N][cin] or null
  const float* fa_shift[2];
  const float* fa_res[2];              // optional fp32 NHWC added after the activation
  float* fa_out[2];                    // optional: the materialised operand, written once (by the tap that maps 1:1)
  int fa_act[2];
  int fa_cin[2];
  int fa_border, fa_border_mode;       // the conv's input = source behind a zero / reflected border of this width
  int H, W, Hs, Ws, conv_stride;       // conv input extents (incl. border), source extents
  int tap_dh[kMaxTaps], tap_dw[kMaxTaps];
  // "halo strip" mode (hs; fused-operand, stride 1): per 64-channel chunk the converters build ONE operand strip of
  // (tile_h + dh span) x (tile_w + dw span) input pixels — every input value is converted once per tile instead of once
  // per tap — and each tap's MMAs address their shifted 16 x 8 pixel window of it directly: descriptor start =
  // strip + ((dh - dh_min) * cols + (dw - dw_min)) * 128 B, group stride (SBO) = cols * 128 B.  Weights stream per
  // (chunk, tap) through the stage ring as in tap mode.
  int hs, hs_rows, hs_cols, hs_plane_bytes, hs_dh_min, hs_dw_min, hs_nbuf;
  int hs_off[kMaxTaps];
  // fa == 2 ("stem" source, vertical-strip mode only): fa_x[0] is the network input, fp32 NCHW [N, stem_C <= 4, Hs, Ws];
  // the converters build the horizontal-window operand lane (s * 8 + c) = pad(x)[n, c, row - pad, col + s - pad] of
  // dlb_stem_window_pack on the fly, so the 21x blown-up operand tensor never exists in HBM.
  int stem_C, stem_S, stem_pad;
  // vertical-strip fused mode with TMA staging (vt): the producer thread streams the fp32 source rows of every strip into
  // a two-slot shared-memory staging ring (half a strip per slot) and the converter warps transform smem -> smem; no
  // global-load latency sits inside the conversion.  Zero border only (TMA out-of-bounds fill = the zero padding).
  int vt, vt_half_rows;
  CUtensorMap a_f32;
  // epilogue through shared memory, behind the stage ring at ets_off.  ets == 1: one 16 KB transpose buffer, the epilogue warps
  // copy it out with coalesced stores.  ets == 2: two buffers, one thread stores each as a TMA box: y_map[a] = the fp32 output of
  // accumulator a as a tiled map {cout, OW, OH, N} (pixel strides = the phase's ys_w / ys_h), box {32, tile_w, tile_h, tile_n},
  // 128-byte swizzle — nothing of the store touches the LSU, which the fused-operand converters need for their loads.
  CUtensorMap y_map[4];
  int ets, ets_off;
  // epi2: plane-fed launches carry a second group of four epilogue warps (warps 6-9); the 32-channel units of a tile alternate
  // between the groups (own transpose buffer, statistics scratch and named barriers each)
  int epi2;
  // halo-strip mode with TMA staging (ht): a dedicated producer thread (warp 14) streams the fp32 strip of every chunk
  // (box 64 ch x cols x rows) into a two-slot staging ring; the converters transform smem -> smem.  This is what makes the
  // fused operand pay off for layers with little MMA work per strip (ConvTranspose phases).  Single plain source, zero border.
  int ht, ht_slot_bytes;
  // hp: halo strips loaded by TMA straight from the operand planes (unfused operands: the box (64 ch, cols, rows) of a
  // 128B-swizzled tensor map lands in exactly the strip layout) — each input value crosses L2 -> SM once per tile, not once
  // per tap.
  int hp;
  // cta2: CTA pair (cluster of 2, cta_group::2).  The two CTAs work on consecutive pixel tiles with the same output channels;
  // ONE M = 256 MMA (issued by the even CTA) multiplies both strips with a weight tile of which each CTA holds half the rows,
  // so every SM loads, stores in shared memory and feeds to the tensor core only half of B.
  int cta2;
  // Merged output-parity phases of a ConvTranspose2d: the phases share the input strip; tap t accumulates into TMEM
  // accumulator tap_acc[t] (nacc x n_tile columns per set) and the epilogue writes accumulator a at output offset acc_ybase[a]
  // with statistics slice base acc_slice[a].
  int nacc;
  int tap_acc[kMaxTaps];
  long long acc_ybase[4];
  int acc_slice[4];
};

struct TileCoord {
  int w0, h0, n0, cout0;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t) {
  TileCoord c;
  int tw = t % p.tiles_w; t /= p.tiles_w;
  int th = t % p.tiles_h; t /= p.tiles_h;
  int tn = t % p.tiles_n; t /= p.tiles_n;
  c.w0 = tw * p.tile_w; c.h0 = th * p.tile_h; c.n0 = tn * p.tile_n; c.cout0 = t * p.n_tile;
  return c;
}

__device__ __forceinline__ float fa_act1(float v, int act) {
  if (act == DLB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLB_ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
  return v;
}

// 4 fp32 values -> 4 hi + 4 lo 16-bit values (hi = rn16(x), lo = rn16(x - hi)): the arithmetic of norm_apply_kernel.
__device__ __forceinline__ void fa_split4(const float (&o)[4], int is_bf16, uint2& hi, uint2& lo) {
  if (is_bf16) {
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(o[k]); l[k] = __float2bfloat16_rn(o[k] - __bfloat162float(h[k])); }
    hi = *reinterpret_cast<uint2*>(h); lo = *reinterpret_cast<uint2*>(l);
  } else {
    __half h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { h[k] = __float2half_rn(o[k]); l[k] = __float2half_rn(o[k] - __half2float(h[k])); }
    hi = *reinterpret_cast<uint2*>(h); lo = *reinterpret_cast<uint2*>(l);
  }
}

// One operand element group: 4 channels of one input pixel of K-source `src`.  (vh, vw): coordinates in the conv's input
// (source + border).  Returns the transformed values (zeros outside the input = the conv's own zero padding, or in a zero
// border) and whether the position is an interior source pixel (eligible for the write-back).
struct FaPix { long long off; bool ok, interior; int n; };

__device__ __forceinline__ FaPix fa_locate(const TcParams& p, int src, int n, int vh, int vw) {
  FaPix r;
  r.n = n;
  r.ok = (n < p.N) && (vh >= 0) && (vh < p.H) && (vw >= 0) && (vw < p.W);
  int sh = vh - p.fa_border, sw = vw - p.fa_border;
  r.interior = r.ok && sh >= 0 && sh < p.Hs && sw >= 0 && sw < p.Ws;
  if (!r.interior) {
    if (p.fa_border_mode == DLB_PAD_REFLECT) {
      if (sh < 0) sh = -sh; if (sh >= p.Hs) sh = 2 * p.Hs - 2 - sh;
      if (sw < 0) sw = -sw; if (sw >= p.Ws) sw = 2 * p.Ws - 2 - sw;
    } else {
      r.ok = false;
    }
  }
  r.off = r.ok ? ((static_cast<long long>(n) * p.Hs + sh) * p.Ws + sw) * p.fa_cin[src] : 0;
  return r;
}

// Column sums of a 32 x 32 tile held one row per lane (a[j] = element (lane, j)): a butterfly that halves the values per
// lane at every step (16 + 8 + 4 + 2 + 1 = 31 shuffles); on return lane l holds the sum of column l.  Fixed tree order.
__device__ __forceinline__ float warp_colsum32(float (&a)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? a[i] : a[i + s];
      const float keep = up ? a[i + s] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return a[0];
}

// CTA2 = true: the CTA-pair instance (launched as clusters of two; contains the cta_group::2 instructions, which make a kernel
// cluster-only — a plain launch of it fails with "cluster misconfiguration", hence two instances).
template <bool CTA2>
__global__ void __launch_bounds__(kThreadsFa, 1) conv_tc_kernel_t(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ __align__(8) uint64_t aready_bar[4];   // hs mode: operand strip buffer written (converter warps / TMA)
  __shared__ __align__(8) uint64_t afree_bar[4];    // hs mode: every MMA reading that strip buffer has retired
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(8) uint64_t sfull_bar[2];    // vt mode: staging slot filled by TMA
  __shared__ __align__(8) uint64_t sempty_bar[2];   // vt mode: staging slot read by every converter warp
  __shared__ float2 st_x_all[2][4][32];        // per epilogue group and warp: (sum, M2) of one 32-channel unit, merged per tile
  __shared__ float st_n_all[2][4];

  // 128B-swizzled TMA/UMMA tiles need 1024 B alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  uint32_t crank = 0u;                                          // 0 = leader (issues the MMAs), 1 = follower
  if constexpr (CTA2) crank = cluster_ctarank();
  const int b_bytes = (CTA2 ? p.n_tile / 2 : p.n_tile) * 128; // rows of the weight tile THIS CTA holds
  const int vs_a_bytes = p.vs_rows * p.tile_w * 128;                 // one plane of an A strip (vs mode)
  const int stage_bytes = p.hs ? p.planes * b_bytes
                               : (p.vs ? p.planes * vs_a_bytes + (p.fa == 2 ? kStemPatchBytes : 0) : p.planes * (kABytes + b_bytes));
  const int kch0 = p.kchunks[0];
  const int strip_bytes = p.planes * p.hs_plane_bytes;              // one operand strip buffer (hs mode)
  const int bres_bytes = p.vs ? p.ntaps * kch0 * p.planes * b_bytes : (p.hs ? p.hs_nbuf * strip_bytes : 0);
  const int vt_slot_bytes = p.vt ? p.vt_half_rows * p.tile_w * 256 : (p.ht ? p.ht_slot_bytes : 0);   // fp32 staging slot
  uint8_t* const vt_stage = smem + bres_bytes;                       // vt mode: two staging slots behind the resident weights
  uint8_t* const stage_base = smem + bres_bytes + 2 * vt_slot_bytes; // resident weights / operand strip first, then the stage ring
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_c;
  const uint32_t acc_cols = static_cast<uint32_t>(p.nacc * p.n_tile);    // one accumulator set
  uint32_t tmem_cols = 32;                                               // two sets, rounded up to a power of two >= 32
  while (tmem_cols < 2u * acc_cols) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nsrc && !p.fa; ++s) {
      prefetch_tensormap(&p.a_hi[s]);
      if (p.planes == 2) prefetch_tensormap(&p.a_lo[s]);
    }
    prefetch_tensormap(&p.b_hi);
    if (p.planes == 2) prefetch_tensormap(&p.b_lo);
    if (blockIdx.x == 0 && p.st_S != nullptr) *p.st_S = p.st_S_total;
    // a stage is full when the TMA bytes have landed (producer's expect_tx arrival) and, in fused-operand mode, every
    // converter warp has written its share of the A planes
    const uint32_t full_count = ((p.fa && p.vs) ? 0u : 1u) + ((p.fa && !p.hs) ? static_cast<uint32_t>(kConvWarps) : 0u);
    for (int b = 0; b < 4; ++b) {
      mbar_init(&aready_bar[b], p.hp ? 1 : (CTA2 ? 2 * kConvWarps : kConvWarps));   // cta2: both CTAs' converters arrive on the leader's
      mbar_init(&afree_bar[b], 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&sfull_bar[b], 1); mbar_init(&sempty_bar[b], kConvWarps); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], full_count); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], (CTA2 || p.epi2) ? 8 : 4); }
    mbar_init(&bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) { tmem_alloc_2sm(&tmem_base_smem, tmem_cols); tmem_relinquish_2sm(); }
    else { tmem_alloc(&tmem_base_smem, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();       // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int s = 0; uint32_t ph = 0;
      if (p.hs) {
        // weights only: one (chunk, tap) tile per stage, chunk-major so the strip of a chunk serves all its taps
        uint32_t g = 0;
        for (int t = blockIdx.x; t - static_cast<int>(crank) < total_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(p, t);
          for (int src = 0; src < p.nsrc; ++src)
            for (int kc = 0; kc < p.kchunks[src]; ++kc, ++g) {
              if (p.hp) {
                // the chunk's operand strip: one TMA box per plane, zero padding = out-of-bounds fill
                const uint32_t buf = g % static_cast<uint32_t>(p.hs_nbuf);
                mbar_wait(&afree_bar[buf], ((g / static_cast<uint32_t>(p.hs_nbuf)) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&aready_bar[buf], static_cast<uint32_t>(p.planes * p.hs_rows * p.hs_cols * 128));
                uint8_t* st = smem + static_cast<size_t>(buf) * strip_bytes;
                tma_load_5d(st, &p.a_hi[src], &aready_bar[buf], kc * kKC, tc.w0 + p.hs_dw_min, tc.h0 + p.hs_dh_min, tc.n0, 0);
                if (p.planes == 2)
                  tma_load_5d(st + p.hs_plane_bytes, &p.a_lo[src], &aready_bar[buf], kc * kKC, tc.w0 + p.hs_dw_min,
                              tc.h0 + p.hs_dh_min, tc.n0, 0);
              }
              for (int tap = 0; tap < p.ntaps; ++tap) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sb = stage_base + static_cast<size_t>(s) * stage_bytes;
                const int kw = p.src_koff[src] + kc * kKC;
                if constexpr (CTA2) {
                  // each CTA loads its half of the weight rows into its own stage; both halves complete on the LEADER's barrier
                  if (crank == 0) mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(2 * stage_bytes));
                  // (one channel tile only in this mode: the pair's weight rows are [0, n_tile), also when the follower's own
                  // tile index lies past the end of an odd tile count)
                  const int co = static_cast<int>(crank) * (p.n_tile / 2);
                  tma_load_3d_2sm(sb, &p.b_hi, &full_bar[s], kw, co, p.tap_w[tap]);
                  if (p.planes == 2) tma_load_3d_2sm(sb + b_bytes, &p.b_lo, &full_bar[s], kw, co, p.tap_w[tap]);
                } else {
                  mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(stage_bytes));
                  tma_load_3d(sb, &p.b_hi, &full_bar[s], kw, tc.cout0, p.tap_w[tap]);
                  if (p.planes == 2) tma_load_3d(sb + b_bytes, &p.b_lo, &full_bar[s], kw, tc.cout0, p.tap_w[tap]);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
              }
            }
        }
      } else if (p.vs) {
        // resident weights: every tap / channel chunk / plane once per CTA
        mbar_arrive_expect_tx(&bres_bar, static_cast<uint32_t>(bres_bytes));
        for (int tap = 0; tap < p.ntaps; ++tap)
          for (int kc = 0; kc < kch0; ++kc) {
            uint8_t* dst = smem + static_cast<size_t>((tap * kch0 + kc) * p.planes) * b_bytes;
            tma_load_3d(dst, &p.b_hi, &bres_bar, kc * kKC, 0, p.tap_w[tap]);
            if (p.planes == 2) tma_load_3d(dst + b_bytes, &p.b_lo, &bres_bar, kc * kKC, 0, p.tap_w[tap]);
          }
        if (p.vt) {
          prefetch_tensormap(&p.a_f32);
          uint32_t hn = 0;                                   // running half-strip number: slot hn & 1, phase (hn >> 1) & 1
          for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(p, t);
            for (int kc = 0; kc < kch0; ++kc)
              for (int half = 0; half < 2; ++half, ++hn) {
                const uint32_t slot = hn & 1u;
                mbar_wait(&sempty_bar[slot], ((hn >> 1) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&sfull_bar[slot], static_cast<uint32_t>(vt_slot_bytes));
                tma_load_4d(vt_stage + slot * vt_slot_bytes, &p.a_f32, &sfull_bar[slot], kc * kKC,
                            tc.w0 + p.tap_off[0][1] - p.fa_border,
                            tc.h0 + p.vs_dh_min - p.fa_border + half * p.vt_half_rows, tc.n0);
              }
          }
        }
        if (!p.fa)
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          const TileCoord tc = decode_tile(p, t);
          for (int kc = 0; kc < kch0; ++kc) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(stage_bytes));
            uint8_t* st = stage_base + static_cast<size_t>(s) * stage_bytes;
            const int cw = tc.w0 + p.tap_off[0][1], chh = tc.h0 + p.vs_dh_min;
            tma_load_5d(st, &p.a_hi[0], &full_bar[s], kc * kKC, cw, chh, tc.n0, 0);
            if (p.planes == 2) tma_load_5d(st + vs_a_bytes, &p.a_lo[0], &full_bar[s], kc * kKC, cw, chh, tc.n0, 0);
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
        }
      } else
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        for (int tap = 0; tap < p.ntaps; ++tap) {
          int base[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) {
            const int sel = p.dim_sel[i];
            base[i] = p.tap_off[tap][i] + (sel == 1 ? tc.w0 : sel == 2 ? tc.h0 : sel == 3 ? tc.n0 : 0);
          }
          for (int src = 0; src < p.nsrc; ++src) {
            for (int kc = 0; kc < p.kchunks[src]; ++kc) {
              mbar_wait(&empty_bar[s], ph ^ 1);
              mbar_arrive_expect_tx(&full_bar[s], static_cast<uint32_t>(p.fa ? p.planes * b_bytes : stage_bytes));
              int c[5];
#pragma unroll
              for (int i = 0; i < 5; ++i) c[i] = base[i] + (p.dim_sel[i] == 0 ? kc * kKC : 0);
              uint8_t* st = stage_base + static_cast<size_t>(s) * stage_bytes;
              if (!p.fa) {
                tma_load_5d(st, &p.a_hi[src], &full_bar[s], c[0], c[1], c[2], c[3], c[4]);
                if (p.planes == 2) tma_load_5d(st + kABytes, &p.a_lo[src], &full_bar[s], c[0], c[1], c[2], c[3], c[4]);
              }
              const int kw = p.src_koff[src] + kc * kKC;
              uint8_t* sb = st + p.planes * kABytes;
              tma_load_3d(sb, &p.b_hi, &full_bar[s], kw, tc.cout0, p.tap_w[tap]);
              if (p.planes == 2) tma_load_3d(sb + b_bytes, &p.b_lo, &full_bar[s], kw, tc.cout0, p.tap_w[tap]);
              if (++s == p.stages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer =====================
      // the whole warp runs the loop (convergent: descriptors may live in uniform registers), one elected lane issues
      int s = 0; uint32_t ph = 0;
      int acc = 0; uint32_t acc_ph = 0;
      const int k_iters = p.ntaps * (p.kchunks[0] + (p.nsrc > 1 ? p.kchunks[1] : 0));
      if (p.hs && crank != 0) {
        // follower of a CTA pair: the leader issues the MMAs for both
      } else if (p.hs) {
        const uint32_t sbo = static_cast<uint32_t>(p.hs_cols) * 128u;
        uint32_t g = 0;                                  // running chunk number: strip buffer g % nbuf, phase (g / nbuf) & 1
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
          tc_fence_after();
          const uint32_t d_set = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
          uint32_t started = 0;                          // bit a: accumulator a has received its first MMA of this tile
          const int nchunks = p.kchunks[0] + (p.nsrc > 1 ? p.kchunks[1] : 0);
          for (int ch = 0; ch < nchunks; ++ch, ++g) {
            const uint32_t buf = g % static_cast<uint32_t>(p.hs_nbuf);
            const uint32_t strip_hi = smem_u32(smem) + buf * static_cast<uint32_t>(strip_bytes);
            const uint32_t strip_lo = strip_hi + static_cast<uint32_t>(p.hs_plane_bytes);
            mbar_wait(&aready_bar[buf], (g / static_cast<uint32_t>(p.hs_nbuf)) & 1u);   // converters wrote this strip
            tc_fence_after();
            for (int tap = 0; tap < p.ntaps; ++tap) {
              mbar_wait(&full_bar[s], ph);              // this tap's weights have landed
              tc_fence_after();
              const uint32_t aoff = static_cast<uint32_t>(p.hs_off[tap]);
              const uint32_t b_hi = smem_u32(stage_base + static_cast<size_t>(s) * stage_bytes);
              const uint32_t b_lo = b_hi + b_bytes;
              const int a_i = p.tap_acc[tap];
              const uint32_t d_tmem = d_set + static_cast<uint32_t>(a_i * p.n_tile);
              uint32_t accumulate = (started >> a_i) & 1u;
              started |= 1u << a_i;
              {
                // one K = 64 chunk of this tap: four K = 16 steps issued from one asm block (ptx.cuh)
                const uint64_t da_hi = make_sw128_kmajor_desc_sbo(strip_hi + aoff, sbo);
                const uint64_t db_hi = make_sw128_kmajor_desc(b_hi);
                if (p.planes == 2) {
                  const uint64_t da_lo = make_sw128_kmajor_desc_sbo(strip_lo + aoff, sbo);
                  const uint64_t db_lo = make_sw128_kmajor_desc(b_lo);
                  // CTA2: M = 256 across the pair — the hardware takes rows 0-127 of A (and of B) from this CTA's shared memory
                  // and rows 128-255 from the same offsets in the peer's
                  if constexpr (CTA2) umma_f16_2sm_k64x3_elect(d_tmem, da_hi, da_lo, db_hi, db_lo, p.idesc, accumulate);
                  else umma_f16_k64x3_elect(d_tmem, da_hi, da_lo, db_hi, db_lo, p.idesc, accumulate);
                } else {
                  if constexpr (CTA2) umma_f16_2sm_k64_elect(d_tmem, da_hi, db_hi, p.idesc, accumulate);
                  else umma_f16_k64_elect(d_tmem, da_hi, db_hi, p.idesc, accumulate);
                }
                accumulate = 1;
              }
              if constexpr (CTA2) umma_commit_2sm_elect(&empty_bar[s]); else umma_commit_elect(&empty_bar[s]);
              if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            if constexpr (CTA2) umma_commit_2sm_elect(&afree_bar[buf]); else umma_commit_elect(&afree_bar[buf]);   // strip buffer reusable once these MMAs retire
          }
          if constexpr (CTA2) umma_commit_2sm_elect(&tfull_bar[acc]); else umma_commit_elect(&tfull_bar[acc]);
          acc ^= 1; if (acc == 0) acc_ph ^= 1;
        }
      } else if (p.vs) {
        mbar_wait(&bres_bar, 0);                      // resident weights have landed
        tc_fence_after();
        const uint32_t bres = smem_u32(smem);
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
          uint32_t accumulate = 0;
          for (int kc = 0; kc < kch0; ++kc) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(stage_base + static_cast<size_t>(s) * stage_bytes);
            const uint32_t a_lo = a_hi + vs_a_bytes;
            for (int tap = 0; tap < p.ntaps; ++tap) {
              const uint32_t aoff = static_cast<uint32_t>(p.vs_row_off[tap]) * p.tile_w * 128u;   // whole image rows
              const uint32_t b_hi = bres + static_cast<uint32_t>((tap * kch0 + kc) * p.planes) * b_bytes;
              const uint32_t b_lo = b_hi + b_bytes;
              {
                const uint64_t da_hi = make_sw128_kmajor_desc(a_hi + aoff);
                const uint64_t db_hi = make_sw128_kmajor_desc(b_hi);
                if (p.planes == 2)
                  umma_f16_k64x3_elect(d_tmem, da_hi, make_sw128_kmajor_desc(a_lo + aoff), db_hi, make_sw128_kmajor_desc(b_lo), p.idesc,
                                       accumulate);
                else
                  umma_f16_k64_elect(d_tmem, da_hi, db_hi, p.idesc, accumulate);
                accumulate = 1;
              }
            }
            umma_commit_elect(&empty_bar[s]);
            if (++s == p.stages) { s = 0; ph ^= 1; }
          }
          umma_commit_elect(&tfull_bar[acc]);
          acc ^= 1; if (acc == 0) acc_ph ^= 1;
        }
      } else
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc) * acc_cols;
        uint32_t accumulate = 0;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[s], ph);                // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_hi = smem_u32(stage_base + static_cast<size_t>(s) * stage_bytes);
          const uint32_t a_lo = a_hi + kABytes;
          const uint32_t b_hi = a_hi + p.planes * kABytes;
          const uint32_t b_lo = b_hi + b_bytes;
          {
            const uint64_t da_hi = make_sw128_kmajor_desc(a_hi);
            const uint64_t db_hi = make_sw128_kmajor_desc(b_hi);
            if (p.planes == 2)                      // small terms first inside every K = 16 step
              umma_f16_k64x3_elect(d_tmem, da_hi, make_sw128_kmajor_desc(a_lo), db_hi, make_sw128_kmajor_desc(b_lo), p.idesc, accumulate);
            else
              umma_f16_k64_elect(d_tmem, da_hi, db_hi, p.idesc, accumulate);
            accumulate = 1;
          }
          umma_commit_elect(&empty_bar[s]);                 // smem stage free once these MMAs retire
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit_elect(&tfull_bar[acc]);                 // accumulator complete -> epilogue
        acc ^= 1; if (acc == 0) acc_ph ^= 1;
      }
    }
  } else if (warp == 14) {
    // ===================== fp32 source staging producer (ht mode) =====================
    if (lane == 0 && p.ht) {
      prefetch_tensormap(&p.a_f32);
      uint32_t g = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        for (int kc = 0; kc < p.kchunks[0]; ++kc, ++g) {
          const uint32_t slot = g & 1u;
          mbar_wait(&sempty_bar[slot], ((g >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sfull_bar[slot], static_cast<uint32_t>(p.hs_rows * p.hs_cols * 256));
          tma_load_4d(vt_stage + slot * vt_slot_bytes, &p.a_f32, &sfull_bar[slot], kc * kKC,
                      tc.w0 + p.hs_dw_min - p.fa_border, tc.h0 + p.hs_dh_min - p.fa_border, tc.n0);
        }
      }
    }
  } else if (warp < 6 || (p.epi2 && warp < 10)) {
    // ===================== epilogue: TMEM -> registers -> (+bias) -> fp32 NHWC =====================
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;                  // epilogue group (1 only with epi2)
    float2 (*st_x)[32] = st_x_all[grp];
    float* st_n = st_n_all[grp];
    auto group_sync = [&]() {                         // the four warps of this group (ids 2 / 3 are used by group 0's two paths)
      if (grp == 0) asm volatile("bar.sync 3, 128;" ::: "memory"); else asm volatile("bar.sync 4, 128;" ::: "memory");
    };
    int acc = 0; uint32_t acc_ph = 0;
    const int ets_lw = 31 - __clz(p.tile_w), ets_lh = 31 - __clz(p.tile_h);
    uint32_t ets_g = 0;                               // 32-channel groups written so far (ets == 2: buffer ets_g & 1)
    const int m = q * 32 + lane;
    const int w_l = m % p.tile_w;
    const int h_l = (m / p.tile_w) % p.tile_h;
    const int n_l = m / (p.tile_w * p.tile_h);
    for (int t = blockIdx.x; t - static_cast<int>(crank) < total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      const int n = tc.n0 + n_l, h = tc.h0 + h_l, w = tc.w0 + w_l;
      const bool valid = (t < total_tiles) && (n < p.N) && (h < p.OH) && (w < p.OW);
      float* const yp0 = p.y + n * p.ys_n + h * p.ys_h + w * p.ys_w + tc.cout0;
      mbar_wait_sleep(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      // ets == 1: the eight buffer rows this thread copies out (row_i = q*32 + i*4 + lane/8): element offset of their pixel in y
      // and whether it lies inside the output — once per tile, not per 32-channel unit (host guarantees 32-bit offsets)
      uint32_t co_off[8]; uint32_t co_ok = 0;
      if (p.ets == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = q * 32 + i * 4 + (lane >> 3);
          const int rw = row & (p.tile_w - 1), rh = (row >> ets_lw) & (p.tile_h - 1), rn = row >> (ets_lw + ets_lh);
          const int n2 = tc.n0 + rn, h2 = tc.h0 + rh, w2 = tc.w0 + rw;
          co_off[i] = static_cast<uint32_t>(n2 * p.ys_n + h2 * p.ys_h + w2 * p.ys_w);
          if (t < total_tiles && n2 < p.N && h2 < p.OH && w2 < p.OW) co_ok |= 1u << i;
        }
      }
      // slice of the statistics workspace this warp's 32 rows belong to (tile_n == 1 when stats are fused)
      const int th_i = tc.h0 / p.tile_h, tw_i = tc.w0 / p.tile_w;
      for (int ai = 0; ai < p.nacc; ++ai) {                // merged ConvTranspose phases: one accumulator per output parity
      float* const yp = yp0 + (p.nacc > 1 ? p.acc_ybase[ai] : p.y_base);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc) * acc_cols +
                             static_cast<uint32_t>(ai * p.n_tile);
      // one statistics slice per CTA tile (tile_n == 1: the tile lies in image tc.n0); the four epilogue warps merge their
      // 32-pixel partials in shared memory before anything is written
      const long long st_row = static_cast<long long>(tc.n0) * p.st_S_cap + (p.nacc > 1 ? p.acc_slice[ai] : p.st_slice_base) +
                               (th_i * p.tiles_w + tw_i);
      for (int c = 0; c < p.n_tile; c += 32) {
        if (p.epi2 && (((ai * (p.n_tile >> 5) + (c >> 5)) & 1) != grp)) continue;      // the other group's unit
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        const bool cvalid = (t < total_tiles) && (tc.cout0 + c) < p.cout_total;
        if (p.bias != nullptr && cvalid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + tc.cout0 + c + j);
            v[j + 0] = __float_as_uint(__uint_as_float(v[j + 0]) + b.x);
            v[j + 1] = __float_as_uint(__uint_as_float(v[j + 1]) + b.y);
            v[j + 2] = __float_as_uint(__uint_as_float(v[j + 2]) + b.z);
            v[j + 3] = __float_as_uint(__uint_as_float(v[j + 3]) + b.w);
          }
        }
        if (p.ets) {
          // ---- epilogue through shared memory: a thread owns one pixel (256 B .. 1 KB apart from its neighbours'), so direct
          // stores are 32 half-used sectors per warp instruction and the LSU becomes the bound of every layer with little MMA
          // work per output byte (measured: ~2 cycles per sector, 2k cycles per 32-channel group).  The group is transposed
          // through one 16 KB buffer instead (row = pixel m, 16-byte chunk j at j ^ (m & 7): conflict-free both ways), stored
          // as full 128-byte lines, and its statistics are taken from the same buffer (lane = channel: 32 conflict-free
          // loads, register adds) instead of two 31-shuffle butterflies.
          uint8_t* const bufp = smem + p.ets_off + (p.ets == 2 ? (ets_g & 1u) * 16384u : static_cast<uint32_t>(grp) * 16384u);
          const uint32_t buf = smem_u32(bufp);
          ++ets_g;
          if (p.ets == 2) {
            if (q == 0 && lane == 0) bulk_wait_group_read1();          // the TMA store that last used this buffer has read it
            group_sync();
          }
          // the buffer is 1024-byte aligned, so "row base + swizzled chunk" is an XOR of bits 4-6: one LOP3 per store
          const uint32_t rowx = (buf + static_cast<uint32_t>(m) * 128u) ^ (static_cast<uint32_t>(m & 7) << 4);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowx ^ (static_cast<uint32_t>(j) << 4)), "r"(v[4 * j]),
                         "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
          if (p.ets == 2) fence_proxy_async();
          group_sync();            // the whole group is in the buffer
          if (p.ets == 2) {
            if (q == 0 && lane == 0 && cvalid) {                      // out-of-range pixels / images are clipped by the tensor map
              tma_store_4d(&p.y_map[ai], bufp, tc.cout0 + c, tc.w0, tc.h0, tc.n0);
              bulk_commit_group();
            }
          } else if (cvalid) {
            float* const yc = p.y + (p.nacc > 1 ? p.acc_ybase[ai] : p.y_base) + tc.cout0 + c + (lane & 7) * 4;
            // row_i & 7 = ((i & 1) * 4 + lane / 8): two swizzle phases, the rest of the address is an immediate
            const uint32_t cbase = buf + static_cast<uint32_t>(q * 32 + (lane >> 3)) * 128u;
            const uint32_t ce = cbase + (static_cast<uint32_t>((lane & 7) ^ (lane >> 3)) << 4);
            const uint32_t cod = cbase + (static_cast<uint32_t>((lane & 7) ^ (4 + (lane >> 3))) << 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              uint32_t x0, x1, x2, x3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                           : "r"(((i & 1) ? cod : ce) + static_cast<uint32_t>(i) * 512u) : "memory");
              if ((co_ok >> i) & 1u) *reinterpret_cast<uint4*>(yc + co_off[i]) = make_uint4(x0, x1, x2, x3);
            }
          }
          if (p.st_partial != nullptr && cvalid) {
            // lane = channel c + lane over this warp's 32 pixel rows (row & 7 == i & 7): (count, sum, M2 about the mean)
            // row q*32 + i, chunk (lane / 4) ^ (i & 7): the XOR touches bits 4-6 only, the row offset i * 128 stays an immediate
            const uint32_t cb = buf + static_cast<uint32_t>(q * 32) * 128u + (static_cast<uint32_t>(lane >> 2) << 4) +
                                static_cast<uint32_t>(lane & 3) * 4u;
            // the additions follow the tree of the shuffle butterfly (rows i and i + 16, then + 8, 4, 2, 1), so both epilogues
            // produce bit-identical statistics and a tile's result does not depend on which one its launch shape selects
            float xs[32], tr[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float x;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"((cb ^ (static_cast<uint32_t>(i & 7) << 4)) + static_cast<uint32_t>(i) * 128u)
                           : "memory");
              xs[i] = x;
              tr[i] = ((vmask >> i) & 1u) ? x : 0.f;
            }
#pragma unroll
            for (int sp = 16; sp >= 1; sp >>= 1) {
#pragma unroll
              for (int i = 0; i < sp; ++i) tr[i] = tr[i] + tr[i + sp];
            }
            const float sum = tr[0];
            const float cntf = static_cast<float>(__popc(vmask));
            const float mean = cntf > 0.f ? sum / cntf : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float d = xs[i] - mean; tr[i] = ((vmask >> i) & 1u) ? d * d : 0.f; }
#pragma unroll
            for (int sp = 16; sp >= 1; sp >>= 1) {
#pragma unroll
              for (int i = 0; i < sp; ++i) tr[i] = tr[i] + tr[i + sp];
            }
            const float m2 = tr[0];
            st_x[q][lane] = make_float2(sum, m2);
            if (lane == 0) st_n[q] = cntf;
          }
          group_sync();            // buffer and partials read / written by everyone
          if (p.st_partial != nullptr && cvalid && q == 0) {
            float nt = 0.f, st = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { nt += st_n[k]; st += st_x[k][lane].x; }
            const float mt = nt > 0.f ? st / nt : 0.f;
            float m2t = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (st_n[k] > 0.f) { const float d = st_x[k][lane].x / st_n[k] - mt; m2t += st_x[k][lane].y + st_n[k] * d * d; }
            p.st_partial[st_row * p.cout_total + tc.cout0 + c + lane] = make_float2(st, m2t);
            if (tc.cout0 == 0 && c == 0 && lane == 0) p.st_cnt[st_row] = nt;
          }
          // (the next group's first barrier orders this merge's reads of st_x before the next partials are written)
          continue;
        } else if (valid && cvalid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = __uint_as_float(v[j + 0]); o.y = __uint_as_float(v[j + 1]);
            o.z = __uint_as_float(v[j + 2]); o.w = __uint_as_float(v[j + 3]);
            *reinterpret_cast<float4*>(yp + c + j) = o;
          }
        }
        if (p.st_partial != nullptr && cvalid) {
          // column statistics over this warp's 32 pixels (lane = pixel, v[j] = channel c + j): two shuffle butterflies,
          // no shared memory; lane l ends up with (sum, M2 about the slice mean) of channel c + l
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = valid ? __uint_as_float(v[j]) : 0.f;
          const float sum = warp_colsum32(a, lane);
          const float cntf = static_cast<float>(__popc(vmask));
          const float mean = cntf > 0.f ? sum / cntf : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) - __shfl_sync(0xffffffffu, mean, j);
            a[j] = valid ? d * d : 0.f;
          }
          const float m2 = warp_colsum32(a, lane);
          st_x[q][lane] = make_float2(sum, m2);
          if (lane == 0) st_n[q] = cntf;
          if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 5, 128;" ::: "memory");            // the four epilogue warps
          if (q == 0) {
            // merge in warp-quarter order (fixed => deterministic): n = sum n_q, S = sum S_q, M2 = sum (M2_q + n_q (mean_q - mean)^2)
            float nt = 0.f, st = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { nt += st_n[k]; st += st_x[k][lane].x; }
            const float mt = nt > 0.f ? st / nt : 0.f;
            float m2t = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (st_n[k] > 0.f) { const float d = st_x[k][lane].x / st_n[k] - mt; m2t += st_x[k][lane].y + st_n[k] * d * d; }
            p.st_partial[st_row * p.cout_total + tc.cout0 + c + lane] = make_float2(st, m2t);
            if (tc.cout0 == 0 && c == 0 && lane == 0) p.st_cnt[st_row] = nt;
          }
          if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 5, 128;" ::: "memory");            // st_x is reused by the next channel group
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) { if (crank == 0) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_remote(&tempty_bar[acc], 0); }
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1; if (acc == 0) acc_ph ^= 1;
    }
    if (p.ets == 2 && q == 0 && lane == 0) bulk_wait_group0();   // the last boxes have left shared memory and reached global memory
  } else if (p.fa && warp < 14) {
    // ===================== operand converters (fused-operand mode) =====================
    // 256 threads; thread = (q: one float4 = 4 channels of the 64-channel chunk, pr: pixel row group).  A warp-wide
    // float4 load covers two pixels x 256 contiguous bytes; each thread writes 8 B of the hi and 8 B of the lo plane at
    // the swizzled position TMA would have used: row m, 16-byte chunk c -> m*128 + ((c ^ (m & 7)) << 4).
    const int ct = threadIdx.x - 6 * 32;
    const int q = ct & 15;
    const int pr = ct >> 4;                                   // 0..15
    const uint32_t sub = static_cast<uint32_t>((q & 1) * 8);
    const int chunk = q >> 1;
    int s = 0; uint32_t ph = 0;
    const int lw = 31 - __clz(p.tile_w), lh = 31 - __clz(p.tile_h);
    if (p.fa == 2) {
      // ---- fused stem: the window operand lane (s * 8 + c) of strip pixel (row, col) is pad(x)[n, c, row, col + s].  Per tile
      // the (vs_rows) x (8 + S - 1) x C input patch is staged once in shared memory (coalesced row loads, zero / reflect
      // border resolved there; the NEXT tile's patch is already in flight in registers while this one is expanded), then
      // every converter thread expands its strip units from the patch: no global latency inside the expansion.
      const int prow = p.vs_rows, pcol = 8 + p.stem_S - 1;          // patch extents (pcol <= 15)
      const int pelems = p.stem_C * prow * pcol;
      const long long plane = static_cast<long long>(p.Hs) * p.Ws;
      constexpr int kPE = 6;                                        // patch elements per thread: 4 x 24 x 15 = 1440 <= 6 x 256
      // element k of this thread: (channel, patch row, patch col) — tile-independent, decoded once (three integer divisions
      // per element would otherwise be paid for every tile)
      int pe_c[kPE], pe_r[kPE], pe_cc[kPE];
#pragma unroll
      for (int k = 0; k < kPE; ++k) {
        const int e = ct + 256 * k;
        pe_c[k] = -1; pe_r[k] = 0; pe_cc[k] = 0;
        if (e < pelems) { pe_c[k] = e / (prow * pcol); const int rem = e - pe_c[k] * (prow * pcol); pe_r[k] = rem / pcol; pe_cc[k] = rem - pe_r[k] * pcol; }
      }
      auto patch_load = [&](int t, float (&v)[kPE]) {
        const TileCoord tc = decode_tile(p, t);
        const float* xn = p.fa_x[0] + static_cast<long long>(tc.n0) * p.stem_C * plane;
#pragma unroll
        for (int k = 0; k < kPE; ++k) {
          v[k] = 0.f;
          if (pe_c[k] >= 0 && tc.n0 < p.N) {
            int hh = tc.h0 + p.vs_dh_min + pe_r[k] - p.stem_pad, ww = tc.w0 + pe_cc[k] - p.stem_pad;
            if (p.fa_border_mode == DLB_PAD_REFLECT) {
              if (hh < 0) hh = -hh; if (hh >= p.Hs) hh = 2 * p.Hs - 2 - hh;
              if (ww < 0) ww = -ww; if (ww >= p.Ws) ww = 2 * p.Ws - 2 - ww;
            }
            if (hh >= 0 && hh < p.Hs && ww >= 0 && ww < p.Ws)          // zero border / tile overhang beyond the image
              v[k] = __ldg(xn + pe_c[k] * plane + hh * p.Ws + ww);
          }
        }
      };
      float pre[kPE], pre2[kPE];                                    // patches of this tile and the next, in flight
      int t = blockIdx.x;
      if (t < total_tiles) patch_load(t, pre);
      if (t + static_cast<int>(gridDim.x) < total_tiles) patch_load(t + gridDim.x, pre2);
      const int npx = p.vs_rows * p.tile_w;
      const int st = q >> 1;
      const bool lane_live = ((q & 1) == 0) && (st < p.stem_S);
      for (; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_hi = stage_base + static_cast<size_t>(s) * stage_bytes;
        uint8_t* a_lo = a_hi + vs_a_bytes;
        float* patch = reinterpret_cast<float*>(a_hi + p.planes * vs_a_bytes);      // [c][row][16]
#pragma unroll
        for (int k = 0; k < kPE; ++k)
          if (pe_c[k] >= 0) patch[(pe_c[k] * 24 + pe_r[k]) * 16 + pe_cc[k]] = pre[k];
        asm volatile("bar.sync 1, 256;" ::: "memory");                 // converter warps only: the patch is complete
#pragma unroll
        for (int k = 0; k < kPE; ++k) pre[k] = pre2[k];
        if (t + 2 * static_cast<int>(gridDim.x) < total_tiles) patch_load(t + 2 * gridDim.x, pre2);   // two tiles ahead
        const bool row_ok = true;
        for (int base = 0; base < npx; base += 16) {
          const int idx = base + pr;
          if (idx >= npx) break;
          const int row = idx >> 3, col = idx & 7;
          const int vh = tc.h0 + p.vs_dh_min + row, vw = tc.w0 + col;
          float o[4] = {0.f, 0.f, 0.f, 0.f};
          if (row_ok && lane_live && tc.n0 < p.N && vh >= 0 && vh < p.H && vw >= 0 && vw < p.W) {
            const float* pp = patch + row * 16 + col + st;
            o[0] = pp[0];
            if (p.stem_C > 1) o[1] = pp[24 * 16];
            if (p.stem_C > 2) o[2] = pp[2 * 24 * 16];
            if (p.stem_C > 3) o[3] = pp[3 * 24 * 16];
          }
          uint2 hi, lo;
          fa_split4(o, p.fa_is_bf16, hi, lo);
          const uint32_t off = static_cast<uint32_t>(idx) * 128u + (static_cast<uint32_t>(chunk ^ (idx & 7)) << 4) + sub;
          *reinterpret_cast<uint2*>(a_hi + off) = hi;
          if (p.planes == 2) *reinterpret_cast<uint2*>(a_lo + off) = lo;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    } else if (p.hs) {
      // Strip units: unit u = (strip pixel u >> 4, float4 q = u & 15); thread owns units ct + 256 * j.  The strip is double-
      // buffered: the converters build chunk c + 1 while the MMAs read chunk c.  Per-unit facts that do not depend on the
      // tile (shared-memory offset, pixel offset inside the strip, "belongs to the tile's own 16 x 8 pixels") are computed
      // once; a tile whose strip lies wholly inside the source needs no per-pixel border logic at all.
      constexpr int kUnits = kHsMaxPx * 16 / 256;              // units per thread for the largest admissible strip
      const int npx = p.hs_rows * p.hs_cols;
      const int nunits = (npx * 16 - ct + 255) / 256;          // units this thread owns
      uint32_t soff[kUnits];                                   // swizzled byte offset inside a strip plane
      int poff[kUnits];                                        // row * Ws + col: source pixel offset relative to the strip origin
      uint32_t own_mask = 0;                                   // bit j: unit j's pixel is one of the tile's own output pixels
#pragma unroll
      for (int j = 0; j < kUnits; ++j) {
        const int px_i = (ct + 256 * j) >> 4;
        const int row = px_i / p.hs_cols, col = px_i - row * p.hs_cols;
        poff[j] = row * p.Ws + col;
        // absolute-address swizzle: the strip buffers are 1024 B aligned, so address bits [7,10) = px_i & 7
        soff[j] = static_cast<uint32_t>(px_i) * 128u + (static_cast<uint32_t>(chunk ^ (px_i & 7)) << 4) + sub;
        const int oh_l = row + p.hs_dh_min - p.fa_border, ow_l = col + p.hs_dw_min - p.fa_border;
        if (oh_l >= 0 && oh_l < p.tile_h && ow_l >= 0 && ow_l < p.tile_w) own_mask |= 1u << j;
      }
      uint32_t g = 0;                                          // running chunk number (same sequence as the MMA issuer)
      for (int t = blockIdx.x; t - static_cast<int>(crank) < total_tiles; t += gridDim.x) {
        TileCoord tc = decode_tile(p, t);
        if (t >= total_tiles) tc.n0 = p.N;                     // odd tile count: the pair's second tile does not exist -> zeros
        const int vh0 = tc.h0 + p.hs_dh_min, vw0 = tc.w0 + p.hs_dw_min;
        const int sh0 = vh0 - p.fa_border, sw0 = vw0 - p.fa_border;      // strip origin in source coordinates
        const bool inside = (tc.n0 < p.N) && sh0 >= 0 && sw0 >= 0 && sh0 + p.hs_rows <= p.Hs && sw0 + p.hs_cols <= p.Ws;
        const long long origin = (static_cast<long long>(tc.n0) * p.Hs + sh0) * p.Ws + sw0;
        for (int src = 0; src < p.nsrc; ++src) {
          const bool wb = (p.fa_out[src] != nullptr) && (tc.cout0 == 0);
          const int C = p.fa_cin[src];
          const float* __restrict__ xs = p.fa_x[src];
          const float* __restrict__ rs = p.fa_res[src];
          const int act = p.fa_act[src];
          for (int kc = 0; kc < p.kchunks[src]; ++kc, ++g) {
            const uint32_t buf = g % static_cast<uint32_t>(p.hs_nbuf);
            uint8_t* const strip_hi = smem + static_cast<size_t>(buf) * strip_bytes;
            uint8_t* const strip_lo = strip_hi + p.hs_plane_bytes;
            const int cbase = kc * kKC + q * 4;
            float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.fa_scale[src] != nullptr && tc.n0 < p.N) {
              sc = __ldg(reinterpret_cast<const float4*>(p.fa_scale[src] + static_cast<long long>(tc.n0) * C + cbase));
              sh = __ldg(reinterpret_cast<const float4*>(p.fa_shift[src] + static_cast<long long>(tc.n0) * C + cbase));
            }
            const float* const xo = xs + origin * C + cbase;          // only dereferenced when `inside`
            const float* const ro = rs != nullptr ? rs + origin * C + cbase : nullptr;
            if (p.ht) {
              // staged source: fp32 strip in shared memory ([pixel][64 ch], plain layout) -> transform -> swizzled strip
              const uint32_t slot = g & 1u;
              mbar_wait(&sfull_bar[slot], (g >> 1) & 1u);
              mbar_wait_sleep(&afree_bar[buf], ((g / static_cast<uint32_t>(p.hs_nbuf)) & 1u) ^ 1u);
              const uint8_t* stg = vt_stage + slot * vt_slot_bytes;
#pragma unroll
              for (int j = 0; j < kUnits; ++j) {
                if (j >= nunits) continue;
                const int px_i = (ct + 256 * j) >> 4;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                bool ok = inside;
                if (!inside) {                                   // border tile: positions outside the source are zero padding
                  const int row = px_i / p.hs_cols, col = px_i - row * p.hs_cols;
                  const int hh = sh0 + row, ww = sw0 + col;
                  ok = tc.n0 < p.N && hh >= 0 && hh < p.Hs && ww >= 0 && ww < p.Ws;
                }
                if (ok) {
                  const float4 v = *reinterpret_cast<const float4*>(stg + static_cast<size_t>(px_i) * 256 + q * 16);
                  o[0] = fa_act1(fmaf(v.x, sc.x, sh.x), act); o[1] = fa_act1(fmaf(v.y, sc.y, sh.y), act);
                  o[2] = fa_act1(fmaf(v.z, sc.z, sh.z), act); o[3] = fa_act1(fmaf(v.w, sc.w, sh.w), act);
                }
                uint2 hi, lo;
                fa_split4(o, p.fa_is_bf16, hi, lo);
                *reinterpret_cast<uint2*>(strip_hi + soff[j]) = hi;
                if (p.planes == 2) *reinterpret_cast<uint2*>(strip_lo + soff[j]) = lo;
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&sempty_bar[slot]);   // this warp has read the staging slot
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
              if constexpr (CTA2) { if (crank == 0) mbar_arrive(&aready_bar[buf]); else mbar_arrive_remote(&aready_bar[buf], 0); }
              else mbar_arrive(&aready_bar[buf]);
            }
              continue;
            }
            bool waited = false;
            // One latency wave per chunk when there is a single fp32 source (12 float4 in flight per thread), two waves of six
            // (value + residual) for the block's first conv: the loads are register-staged, so in-flight bytes are what the
            // 128-register budget allows.
            auto convert = [&](auto batch_tag) {
              constexpr int kBatch = decltype(batch_tag)::value;
              constexpr bool kRes = kBatch == 6;
#pragma unroll
              for (int j0 = 0; j0 < kUnits; j0 += kBatch) {
                float4 xv[kBatch], rv[kRes ? kBatch : 1]; int goff[kBatch]; uint32_t okm = 0, inm = 0;
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                  const int j = j0 + u;
                  xv[u] = make_float4(0.f, 0.f, 0.f, 0.f); goff[u] = 0;
                  if (kRes) rv[u] = xv[u];
                  if (j < nunits) {
                    if (inside) {
                      goff[u] = poff[j] * C;
                      okm |= 1u << u; inm |= 1u << u;
                      xv[u] = __ldg(reinterpret_cast<const float4*>(xo + goff[u]));
                      if (kRes) rv[u] = __ldg(reinterpret_cast<const float4*>(ro + goff[u]));
                    } else {
                      const int px_i = (ct + 256 * j) >> 4;               // border tile: per-pixel zero / reflect logic
                      const int row = px_i / p.hs_cols, col = px_i - row * p.hs_cols;
                      const FaPix px = fa_locate(p, src, tc.n0, vh0 + row, vw0 + col);
                      if (px.ok) {
                        okm |= 1u << u; if (px.interior) inm |= 1u << u;
                        goff[u] = static_cast<int>(px.off - origin * C);  // same addressing as the fast path
                        xv[u] = __ldg(reinterpret_cast<const float4*>(xs + px.off + cbase));
                        if (kRes) rv[u] = __ldg(reinterpret_cast<const float4*>(rs + px.off + cbase));
                      }
                    }
                  }
                }
                if (!waited) {
                  // the first loads of the chunk are in flight before the (rarely blocking) wait for the strip buffer
                  mbar_wait_sleep(&afree_bar[buf], ((g / static_cast<uint32_t>(p.hs_nbuf)) & 1u) ^ 1u);
                  waited = true;
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                  const int j = j0 + u;
                  if (j >= nunits) continue;
                  float o[4] = {0.f, 0.f, 0.f, 0.f};
                  if ((okm >> u) & 1u) {
                    o[0] = fa_act1(fmaf(xv[u].x, sc.x, sh.x), act); o[1] = fa_act1(fmaf(xv[u].y, sc.y, sh.y), act);
                    o[2] = fa_act1(fmaf(xv[u].z, sc.z, sh.z), act); o[3] = fa_act1(fmaf(xv[u].w, sc.w, sh.w), act);
                    if (kRes) { o[0] += rv[u].x; o[1] += rv[u].y; o[2] += rv[u].z; o[3] += rv[u].w; }
                    // write-back of the evaluated operand: the tile's own pixels only (each source pixel belongs to exactly
                    // one tile); `interior` excludes reflected border positions, which alias other pixels
                    if (wb && ((own_mask >> j) & 1u) && ((inm >> u) & 1u))
                      *reinterpret_cast<float4*>(p.fa_out[src] + origin * C + cbase + goff[u]) = make_float4(o[0], o[1], o[2], o[3]);
                  }
                  uint2 hi, lo;
                  fa_split4(o, p.fa_is_bf16, hi, lo);
                  *reinterpret_cast<uint2*>(strip_hi + soff[j]) = hi;
                  if (p.planes == 2) *reinterpret_cast<uint2*>(strip_lo + soff[j]) = lo;
                }
              }
            };
            if (rs != nullptr) convert(std::integral_constant<int, 6>{});
            else convert(std::integral_constant<int, 12>{});
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTA2) { if (crank == 0) mbar_arrive(&aready_bar[buf]); else mbar_arrive_remote(&aready_bar[buf], 0); }
              else mbar_arrive(&aready_bar[buf]);
            }
          }
        }
      }
    } else if (p.vt) {
      // ---- vertical strip fed from the TMA staging ring: smem (fp32, plain layout [row][8 px][64 ch]) -> transform ->
      // swizzled hi / lo strip.  A position outside the source is the conv's zero padding (the staged zeros must NOT go
      // through act(0 * scale + shift)), so validity is recomputed from the coordinates.
      uint32_t hn = 0;
      const int half_px = p.vt_half_rows * p.tile_w;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = decode_tile(p, t);
        const int sw0 = tc.w0 + p.tap_off[0][1] - p.fa_border, sh0 = tc.h0 + p.vs_dh_min - p.fa_border;
        for (int kc = 0; kc < kch0; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_hi = stage_base + static_cast<size_t>(s) * stage_bytes;
          uint8_t* a_lo = a_hi + vs_a_bytes;
          const int cbase = kc * kKC + q * 4;
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.fa_scale[0] != nullptr && tc.n0 < p.N) {
            sc = __ldg(reinterpret_cast<const float4*>(p.fa_scale[0] + static_cast<long long>(tc.n0) * p.fa_cin[0] + cbase));
            sh = __ldg(reinterpret_cast<const float4*>(p.fa_shift[0] + static_cast<long long>(tc.n0) * p.fa_cin[0] + cbase));
          }
          for (int half = 0; half < 2; ++half, ++hn) {
            const uint32_t slot = hn & 1u;
            mbar_wait(&sfull_bar[slot], (hn >> 1) & 1u);
            const uint8_t* st = vt_stage + slot * vt_slot_bytes;
            const int rows_here = min(p.vt_half_rows, p.vs_rows - half * p.vt_half_rows);
            const int npx_h = rows_here * p.tile_w;
            for (int i = pr; i < npx_h; i += 16) {
              const int row = half * p.vt_half_rows + (i >> 3), col = i & 7;
              const int hh = sh0 + row, ww = sw0 + col;
              float o[4] = {0.f, 0.f, 0.f, 0.f};
              if (tc.n0 < p.N && hh >= 0 && hh < p.Hs && ww >= 0 && ww < p.Ws) {
                const float4 v = *reinterpret_cast<const float4*>(st + static_cast<size_t>(i) * 256 + q * 16);
                o[0] = fa_act1(fmaf(v.x, sc.x, sh.x), p.fa_act[0]); o[1] = fa_act1(fmaf(v.y, sc.y, sh.y), p.fa_act[0]);
                o[2] = fa_act1(fmaf(v.z, sc.z, sh.z), p.fa_act[0]); o[3] = fa_act1(fmaf(v.w, sc.w, sh.w), p.fa_act[0]);
              }
              uint2 hi, lo;
              fa_split4(o, p.fa_is_bf16, hi, lo);
              const int idx = half * half_px + i;
              const uint32_t off = static_cast<uint32_t>(idx) * 128u + (static_cast<uint32_t>(chunk ^ (idx & 7)) << 4) + sub;
              *reinterpret_cast<uint2*>(a_hi + off) = hi;
              if (p.planes == 2) *reinterpret_cast<uint2*>(a_lo + off) = lo;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sempty_bar[slot]);      // this warp has read the slot
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    } else
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      if (p.vs) {
        const int npx = p.vs_rows * p.tile_w;                 // tile_w == 8: one image row = one 1024 B swizzle atom
        const int vw0 = tc.w0 + p.tap_off[0][1], vh0 = tc.h0 + p.vs_dh_min;
        for (int kc = 0; kc < kch0; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);                     // short stage cycles: poll (a suspended wait wakes up too coarsely)
          uint8_t* a_hi = stage_base + static_cast<size_t>(s) * stage_bytes;
          uint8_t* a_lo = a_hi + vs_a_bytes;
          const int cbase = kc * kKC + q * 4;
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.fa_scale[0] != nullptr && tc.n0 < p.N) {
            sc = __ldg(reinterpret_cast<const float4*>(p.fa_scale[0] + static_cast<long long>(tc.n0) * p.fa_cin[0] + cbase));
            sh = __ldg(reinterpret_cast<const float4*>(p.fa_shift[0] + static_cast<long long>(tc.n0) * p.fa_cin[0] + cbase));
          }
          // One latency wave per strip: every unit of this thread (a strip has at most 16 x 12 pixels) is loaded before any is
          // used — with waves of four the strip cost three global round trips and the converters, not the MMAs, set the pace.
          constexpr int kB = 12;
          const bool has_res = p.fa_res[0] != nullptr;
          for (int base = 0; base < npx; base += 16 * kB) {
            float4 xv[kB]; uint32_t okm = 0; long long offs[kB];
#pragma unroll
            for (int u = 0; u < kB; ++u) {
              const int idx = base + u * 16 + pr;
              xv[u] = make_float4(0.f, 0.f, 0.f, 0.f); offs[u] = 0;
              if (idx < npx) {
                const FaPix px = fa_locate(p, 0, tc.n0, vh0 + (idx >> 3), vw0 + (idx & 7));
                if (px.ok) {
                  okm |= 1u << u; offs[u] = px.off;
                  xv[u] = __ldg(reinterpret_cast<const float4*>(p.fa_x[0] + px.off + cbase));
                }
              }
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {
              const int idx = base + u * 16 + pr;
              if (idx >= npx) continue;
              float o[4] = {0.f, 0.f, 0.f, 0.f};
              if ((okm >> u) & 1u) {
                float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_res) rv = __ldg(reinterpret_cast<const float4*>(p.fa_res[0] + offs[u] + cbase));
                o[0] = fa_act1(fmaf(xv[u].x, sc.x, sh.x), p.fa_act[0]) + rv.x;
                o[1] = fa_act1(fmaf(xv[u].y, sc.y, sh.y), p.fa_act[0]) + rv.y;
                o[2] = fa_act1(fmaf(xv[u].z, sc.z, sh.z), p.fa_act[0]) + rv.z;
                o[3] = fa_act1(fmaf(xv[u].w, sc.w, sh.w), p.fa_act[0]) + rv.w;
              }
              uint2 hi, lo;
              fa_split4(o, p.fa_is_bf16, hi, lo);
              const uint32_t off = static_cast<uint32_t>(idx) * 128u + (static_cast<uint32_t>(chunk ^ (idx & 7)) << 4) + sub;
              *reinterpret_cast<uint2*>(a_hi + off) = hi;
              if (p.planes == 2) *reinterpret_cast<uint2*>(a_lo + off) = lo;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      } else {
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
          for (int src = 0; src < p.nsrc; ++src) {
            const bool wb = (p.fa_out[src] != nullptr) && (tap == p.fa_wb_tap) && (tc.cout0 == 0);
            for (int kc = 0; kc < p.kchunks[src]; ++kc) {
              mbar_wait_sleep(&empty_bar[s], ph ^ 1);
              uint8_t* a_hi = stage_base + static_cast<size_t>(s) * stage_bytes;
              uint8_t* a_lo = a_hi + kABytes;
              const int cbase = kc * kKC + q * 4;
              float4 xv[8], rv[8];
              FaPix px[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = pr + 16 * i;
                const int w_l = m & (p.tile_w - 1), h_l = (m >> lw) & (p.tile_h - 1), n_l = m >> (lw + lh);
                px[i] = fa_locate(p, src, tc.n0 + n_l, (tc.h0 + h_l) * p.conv_stride + dh, (tc.w0 + w_l) * p.conv_stride + dw);
                xv[i] = make_float4(0.f, 0.f, 0.f, 0.f); rv[i] = xv[i];
                if (px[i].ok) {
                  xv[i] = __ldg(reinterpret_cast<const float4*>(p.fa_x[src] + px[i].off + cbase));
                  if (p.fa_res[src] != nullptr) rv[i] = __ldg(reinterpret_cast<const float4*>(p.fa_res[src] + px[i].off + cbase));
                }
              }
              const bool has_ss = p.fa_scale[src] != nullptr;
              float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
              int sc_n = -1;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int m = pr + 16 * i;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
                if (px[i].ok) {
                  if (has_ss && px[i].n != sc_n) {
                    sc_n = px[i].n;
                    sc = __ldg(reinterpret_cast<const float4*>(p.fa_scale[src] + static_cast<long long>(sc_n) * p.fa_cin[src] + cbase));
                    sh = __ldg(reinterpret_cast<const float4*>(p.fa_shift[src] + static_cast<long long>(sc_n) * p.fa_cin[src] + cbase));
                  }
                  o[0] = fa_act1(fmaf(xv[i].x, sc.x, sh.x), p.fa_act[src]) + rv[i].x;
                  o[1] = fa_act1(fmaf(xv[i].y, sc.y, sh.y), p.fa_act[src]) + rv[i].y;
                  o[2] = fa_act1(fmaf(xv[i].z, sc.z, sh.z), p.fa_act[src]) + rv[i].z;
                  o[3] = fa_act1(fmaf(xv[i].w, sc.w, sh.w), p.fa_act[src]) + rv[i].w;
                  if (wb && px[i].interior)
                    *reinterpret_cast<float4*>(p.fa_out[src] + px[i].off + cbase) = make_float4(o[0], o[1], o[2], o[3]);
                }
                uint2 hi, lo;
                fa_split4(o, p.fa_is_bf16, hi, lo);
                const uint32_t off = static_cast<uint32_t>(m) * 128u + (static_cast<uint32_t>(chunk ^ (m & 7)) << 4) + sub;
                *reinterpret_cast<uint2*>(a_hi + off) = hi;
                if (p.planes == 2) *reinterpret_cast<uint2*>(a_lo + off) = lo;
              }
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(&full_bar[s]);
              if (++s == p.stages) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();       // nobody leaves while the peer may still arrive on / read from this CTA
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2sm(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// --------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------
int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// fp32 map: source staging (fused-operand vertical-strip mode, no swizzle) or an output tile store (128-byte swizzle).
bool encode_f32_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(fp32) failed with CUresult %d (rank %d)", (int)r, rank);
    set_error(buf);
    return false;
  }
  return true;
}

bool encode_tiled_map(CUtensorMap* map, const void* base, int is_bf16, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                  const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
    set_error(buf);
    return false;
  }
  return true;
}


// 128 output pixels per CTA tile = tile_n x tile_h x tile_w (shared with api.cu for the statistics slices).
// Returns 1 when the phase runs in vertical-strip mode (see TcParams::vs): R x 1 filter, stride 1, one source,
// all output channels in one N tile, resident weights + two A strips fit in shared memory.
// UMMA N when the caller leaves it open.  Large layers take the widest tile (A is read once per 256 output channels).
// Layers with at most a few hundred output pixels per phase (the inner UNet levels, the PatchGAN tail at small batch)
// are bound by the latency of streaming the weights through a handful of CTAs: there the tile narrows (down to 64)
// until the grid covers the SMs — four times the CTAs, each with a quarter of the weights and a deeper TMA ring.
static int auto_n_tile(int cout, long long out_px) {
  int n = cout >= 256 ? 256 : (cout >= 128 ? 128 : (cout > 32 ? 64 : 32));
  const long long m_tiles = (out_px + 127) / 128;
  while (n > 64 && m_tiles <= 4 && m_tiles * ((cout + n - 1) / n) < 148) n >>= 1;
  return n;
}

// Halo-strip eligibility (fused-operand mode only): stride 1, several taps, image at least one 16 x 8 tile, the strip of
// (16 + dh span) x (8 + dw span) pixels within kHsMaxPx, and room for the strip plus two weight stages.
static bool hs_eligible(const PhaseGeom& g, int split, int n_tile) {
  if (g.stride != 1 || g.ntaps < 1 || g.ntaps > kMaxTaps || g.OW < 8 || g.OH < 16) return false;
  int dh0 = g.tap_dh[0], dh1 = g.tap_dh[0], dw0 = g.tap_dw[0], dw1 = g.tap_dw[0];
  for (int t = 0; t < g.ntaps; ++t) {
    dh0 = min(dh0, g.tap_dh[t]); dh1 = max(dh1, g.tap_dh[t]); dw0 = min(dw0, g.tap_dw[t]); dw1 = max(dw1, g.tap_dw[t]);
  }
  const int rows = 16 + dh1 - dh0, cols = 8 + dw1 - dw0;
  if (rows * cols > kHsMaxPx) return false;
  const int planes = split ? 2 : 1;
  const long long strip = static_cast<long long>(planes) * ((rows * cols * 128 + 1023) / 1024 * 1024);
  return strip + 2LL * planes * n_tile * 128 + 1024 <= kMaxDynSmem;       // one strip buffer + two weight stages at least
}

// Would a halo-strip phase fit the TMA staging ring (two fp32 slots) next to its strips and two weight stages?
bool hs_staging_fits(const PhaseGeom& g, int split, int n_tile) {
  int dh0 = g.tap_dh[0], dh1 = dh0, dw0 = g.tap_dw[0], dw1 = dw0;
  for (int t = 0; t < g.ntaps; ++t) {
    dh0 = min(dh0, g.tap_dh[t]); dh1 = max(dh1, g.tap_dh[t]); dw0 = min(dw0, g.tap_dw[t]); dw1 = max(dw1, g.tap_dw[t]);
  }
  const int rows = 16 + dh1 - dh0, cols = 8 + dw1 - dw0, planes = split ? 2 : 1;
  const long long plane = (rows * cols * 128 + 1023) / 1024 * 1024, slot = (rows * cols * 256 + 1023) / 1024 * 1024;
  const long long nbuf = (2 * planes * plane + 2LL * planes * n_tile * 128 + 1024 <= kMaxDynSmem) ? 2 : 1;
  return nbuf * planes * plane + 2 * slot + 2LL * planes * n_tile * 128 + 1024 <= kMaxDynSmem;
}

int tc_plan_tiles(const PhaseGeom& g, int nsrc, const int* cin, int cout, int split, int n_tile_req, int* tile_w,
                  int* tile_h, int* tile_n, int* n_tile_out, int fa, int allow_hp) {
  int n_tile = n_tile_req;
  if (n_tile == 0) n_tile = auto_n_tile(cout, static_cast<long long>(g.N) * g.OH * g.OW);
  *n_tile_out = n_tile;
  int tw = g.OW >= 128 ? 128 : pow2_ceil(g.OW);
  int th = 128 / tw; { int hp = pow2_ceil(g.OH); if (th > hp) th = hp; }
  *tile_w = tw; *tile_h = th; *tile_n = 128 / (tw * th);
  // ---- vertical-strip eligibility: R x 1 filter, one source, all output channels in one N tile, resident weights ----
  bool vs = !(g.stride != 1 || nsrc != 1 || g.ntaps < 2 || g.ntaps > kMaxTaps || g.OW < 8 || g.OH < 16 || cout > n_tile);
  if (vs) {
    int dh_min = g.tap_dh[0], dh_max = g.tap_dh[0];
    for (int t = 0; t < g.ntaps; ++t) {
      if (g.tap_dw[t] != g.tap_dw[0]) vs = false;
      dh_min = min(dh_min, g.tap_dh[t]); dh_max = max(dh_max, g.tap_dh[t]);
    }
    const int planes = split ? 2 : 1;
    const int kch = cin[0] / kKC;
    const long long resident = static_cast<long long>(g.ntaps) * kch * planes * n_tile * 128;
    const long long strip = static_cast<long long>(planes) * (16 + dh_max - dh_min) * 8 * 128;
    if (resident + 2 * strip + 1024 > kMaxDynSmem) vs = false;
  }
  if (vs) { *tile_w = 8; *tile_h = 16; *tile_n = 1; return 1; }
  // ---- halo strip: any stride-1 tap set whose strip fits.  Fused operands: built by the converter warps; plane operands
  // (hp): loaded by TMA, worthwhile from two taps on (one tap = the strip is the tile) ----
  if ((fa || (allow_hp && g.ntaps >= 2)) && hs_eligible(g, split, n_tile)) { *tile_w = 8; *tile_h = 16; *tile_n = 1; return 2; }
  return 0;
}

// One phase of a convolution on the tensor cores.  See internal.h for the argument contract.
static void tc_plan_tiles_novs(const TcPhase& ph, int* tile_w, int* tile_h, int* tile_n, int* n_tile_out) {
  int n_tile = ph.n_tile;
  if (n_tile == 0) n_tile = auto_n_tile(ph.cout, static_cast<long long>(ph.N) * ph.OH * ph.OW);
  *n_tile_out = n_tile;
  int tw = ph.OW >= 128 ? 128 : pow2_ceil(ph.OW);
  int th = 128 / tw; { int hp = pow2_ceil(ph.OH); if (th > hp) th = hp; }
  *tile_w = tw; *tile_h = th; *tile_n = 128 / (tw * th);
}

int launch_conv_tc_phase(const TcPhase& ph, cudaStream_t stream) {
  int num_sms = 0;
  if (device_num_sms(&num_sms) != 0) return DLB_ERR_CUDA;
  if (ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_kernel_t<false>), kMaxDynSmem, kSlotConvTc) != 0) return DLB_ERR_CUDA;
  TcParams p;
  memset(&p, 0, sizeof(p));
  const int is_bf16 = (ph.fmt == DLB_FMT_BF16);
  p.planes = ph.split ? 2 : 1;
  p.nsrc = ph.nsrc;
  p.ntaps = ph.ntaps;
  if (ph.ntaps < 1 || ph.ntaps > kMaxTaps) return set_error("conv_tc: 1..16 taps per phase");
  int cin_total = 0;
  for (int s = 0; s < ph.nsrc; ++s) {
    if (ph.cin[s] % kKC != 0) return set_error("conv_tc: every source needs Cin % 64 == 0");
    p.kchunks[s] = ph.cin[s] / kKC;
    p.src_koff[s] = cin_total;
    cin_total += ph.cin[s];
  }
  if (ph.cout % 32 != 0) return set_error("conv_tc: Cout % 32 == 0 required");

  // ---- tile shape: 128 output pixels = tile_n x tile_h x tile_w ---------------------------------
  int tile_w, tile_h, tile_n, n_tile_plan;
  const int mode = ph.no_vs ? (tc_plan_tiles_novs(ph, &tile_w, &tile_h, &tile_n, &n_tile_plan), 0)
                           : tc_plan_tiles(ph, ph.nsrc, ph.cin, ph.cout, ph.split, ph.n_tile, &tile_w, &tile_h, &tile_n, &n_tile_plan, ph.fa,
                                           !ph.no_hs);
  const int use_vs = mode == 1, use_hs = mode == 2;
  // halo-strip extents (needed by the plane tensor maps below)
  int hs_dh0 = ph.tap_dh[0], hs_dh1 = hs_dh0, hs_dw0 = ph.tap_dw[0], hs_dw1 = hs_dw0;
  for (int t = 0; t < ph.ntaps; ++t) {
    hs_dh0 = min(hs_dh0, ph.tap_dh[t]); hs_dh1 = max(hs_dh1, ph.tap_dh[t]);
    hs_dw0 = min(hs_dw0, ph.tap_dw[t]); hs_dw1 = max(hs_dw1, ph.tap_dw[t]);
  }
  p.nacc = ph.nacc > 1 ? ph.nacc : 1;
  if (p.nacc > 1 && !use_hs) return set_error("conv_tc: merged phases need the halo-strip mode");
  for (int t = 0; t < ph.ntaps; ++t) p.tap_acc[t] = p.nacc > 1 ? ph.tap_acc[t] : 0;
  for (int a = 0; a < 4; ++a) { p.acc_ybase[a] = ph.acc_ybase[a]; p.acc_slice[a] = ph.acc_slice[a]; }
  p.tile_w = tile_w; p.tile_h = tile_h; p.tile_n = tile_n;
  p.tiles_w = (ph.OW + tile_w - 1) / tile_w;
  p.tiles_h = (ph.OH + tile_h - 1) / tile_h;
  p.tiles_n = (ph.N + tile_n - 1) / tile_n;

  int n_tile = n_tile_plan;
  if (n_tile != 32 && n_tile != 64 && n_tile != 128 && n_tile != 256) return set_error("conv_tc: n_tile must be 32/64/128/256");
  p.n_tile = n_tile;
  p.tiles_c = (ph.cout + n_tile - 1) / n_tile;
  p.N = ph.N; p.OH = ph.OH; p.OW = ph.OW; p.cout_total = ph.cout;
  p.ys_n = ph.ys_n; p.ys_h = ph.ys_h; p.ys_w = ph.ys_w; p.y_base = ph.y_base;
  p.y = ph.y; p.bias = ph.bias;
  if (2 * p.nacc * n_tile > 512) return set_error("conv_tc: merged phases need 2 * nacc * n_tile <= 512 TMEM columns");
  // CTA pair (cta_group::2) for the wide fused-operand halo-strip convolutions (the 256 -> 256 trunk): each SM then loads,
  // stores and feeds the tensor core only half of every weight tile.  DLB_CTA2=0 switches it off (A/B measurements).
  static const bool cta2_enabled = []() { const char* e = getenv("DLB_CTA2"); return !(e != nullptr && e[0] == '0'); }();
  const bool use_cta2 = cta2_enabled && use_hs && ph.fa == 1 && ph.nsrc == 1 && n_tile == 256 && ph.cout == 256 && p.nacc == 1 &&
                        !ph.no_cta2 && (num_sms % 2 == 0);
  p.cta2 = use_cta2 ? 1 : 0;
  p.idesc = make_idesc_f16(use_cta2 ? 256 : 128, n_tile, is_bf16);
  p.st_partial = ph.st_partial; p.st_cnt = ph.st_cnt; p.st_S = ph.st_S;
  p.st_S_cap = ph.st_S_cap; p.st_slice_base = ph.st_slice_base; p.st_S_total = ph.st_S_total;
  if (ph.st_partial != nullptr && tile_n != 1) return set_error("conv_tc: fused statistics need H*W >= 128 per image");

  // ---- activation tensor maps ---------------------------------------------------------------------
  for (int s = 0; s < ph.nsrc; ++s) {
    const uint64_t C = ph.cin[s], W = ph.W, H = ph.H, N = ph.N;
    uint64_t dims[5], strides[4]; uint32_t box[5];
    if (ph.stride == 1) {
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = N; dims[4] = 1;
      strides[0] = C * 2; strides[1] = W * C * 2; strides[2] = H * W * C * 2; strides[3] = N * H * W * C * 2;
      box[0] = kKC; box[1] = tile_w; box[2] = tile_h; box[3] = tile_n; box[4] = 1;
      if (use_vs) {
        int dmin = ph.tap_dh[0], dmax = ph.tap_dh[0];
        for (int t = 0; t < ph.ntaps; ++t) { dmin = min(dmin, ph.tap_dh[t]); dmax = max(dmax, ph.tap_dh[t]); }
        p.vs = 1; p.vs_dh_min = dmin; p.vs_rows = tile_h + dmax - dmin;
        for (int t = 0; t < ph.ntaps; ++t) p.vs_row_off[t] = ph.tap_dh[t] - dmin;
        box[2] = p.vs_rows;
      }
      if (use_hs && !ph.fa) { box[1] = tile_w + hs_dw1 - hs_dw0; box[2] = tile_h + hs_dh1 - hs_dh0; box[3] = 1; }
      p.dim_sel[0] = 0; p.dim_sel[1] = 1; p.dim_sel[2] = 2; p.dim_sel[3] = 3; p.dim_sel[4] = 4;
    } else {
      if ((W & 1) || (H & 1)) return set_error("conv_tc: stride-2 needs even H and W");
      // x[n, 2*hh+hp, 2*ww+wp, c] viewed as (wp*C + c, ww, hp, hh, n)
      dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
      strides[0] = 2 * C * 2; strides[1] = W * C * 2; strides[2] = 2 * W * C * 2; strides[3] = H * W * C * 2;
      box[0] = kKC; box[1] = tile_w; box[2] = 1; box[3] = tile_h; box[4] = tile_n;
      p.dim_sel[0] = 0; p.dim_sel[1] = 1; p.dim_sel[2] = 4; p.dim_sel[3] = 2; p.dim_sel[4] = 3;
    }
    if (ph.fa) continue;                       // fused-operand mode: no activation planes, no tensor maps
    if (!encode_tiled_map(&p.a_hi[s], ph.x_hi[s], is_bf16, 5, dims, strides, box)) return -1;
    if (ph.split && !encode_tiled_map(&p.a_lo[s], ph.x_lo[s], is_bf16, 5, dims, strides, box)) return -1;
  }
  if (ph.fa) {
    p.fa = ph.fa; p.fa_is_bf16 = is_bf16; p.fa_wb_tap = -1;
    p.stem_C = ph.stem_C; p.stem_S = ph.stem_S; p.stem_pad = ph.stem_pad;
    p.fa_border = ph.fa_border; p.fa_border_mode = ph.fa_border_mode;
    p.H = ph.H; p.W = ph.W; p.Hs = ph.H - 2 * ph.fa_border; p.Ws = ph.W - 2 * ph.fa_border; p.conv_stride = ph.stride;
    if (ph.fa == 2) {
      // stem source: H = Hs + 2 * pad rows of the window operand, W = Ws columns (the horizontal padding lives in the lanes)
      p.Hs = ph.H - 2 * ph.stem_pad; p.Ws = ph.W;
      if (!use_vs) return set_error("conv_tc: the fused stem needs the vertical-strip mode (image at least 16 x 8, S x 1 filter)");
      if (ph.stem_C < 1 || ph.stem_C > 4 || ph.stem_S < 1 || ph.stem_S > 8 || ph.cin[0] != 64 || ph.nsrc != 1)
        return set_error("conv_tc: fused stem needs C <= 4, S <= 8 and a 64-lane operand");
      if (tile_h + ph.stem_S - 1 > 24) return set_error("conv_tc: fused stem patch rows exceed 24");
    }
    if (p.Hs < 1 || p.Ws < 1) return set_error("conv_tc: fused operand border larger than the input");
    if (ph.fa_border_mode == DLB_PAD_REFLECT && (ph.fa_border >= p.Hs || ph.fa_border >= p.Ws))
      return set_error("conv_tc: reflect border must be smaller than the source");
    bool any_out = false;
    for (int s = 0; s < ph.nsrc; ++s) {
      p.fa_x[s] = ph.fa_x[s]; p.fa_scale[s] = ph.fa_scale[s]; p.fa_shift[s] = ph.fa_shift[s]; p.fa_res[s] = ph.fa_res[s];
      p.fa_out[s] = ph.fa_out[s]; p.fa_act[s] = ph.fa_act[s]; p.fa_cin[s] = ph.cin[s];
      if (ph.fa_x[s] == nullptr) return set_error("conv_tc: fused operand source is null");
      if ((ph.fa_scale[s] == nullptr) != (ph.fa_shift[s] == nullptr)) return set_error("conv_tc: scale and shift come together");
      any_out = any_out || ph.fa_out[s] != nullptr;
    }
    for (int t = 0; t < ph.ntaps; ++t) {
      p.tap_dh[t] = ph.tap_dh[t]; p.tap_dw[t] = ph.tap_dw[t];
      // the tap that reads input pixel (oh + b, ow + b) for output (oh, ow) visits every source pixel exactly once
      if (ph.stride == 1 && ph.tap_dh[t] == ph.fa_border && ph.tap_dw[t] == ph.fa_border) p.fa_wb_tap = t;
    }
    if (any_out && (use_vs || p.fa_wb_tap < 0 || ph.OH != p.Hs || ph.OW != p.Ws))
      return set_error("conv_tc: operand write-back needs a stride-1 'same' convolution in tap / halo-strip mode");
    if (use_hs) {
      const int dh0 = hs_dh0, dh1 = hs_dh1, dw0 = hs_dw0, dw1 = hs_dw1;
      p.hs = 1; p.hs_rows = tile_h + dh1 - dh0; p.hs_cols = tile_w + dw1 - dw0; p.hs_dh_min = dh0; p.hs_dw_min = dw0;
      p.hs_plane_bytes = (p.hs_rows * p.hs_cols * 128 + 1023) / 1024 * 1024;
      // two strip buffers when they still leave room for two weight stages (the 256 -> 256 trunk conv: 2 x 46 KB + 2 x 64 KB)
      p.hs_nbuf = (2LL * p.planes * p.hs_plane_bytes + 2LL * p.planes * (p.cta2 ? n_tile / 2 : n_tile) * 128 + 1024 <= kMaxDynSmem) ? 2 : 1;
      // TMA staging of the fp32 source (ht) when it is a plain single source and two slots fit next to the strips and two
      // weight stages
      const int slot = (p.hs_rows * p.hs_cols * 256 + 1023) / 1024 * 1024;
      if (ph.nsrc == 1 && ph.fa_res[0] == nullptr && ph.fa_out[0] == nullptr &&
          (ph.fa_border == 0 || ph.fa_border_mode == DLB_PAD_ZERO) &&
          static_cast<long long>(p.hs_nbuf) * p.planes * p.hs_plane_bytes + 2LL * slot + 2LL * p.planes * n_tile * 128 + 1024 <= kMaxDynSmem) {
        const uint64_t C = ph.cin[0];
        uint64_t dims[4] = {C, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)ph.N};
        uint64_t strides[3] = {C * 4, (uint64_t)p.Ws * C * 4, (uint64_t)p.Hs * p.Ws * C * 4};
        uint32_t box[4] = {(uint32_t)kKC, (uint32_t)p.hs_cols, (uint32_t)p.hs_rows, 1};
        if (!encode_f32_map(&p.a_f32, ph.fa_x[0], 4, dims, strides, box)) return -1;
        p.ht = 1; p.ht_slot_bytes = slot;
      }
      for (int t = 0; t < ph.ntaps; ++t) p.hs_off[t] = ((ph.tap_dh[t] - dh0) * p.hs_cols + (ph.tap_dw[t] - dw0)) * 128;
    }
  }
  const bool epi2_wanted = []() { const char* e = getenv("DLB_EPI2"); return e != nullptr && e[0] == '1'; }();
  if (use_hs && !ph.fa) {
    // plane-fed halo strips: the strip of a chunk is one TMA box per plane (tensor maps above)
    p.hs = 1; p.hp = 1;
    p.hs_rows = tile_h + hs_dh1 - hs_dh0; p.hs_cols = tile_w + hs_dw1 - hs_dw0; p.hs_dh_min = hs_dh0; p.hs_dw_min = hs_dw0;
    p.hs_plane_bytes = (p.hs_rows * p.hs_cols * 128 + 1023) / 1024 * 1024;
    // strip buffers: as many as fit (up to 4) while at least 4 weight stages remain — a strip load has ~1-2 us of TMA latency
    // and a narrow layer spends no more than that on a chunk's MMAs, so two buffers are not enough to stay ahead
    p.hs_nbuf = 1;
    for (int nb = 2; nb <= 4; ++nb) {
      const long long need = static_cast<long long>(nb) * p.planes * p.hs_plane_bytes +
                             (nb == 2 ? 2LL : 4LL) * p.planes * n_tile * 128 + 1024;
      if (need + (epi2_wanted ? 2 : 1) * 16384 + 1024 <= kMaxDynSmem) p.hs_nbuf = nb;   // room for the epilogue's transpose buffer(s) (below)
    }
    for (int t = 0; t < ph.ntaps; ++t) p.hs_off[t] = ((ph.tap_dh[t] - hs_dh0) * p.hs_cols + (ph.tap_dw[t] - hs_dw0)) * 128;
  }
  // ---- weight tensor map: [taps_total][Cout][Cin_total] ---------------------------------------------
  {
    uint64_t dims[3] = {(uint64_t)cin_total, (uint64_t)ph.cout, (uint64_t)ph.w_taps};
    uint64_t strides[2] = {(uint64_t)cin_total * 2, (uint64_t)cin_total * ph.cout * 2};
    uint32_t box[3] = {(uint32_t)kKC, (uint32_t)(p.cta2 ? n_tile / 2 : n_tile), 1};   // cta2: each CTA loads half the rows
    if (!encode_tiled_map(&p.b_hi, ph.w_hi, is_bf16, 3, dims, strides, box)) return -1;
    if (ph.split && !encode_tiled_map(&p.b_lo, ph.w_lo, is_bf16, 3, dims, strides, box)) return -1;
  }
  // ---- taps ---------------------------------------------------------------------------------------
  for (int t = 0; t < ph.ntaps; ++t) {
    const int dh = ph.tap_dh[t], dw = ph.tap_dw[t];
    p.tap_w[t] = ph.tap_widx[t];
    if (ph.stride == 1) {
      p.tap_off[t][0] = 0; p.tap_off[t][1] = dw; p.tap_off[t][2] = dh; p.tap_off[t][3] = 0; p.tap_off[t][4] = 0;
    } else {
      // input row = 2*i + dh  ->  hp = dh mod 2, hh = i + floor(dh / 2)
      const int hp = ((dh % 2) + 2) % 2, wp = ((dw % 2) + 2) % 2;
      const int dhh = (dh - hp) / 2, dww = (dw - wp) / 2;
      // channel offset wp*C only valid for a single source per distinct C; each source has its own map
      p.tap_off[t][0] = 0;  // patched per source below through tap_wp
      p.tap_off[t][1] = dww; p.tap_off[t][2] = hp; p.tap_off[t][3] = dhh; p.tap_off[t][4] = 0;
      if (ph.nsrc != 1) return set_error("conv_tc: stride-2 supports one source");
      p.tap_off[t][0] = wp * ph.cin[0];
    }
  }

  const int b_bytes = (p.cta2 ? n_tile / 2 : n_tile) * 128;
  int bres_bytes = use_vs ? ph.ntaps * p.kchunks[0] * p.planes * b_bytes
                          : (use_hs ? p.hs_nbuf * p.planes * p.hs_plane_bytes : 0);
  if (p.ht) bres_bytes += 2 * p.ht_slot_bytes;               // the staging ring sits between the strips and the weight stages
  // vertical strip + plain fused source (no residual, zero border): feed the converters from a TMA staging ring
  if (use_vs && ph.fa == 1 && ph.nsrc == 1 && ph.fa_res[0] == nullptr && ph.fa_out[0] == nullptr &&
      (ph.fa_border == 0 || ph.fa_border_mode == DLB_PAD_ZERO) && ph.cin[0] % kKC == 0) {
    const int half_rows = (p.vs_rows + 1) / 2;
    const int slot = half_rows * tile_w * 256;
    const int strip = p.planes * p.vs_rows * tile_w * 128;
    if (bres_bytes + 2 * slot + 2 * strip + 1024 <= kMaxDynSmem) {
      const uint64_t C = ph.cin[0];
      uint64_t dims[4] = {C, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)ph.N};
      uint64_t strides[3] = {C * 4, (uint64_t)p.Ws * C * 4, (uint64_t)p.Hs * p.Ws * C * 4};
      uint32_t box[4] = {(uint32_t)kKC, (uint32_t)tile_w, (uint32_t)half_rows, 1};
      if (!encode_f32_map(&p.a_f32, ph.fa_x[0], 4, dims, strides, box)) return -1;
      p.vt = 1; p.vt_half_rows = half_rows;
      bres_bytes += 2 * slot;                                  // the staging ring sits behind the resident weights
    }
  }
  const int stage_bytes = use_hs ? p.planes * b_bytes
                                 : (use_vs ? p.planes * p.vs_rows * tile_w * 128 + (ph.fa == 2 ? kStemPatchBytes : 0)
                                           : p.planes * (kABytes + b_bytes));
  // epilogue through shared memory (statistics from the buffer): plane-fed launches copy the buffer out with coalesced stores
  // (17 KB); fused-operand launches store it by TMA (33 KB), leaving the LSU to the converters.  Only where the buffers can be
  // spared without making the pipeline shallower than 3 stages.  DLB_EPI_SMEM=0 switches it off, =1 plane-fed launches only.
  static const int ets_env = []() { const char* e = getenv("DLB_EPI_SMEM"); return e != nullptr ? atoi(e) : 2; }();
  int ets_mode = (ets_env >= 1 && !ph.fa) ? 1 : ((ets_env >= 2 && ph.fa) ? 2 : 0);
  if ((ph.ys_w % 4) || (ph.ys_h % 4) || (ph.ys_n % 4) || (ph.cout % 32) || (p.nacc <= 1 && ph.y_base % 4)) ets_mode = 0;
  for (int a = 0; a < p.nacc && p.nacc > 1; ++a) if (ph.acc_ybase[a] % 4) ets_mode = 0;
  // the copy-out keeps 32-bit element offsets per thread
  if (ets_mode == 1 && (static_cast<long long>(ph.N) + 1) * ph.ys_n + static_cast<long long>(ph.OH + 32) * ph.ys_h >= (1LL << 31)) ets_mode = 0;
  // second epilogue group: measured +0.4 % on the headline step and on UNet-256 (the narrow layers are bound by the MMA issue
  // rate, ~60 cycles per N = 64 MMA, not by their epilogue), so it is off unless DLB_EPI2=1
  static const bool epi2_env = []() { const char* e = getenv("DLB_EPI2"); return e != nullptr && e[0] == '1'; }();
  p.epi2 = (epi2_env && !ph.fa && p.nacc * (n_tile / 32) >= 2) ? 1 : 0;
  const int ets_need = (ets_mode == 2 || p.epi2) ? 2 * 16384 + 1024 : 16384 + 1024;
  {
    const int cap = ph.max_stages > 0 ? ph.max_stages : kMaxStages;
    int s_without = (kMaxDynSmem - 1024 - bres_bytes) / stage_bytes; if (s_without > cap) s_without = cap;
    int s_with = (kMaxDynSmem - 1024 - bres_bytes - ets_need) / stage_bytes; if (s_with > cap) s_with = cap;
    if (s_with < 2 || (s_with < 3 && s_with < s_without)) ets_mode = 0;
  }
  int stages = (kMaxDynSmem - 1024 - bres_bytes - (ets_mode ? ets_need : 0)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (ph.max_stages > 0 && stages > ph.max_stages) stages = ph.max_stages;
  if (stages < 2) return set_error("conv_tc: not enough shared memory for 2 stages");
  p.stages = stages;
  p.ets = ets_mode;
  p.ets_off = ets_mode ? (bres_bytes + stages * stage_bytes + 1023) / 1024 * 1024 : 0;
  if (ets_mode == 2) {
    for (int a = 0; a < p.nacc; ++a) {
      const long long base = p.nacc > 1 ? ph.acc_ybase[a] : ph.y_base;
      const uint64_t dims[4] = {(uint64_t)ph.cout, (uint64_t)ph.OW, (uint64_t)ph.OH, (uint64_t)ph.N};
      const uint64_t strides[3] = {(uint64_t)ph.ys_w * 4, (uint64_t)ph.ys_h * 4, (uint64_t)ph.ys_n * 4};
      const uint32_t box[4] = {32, (uint32_t)tile_w, (uint32_t)tile_h, (uint32_t)tile_n};
      if (!encode_f32_map(&p.y_map[a], ph.y + base, 4, dims, strides, box, 1)) return -1;
    }
  }
  const int smem_bytes = (ets_mode ? p.ets_off + ((ets_mode == 2 || p.epi2) ? 2 : 1) * 16384 : bres_bytes + stages * stage_bytes) + 1024;
  const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.tiles_c;
  int grid = total_tiles < num_sms ? total_tiles : num_sms;
  if (ph.max_ctas > 0 && grid > ph.max_ctas) grid = ph.max_ctas;
  if (p.cta2) {
    // clusters of two CTAs (consecutive blockIdx.x = consecutive tiles); an even grid, at most one CTA per SM
    grid = (grid + 1) & ~1;
    if (grid > num_sms) grid = num_sms & ~1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(kThreadsFa, 1, 1); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_kernel_t<true>), kMaxDynSmem, kSlotConvTc2) != 0) return DLB_ERR_CUDA;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel_t<true>, p);
    if (e != cudaSuccess) {
      char buf[200];
      snprintf(buf, sizeof(buf), "conv_tc_kernel cluster launch (grid %d, smem %d): %s", grid, smem_bytes, cudaGetErrorString(e));
      cudaGetLastError();
      set_error(buf);
      return DLB_ERR_CUDA;
    }
    return 0;
  }
  const int threads = ph.fa ? kThreadsFa : (p.epi2 ? kThreads + 128 : kThreads);
  conv_tc_kernel_t<false><<<grid, threads, smem_bytes, stream>>>(p);
  {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      char buf[200];
      snprintf(buf, sizeof(buf), "conv_tc_kernel launch (grid %d, threads %d, smem %d): %s", grid, threads,
               smem_bytes, cudaGetErrorString(e));
      set_error(buf);
      return DLB_ERR_CUDA;
    }
  }
  return 0;
}

}  // namespace dlb

// Row-streaming tensor-core convolution for the generator head: ReflectionPad2d(3) / ZeroPad2d(3) + Conv2d(64, CO <= 3, 7)
// + bias + Tanh (reference networks.py:438-444), reading the producer's RAW fp32 NHWC output and evaluating its
// normalisation + activation while loading, writing the network output fp32 NCHW.  One kernel replaces
// dlb_norm_apply / dlb_conv_tc_fwd_fused (vertical-strip mode) + dlb_head_finish for this layer.
//
// Formulation.  out[r, c, co] = sum_{kh, kw, ci} a[pad(r + kh - 3), pad(c + kw - 3), ci] * w[co, ci, kh, kw] is split as
//     z[r', c, (kh, co)] = sum_{kw, ci} a[r', pad(c + kw - 3), ci] * w[co, ci, kh, kw]          (tensor cores)
//     out[r, c, co]      = sum_kh z[pad(r + kh - 3), c, (kh, co)]                               (epilogue registers)
// i.e. the horizontal taps are K (7 x 64) and the vertical taps are 7 x 4 = 28 -> 32 output columns of the MMA.  A CTA
// walks DOWN a 128-pixel-wide column strip one input row at a time:
//   producer   (warp 0)      one 1-D bulk copy per input row: 134 px x 64 ch fp32 (contiguous in NHWC) into a 3-deep ring
//   converters (warps 6-13)  act(x * scale + shift) -> bf16 hi / lo planes, 128-byte-swizzled rows (one row per pixel), the
//                            zero / reflected COLUMN border resolved here
//   MMA        (warp 1)      per row 56 x tcgen05.mma M128 K16: tap kw reads the A rows shifted by kw pixels (descriptor
//                            start + kw * 128 B: the SW128 pattern is a function of the absolute address, ptx.cuh); split
//                            precision as a_hi x [w_hi | w_lo] (N = 64) + a_lo x w_hi (N = 32), weights resident in shared
//                            memory; the warp runs convergent and one elected lane issues (uniform-register descriptors)
//   epilogue   (warps 2-5)   thread = one pixel column: a 7-row sliding window of output sums in registers; input row r'
//                            adds z[kh] to window row r' - kh + 3 (plus the reflected ROW border terms), the finished row
//                            gets bias + tanh and a coalesced NCHW store.
// Work split: all N x strips x H strip rows form one sequence cut into gridDim.x contiguous ranges (a range may span two
// strips); each range recomputes z for the 3 input rows above and below it.
// Algorithmic bytes: 256 B read per input pixel + 4 * CO written per output pixel; the z tensor never exists in HBM.
#include <cuda_bf16.h>

#include "internal.h"
#include "ptx.cuh"

namespace dlb {
namespace {

constexpr int kHcThreads = 448;                      // 14 warps
constexpr int kHcTW = 128;                           // strip width = MMA M
constexpr int kHcS = 7, kHcHalo = 3;
constexpr int kHcPx = kHcTW + 2 * kHcHalo;           // 134 operand rows (pixels) per input row
constexpr uint32_t kHcPlaneBytes = 136 * 128;        // 17 groups of 8 rows x 128 B
constexpr uint32_t kHcStageBytes = kHcPx * 256;      // fp32 staging slot
constexpr int kHcNS = 3;                             // staging ring
constexpr int kHcNB = 2;                             // operand plane ring
constexpr uint32_t kHcWTile = 64 * 128;              // one kw weight tile: rows 0-31 hi, 32-63 lo of (kh, co), x 64 ci
constexpr uint32_t kHcWBytes = kHcS * kHcWTile;      // 57344
constexpr uint32_t kHcOffA = kHcWBytes;
constexpr uint32_t kHcOffS = kHcOffA + kHcNB * 2 * kHcPlaneBytes;
constexpr uint32_t kHcOffBar = kHcOffS + kHcNS * kHcStageBytes;
constexpr uint32_t kHcSmem = kHcOffBar + 256 + 1024;  // + barriers + alignment slack

struct HeadParams {
  const float* x; const float* scale; const float* shift; const uint8_t* wpk; const float* bias; float* y;
  int N, H, W, CO, act, out_act, border_mode, strips;
  long long rows_total;
};

struct Piece { int n, c0, ra, rb, ia, ib; };

__device__ __forceinline__ bool next_piece(const HeadParams& p, long long& cur, long long end, Piece& pc) {
  if (cur >= end) return false;
  const long long col = cur / p.H;
  const int r0 = static_cast<int>(cur - col * p.H);
  const long long left = end - cur;
  const int len = left < (p.H - r0) ? static_cast<int>(left) : (p.H - r0);
  pc.n = static_cast<int>(col / p.strips);
  pc.c0 = static_cast<int>(col % p.strips) * kHcTW;
  pc.ra = r0; pc.rb = r0 + len;
  pc.ia = r0 - kHcHalo < 0 ? 0 : r0 - kHcHalo;
  pc.ib = pc.rb + kHcHalo > p.H ? p.H : pc.rb + kHcHalo;
  cur += len;
  return true;
}

__device__ __forceinline__ float hc_act(float v, int act) {
  if (act == DLB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLB_ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
  if (act == DLB_ACT_TANH) return tanhf(v);
  return v;
}

// 4 fp32 -> 4 bf16 hi + 4 bf16 lo (hi = rn(v), lo = rn(v - hi)): the arithmetic of norm_apply_kernel
__device__ __forceinline__ void hc_split4(const float (&o)[4], uint2& hi, uint2& lo) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { h[k] = __float2bfloat16_rn(o[k]); l[k] = __float2bfloat16_rn(o[k] - __bfloat162float(h[k])); }
  hi = *reinterpret_cast<uint2*>(h); lo = *reinterpret_cast<uint2*>(l);
}

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint2 v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// One horizontal tap (K = 64: four K = 16 steps) from one asm block: a_hi x [w_hi | w_lo] (N = 64) then a_lo x w_hi (N = 32, the
// first 32 rows of the same weight tile) per step; the descriptors cross to uniform registers once and step by 32 bytes there.
__device__ __forceinline__ void hc_mma_tap(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t w, uint32_t idesc64, uint32_t idesc32,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e, t;\n\t.reg .b64 ah, al, wb;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, 0, 0;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "mov.b64 ah, %1;\n\tmov.b64 al, %2;\n\tmov.b64 wb, %3;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, wb, %4, p;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, wb, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 wb, wb, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, wb, %4, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, wb, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 wb, wb, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, wb, %4, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, wb, %5, t;\n\t"
      "add.s64 ah, ah, 2;\n\tadd.s64 al, al, 2;\n\tadd.s64 wb, wb, 2;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah, wb, %4, t;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al, wb, %5, t;\n\t"
      "}"
      :
      : "r"(tmem_d), "l"(a_hi), "l"(a_lo), "l"(w), "r"(idesc64), "r"(idesc32), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kHcThreads, 1) head_conv_kernel(const HeadParams p) {
  extern __shared__ uint8_t hc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sW = smem;
  uint8_t* sA = smem + kHcOffA;
  uint8_t* sS = smem + kHcOffS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHcOffBar);
  uint64_t* wbar = bars;                 // weights landed
  uint64_t* sfull = bars + 1;            // [kHcNS] staging slot filled (bulk copy bytes)
  uint64_t* sempty = bars + 4;           // [kHcNS] staging slot consumed (8 converter warps)
  uint64_t* aready = bars + 7;           // [kHcNB] operand planes written (8 converter warps)
  uint64_t* afree = bars + 9;            // [kHcNB] MMAs that read the planes retired
  uint64_t* tfull = bars + 11;           // [2] accumulator complete
  uint64_t* tempty = bars + 13;          // [2] accumulator drained (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long T = p.rows_total;
  const long long beg = T * blockIdx.x / gridDim.x, end = T * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(wbar, 1);
    for (int i = 0; i < kHcNS; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], 8); }
    for (int i = 0; i < kHcNB; ++i) { mbar_init(&aready[i], 8); mbar_init(&afree[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: weights once, then one bulk copy per input row =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(wbar, kHcWBytes);
      bulk_copy_g2s(sW, p.wpk, kHcWBytes, wbar);
      uint32_t it = 0;
      long long cur = beg; Piece pc;
      while (next_piece(p, cur, end, pc)) {
        const int clo = pc.c0 - kHcHalo < 0 ? 0 : pc.c0 - kHcHalo;
        const int chi = pc.c0 + kHcTW + kHcHalo > p.W ? p.W : pc.c0 + kHcTW + kHcHalo;
        const uint32_t bytes = static_cast<uint32_t>(chi - clo) * 256u;
        for (int i = pc.ia; i < pc.ib; ++i, ++it) {
          const uint32_t st = it % kHcNS;
          mbar_wait(&sempty[st], ((it / kHcNS) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&sfull[st], bytes);
          bulk_copy_g2s(sS + st * kHcStageBytes + static_cast<uint32_t>(clo - (pc.c0 - kHcHalo)) * 256u,
                        p.x + ((static_cast<long long>(pc.n) * p.H + i) * p.W + clo) * 64, bytes, &sfull[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      mbar_wait(wbar, 0);
      const uint32_t idesc64 = make_idesc_f16(128, 64, 1), idesc32 = make_idesc_f16(128, 32, 1);
      const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(sW));
      uint32_t it = 0;
      long long cur = beg; Piece pc;
      while (next_piece(p, cur, end, pc)) {
        for (int i = pc.ia; i < pc.ib; ++i, ++it) {
          const uint32_t b = it % kHcNB, t = it & 1u;
          mbar_wait(&aready[b], (it / kHcNB) & 1u);
          mbar_wait(&tempty[t], ((it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          // descriptors differ from the row's base only in the start-address field (units of 16 B): plain 64-bit adds
          const uint64_t da_hi = make_sw128_kmajor_desc_sbo(smem_u32(sA + b * 2 * kHcPlaneBytes), 1024);
          const uint64_t da_lo = da_hi + (kHcPlaneBytes >> 4);
          const uint32_t d_tmem = tmem_base + t * 64u;
#pragma unroll
          for (int kw = 0; kw < kHcS; ++kw)
            // a_hi x [w_hi | w_lo] -> columns 0..63; a_lo x w_hi -> columns 0..31 (the epilogue adds the two halves)
            hc_mma_tap(d_tmem, da_hi + ((kw * 128) >> 4), da_lo + ((kw * 128) >> 4), db0 + ((kw * kHcWTile) >> 4), idesc64, idesc32,
                       kw != 0);
          umma_commit_elect(&afree[b]);
          umma_commit_elect(&tfull[t]);
        }
      }
    }
  } else if (warp < 6) {
    // ===================== epilogue: vertical taps in a register window, bias + act, NCHW store =====================
    const int q = warp & 3;
    const int px = q * 32 + lane;
    float bias[3] = {0.f, 0.f, 0.f};
    for (int co = 0; co < 3; ++co) if (co < p.CO && p.bias != nullptr) bias[co] = __ldg(p.bias + co);
    const bool refl = p.border_mode == DLB_PAD_REFLECT;
    const long long plane = static_cast<long long>(p.H) * p.W;
    uint32_t it = 0;
    long long cur = beg; Piece pc;
    while (next_piece(p, cur, end, pc)) {
      float win[7][3];
#pragma unroll
      for (int k = 0; k < 7; ++k) { win[k][0] = 0.f; win[k][1] = 0.f; win[k][2] = 0.f; }
      const int col = pc.c0 + px;
      const bool colok = col < p.W;
      float* const yb = p.y + static_cast<long long>(pc.n) * p.CO * plane + col;
      auto emit = [&](int r, const float (&a)[3]) {
        if (r >= pc.ra && r < pc.rb && colok) {
#pragma unroll
          for (int co = 0; co < 3; ++co)
            if (co < p.CO) yb[co * plane + static_cast<long long>(r) * p.W] = hc_act(a[co] + bias[co], p.out_act);
        }
      };
      for (int i = pc.ia; i < pc.ib; ++i, ++it) {
        const uint32_t t = it & 1u;
        mbar_wait_sleep(&tfull[t], (it >> 1) & 1u);
        tc_fence_after();
        uint32_t v[32], v2[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64u, v);
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 64u + 32u, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[t]);
#define HC_Z(kh, co) (__uint_as_float(v[(kh) * 4 + (co)]) + __uint_as_float(v2[(kh) * 4 + (co)]))
        // input row i feeds output row i - kh + 3 = window row 6 - kh
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
          for (int co = 0; co < 3; ++co) win[6 - kh][co] += HC_Z(kh, co);
        }
        if (refl) {
          // reflected row border: input row i is also pad(r + kh - 3) for r + kh - 3 = -i (top) or 2(H-1) - i (bottom)
          if (i == 1) {
#pragma unroll
            for (int co = 0; co < 3; ++co) { win[4][co] += HC_Z(0, co); win[3][co] += HC_Z(1, co); win[2][co] += HC_Z(2, co); }
          } else if (i == 2) {
#pragma unroll
            for (int co = 0; co < 3; ++co) { win[2][co] += HC_Z(0, co); win[1][co] += HC_Z(1, co); }
          } else if (i == 3) {
#pragma unroll
            for (int co = 0; co < 3; ++co) win[0][co] += HC_Z(0, co);
          }
          const int d = p.H - 1 - i;
          if (d == 1) {
#pragma unroll
            for (int co = 0; co < 3; ++co) { win[4][co] += HC_Z(4, co); win[3][co] += HC_Z(5, co); win[2][co] += HC_Z(6, co); }
          } else if (d == 2) {
#pragma unroll
            for (int co = 0; co < 3; ++co) { win[5][co] += HC_Z(5, co); win[4][co] += HC_Z(6, co); }
          } else if (d == 3) {
#pragma unroll
            for (int co = 0; co < 3; ++co) win[6][co] += HC_Z(6, co);
          }
        }
#undef HC_Z
        emit(i - 3, win[0]);
        if (i == pc.ib - 1) { emit(i - 2, win[1]); emit(i - 1, win[2]); emit(i, win[3]); }
#pragma unroll
        for (int k = 0; k < 6; ++k) { win[k][0] = win[k + 1][0]; win[k][1] = win[k + 1][1]; win[k][2] = win[k + 1][2]; }
        win[6][0] = 0.f; win[6][1] = 0.f; win[6][2] = 0.f;
      }
    }
  } else {
    // ===================== converters: fp32 staging -> act(x*scale + shift) -> swizzled bf16 hi / lo rows =====================
    // lane = (qq: one of 4 pixels, g: channels 4g..4g+3 and 32+4g..32+4g+3).  A warp instruction covers the pixels
    // {0,1,4,5} or {2,3,6,7} of an 8-row group: loads are 4 x 128 contiguous bytes, the 8-byte stores of the two rows
    // that share a swizzle half land on disjoint banks.
    const int cw = warp - 6;
    const int qq = lane >> 3, g = lane & 7;
    const int rsub = (qq & 1) + 4 * (qq >> 1);
    uint32_t it = 0;
    long long cur = beg; Piece pc;
    while (next_piece(p, cur, end, pc)) {
      float sc[8], sh[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int ch = (k < 4 ? 4 * g + k : 32 + 4 * g + (k - 4));
        sc[k] = p.scale != nullptr ? __ldg(p.scale + pc.n * 64 + ch) : 1.f;
        sh[k] = p.shift != nullptr ? __ldg(p.shift + pc.n * 64 + ch) : 0.f;
      }
      const int cbase = pc.c0 - kHcHalo;
      const bool refl = p.border_mode == DLB_PAD_REFLECT;
      for (int i = pc.ia; i < pc.ib; ++i, ++it) {
        const uint32_t st = it % kHcNS, b = it % kHcNB;
        mbar_wait(&sfull[st], (it / kHcNS) & 1u);
        mbar_wait(&afree[b], ((it / kHcNB) & 1u) ^ 1u);
        const uint32_t s_base = smem_u32(sS + st * kHcStageBytes);
        const uint32_t a_hi = smem_u32(sA + b * 2 * kHcPlaneBytes);
        const uint32_t a_lo = a_hi + kHcPlaneBytes;
        for (int u = cw; u < 34; u += 8) {
          const int j = (u >> 1) * 8 + (u & 1) * 2 + rsub;
          if (j < kHcPx) {
            const int cj = cbase + j;
            int cr = cj;
            bool inside = cj >= 0 && cj < p.W;
            if (refl && !inside && cj >= -kHcHalo && cj < p.W + kHcHalo) { cr = cj < 0 ? -cj : 2 * p.W - 2 - cj; inside = true; }
            float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
            if (inside) {
              const uint32_t sp = s_base + static_cast<uint32_t>(cr - cbase) * 256u + static_cast<uint32_t>(g) * 16u;
              const float4 v0 = ld_shared_f4(sp), v1 = ld_shared_f4(sp + 128u);
              o0[0] = hc_act(fmaf(v0.x, sc[0], sh[0]), p.act); o0[1] = hc_act(fmaf(v0.y, sc[1], sh[1]), p.act);
              o0[2] = hc_act(fmaf(v0.z, sc[2], sh[2]), p.act); o0[3] = hc_act(fmaf(v0.w, sc[3], sh[3]), p.act);
              o1[0] = hc_act(fmaf(v1.x, sc[4], sh[4]), p.act); o1[1] = hc_act(fmaf(v1.y, sc[5], sh[5]), p.act);
              o1[2] = hc_act(fmaf(v1.z, sc[6], sh[6]), p.act); o1[3] = hc_act(fmaf(v1.w, sc[7], sh[7]), p.act);
            }
            uint2 h0, l0, h1, l1;
            hc_split4(o0, h0, l0); hc_split4(o1, h1, l1);
            const uint32_t row = static_cast<uint32_t>(j) * 128u, sw = static_cast<uint32_t>(j & 7);
            const uint32_t off0 = row + (((static_cast<uint32_t>(g) >> 1) ^ sw) << 4) + (static_cast<uint32_t>(g) & 1u) * 8u;
            const uint32_t off1 = row + (((4u + (static_cast<uint32_t>(g) >> 1)) ^ sw) << 4) + (static_cast<uint32_t>(g) & 1u) * 8u;
            st_shared_v2(a_hi + off0, h0); st_shared_v2(a_hi + off1, h1);
            st_shared_v2(a_lo + off0, l0); st_shared_v2(a_lo + off1, l1);
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) { mbar_arrive(&aready[b]); mbar_arrive(&sempty[st]); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// w fp32 [CO][64][7][7] -> [kw][row = plane * 32 + kh*4 + co (64)][ci 64] bf16 (plane 0 = hi, 1 = lo), rows of 128 B with
// the 128-byte swizzle (16-byte chunk c of row r stored at chunk c ^ (r & 7)); rows with co >= CO or kh >= 7 are zero.
__global__ void head_pack_kernel(const float* __restrict__ w, int CO, __nv_bfloat16* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kHcS * 64 * 64) return;
  const int ci = e & 63, prow = (e >> 6) & 63, kw = e >> 12;
  const int plane = prow >> 5, row = prow & 31;
  const int kh = row >> 2, co = row & 3;
  float v = 0.f;
  if (kh < kHcS && co < CO) v = w[((co * 64 + ci) * kHcS + kh) * kHcS + kw];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 val = plane == 0 ? hi : __float2bfloat16_rn(v - __bfloat162float(hi));
  const int chunk = (ci >> 3) ^ (prow & 7);
  out[(static_cast<size_t>(kw) * 64 + prow) * 64 + chunk * 8 + (ci & 7)] = val;
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" size_t dlb_head_conv_weight_bytes(void) { return kHcWBytes; }

extern "C" int dlb_head_conv_pack_weights(const float* w, int CO, int C, int R, int S, void* out, dlb_stream_t stream) {
  if (C != 64 || R != kHcS || S != kHcS || CO < 1 || CO > 3) return set_error("dlb_head_conv_pack_weights: needs a [CO<=3][64][7][7] filter");
  head_pack_kernel<<<(kHcS * 64 * 64 + 255) / 256, 256, 0, stream>>>(w, CO, reinterpret_cast<__nv_bfloat16*>(out));
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("head_pack_kernel launch");
  return 0;
}

extern "C" int dlb_head_conv_fwd(const float* x, const float* scale, const float* shift, int act, int N, int H, int W, int C,
                                 const void* w_packed, const float* bias, int CO, int border_mode, int out_act, float* y_nchw,
                                 dlb_stream_t stream) {
  if (C != 64 || CO < 1 || CO > 3) return set_error("dlb_head_conv_fwd: needs C == 64 and CO <= 3");
  if (H < 8 || W < 8 || N < 1) return set_error("dlb_head_conv_fwd: needs H, W >= 8");
  if (border_mode != DLB_PAD_ZERO && border_mode != DLB_PAD_REFLECT) return set_error("dlb_head_conv_fwd: bad border mode");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w_packed) & 15))
    return set_error("dlb_head_conv_fwd: x and w_packed must be 16-byte aligned");
  int sms = 0;
  if (int rc = device_num_sms(&sms)) return rc;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(head_conv_kernel), static_cast<int>(kHcSmem), kSlotHeadConv)) return rc;
  HeadParams p;
  p.x = x; p.scale = scale; p.shift = shift; p.wpk = static_cast<const uint8_t*>(w_packed); p.bias = bias; p.y = y_nchw;
  p.N = N; p.H = H; p.W = W; p.CO = CO; p.act = act; p.out_act = out_act; p.border_mode = border_mode;
  p.strips = (W + kHcTW - 1) / kHcTW;
  p.rows_total = static_cast<long long>(N) * p.strips * H;
  long long grid = p.rows_total / 8;                 // at least ~8 strip rows per CTA (each range adds up to 6 halo rows)
  if (grid < 1) grid = 1;
  if (grid > sms) grid = sms;
  head_conv_kernel<<<static_cast<unsigned>(grid), kHcThreads, kHcSmem, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("head_conv_kernel launch");
  return 0;
}

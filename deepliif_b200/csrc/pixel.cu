// Byte/integer ends of the path: the uint8 <-> fp32 tile conversions and the fused segmentation finish.
//   dlb_u8_to_f32   deepliif.data.transform (data/__init__.py:133-138): ToTensor (/255) + Normalize(0.5, 0.5)
//   dlb_f32_to_u8   util.tensor2im (util/util.py:130-135): trunc((x + 1) / 2 * 255) in fp32
//   dlb_seg_finish  run_dask aggregation (models/__init__.py:338) + tensor2im + create_posneg_mask
//                   (postprocessing.py:163-190, labels :87-95)
// Pure HBM streams; bit-exact against oracle/pixel.py (same fp32 operation order, no FMA contraction).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace dlb {
namespace {

__device__ __forceinline__ uint8_t quant_u8(float x) {
  // numpy: (x + 1) / 2.0 * 255.0 in float32, then astype(uint8) (truncation toward zero)
  float v = __fmul_rn(__fdiv_rn(__fadd_rn(x, 1.0f), 2.0f), 255.0f);
  int i = static_cast<int>(v);            // trunc
  return static_cast<uint8_t>(i);         // numpy wraps out-of-range modulo 256 on x86; inputs are tanh-bounded
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ img, float* __restrict__ out, int N, int H, int W) {
  const long long total = static_cast<long long>(N) * H * W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = idx / (static_cast<long long>(H) * W);
    const long long hw = idx % (static_cast<long long>(H) * W);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn(static_cast<float>(img[idx * 3 + c]), 255.0f);
      out[(n * 3 + c) * H * W + hw] = __fdiv_rn(__fsub_rn(v, 0.5f), 0.5f);
    }
  }
}

__global__ void f32_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int N, int H, int W) {
  const long long total = static_cast<long long>(N) * H * W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = idx / (static_cast<long long>(H) * W);
    const long long hw = idx % (static_cast<long long>(H) * W);
#pragma unroll
    for (int c = 0; c < 3; ++c) out[idx * 3 + c] = quant_u8(x[(n * 3 + c) * H * W + hw]);
  }
}

struct SegParams {
  const float* segs[8];
  float w[8];
  int nseg, N, H, W, thresh;
  float* seg_f32; uint8_t* seg_u8; uint8_t* mask;
};

__global__ void seg_finish_kernel(const SegParams p) {
  const long long plane = static_cast<long long>(p.H) * p.W;
  const long long total = static_cast<long long>(p.N) * plane;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = idx / plane, hw = idx % plane;
    int u[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const long long off = (n * 3 + c) * plane + hw;
      float acc = 0.f;
      for (int k = 0; k < p.nseg; ++k) acc = __fadd_rn(acc, __fmul_rn(p.segs[k][off], p.w[k]));
      if (p.seg_f32) p.seg_f32[off] = acc;
      const uint8_t q = quant_u8(acc);
      u[c] = q;
      if (p.seg_u8) p.seg_u8[idx * 3 + c] = q;
    }
    if (p.mask) {
      uint8_t m = 50;                                   // LABEL_UNKNOWN
      if (u[0] + u[2] > p.thresh && u[1] <= 80) m = (u[0] >= u[2]) ? 200 : 150;   // POSITIVE : NEGATIVE
      p.mask[idx] = m;
    }
  }
}

// Head finish: y[n, co, h, w] = act(bias[co] + sum_s z[n, h, w + s, s*4 + co]).
// z is the output of the "horizontal taps as output channels" convolution (see dlb_head_finish in the header):
// fp32 NHWC [N, H, W + S - 1, 32]; thread = one output pixel, S float4 loads.
__global__ void __launch_bounds__(256) head_finish_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                                                          int N, int H, int W, int S, int CO, int act,
                                                          float* __restrict__ y) {
  const int WZ = W + S - 1;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(N) * plane;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % W);
    const int h = static_cast<int>((idx / W) % H);
    const int n = static_cast<int>(idx / plane);
    const float* zr = z + ((static_cast<long long>(n) * H + h) * WZ + w) * 32;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int s = 0; s < S; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(zr + s * 32 + s * 4));
      a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    const float acc[4] = {a0, a1, a2, a3};
    for (int co = 0; co < CO; ++co) {
      float v = acc[co] + (bias ? __ldg(bias + co) : 0.f);
      if (act == DLB_ACT_TANH) v = tanhf(v);
      else if (act == DLB_ACT_RELU) v = fmaxf(v, 0.f);
      y[(static_cast<long long>(n) * CO + co) * plane + static_cast<long long>(h) * W + w] = v;
    }
  }
}

// Backward of head_finish: dz[n, h, u, s*4 + co] = dzz[n, co, h, u - s] (0 <= u - s < W), other lanes 0; written as
// 64-channel hi/lo planes [N, H, W+S-1, 64] = the operand of the head's dgrad / wgrad GEMMs.  S = 1, CO = 1 is the
// PatchGAN last conv, S = 1, CO = 3 the UNet outermost ConvTranspose.  thread = one (n, h, u) pixel.
template <typename T16>
__global__ void __launch_bounds__(256) head_bwd_pack_kernel(const float* __restrict__ dzz, int N, int H, int W, int S,
                                                            int CO, T16* __restrict__ out_hi, T16* __restrict__ out_lo) {
  const int WZ = W + S - 1;
  const long long plane = static_cast<long long>(H) * W;
  const long long total = static_cast<long long>(N) * H * WZ;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int u = static_cast<int>(idx % WZ);
    const int h = static_cast<int>((idx / WZ) % H);
    const int n = static_cast<int>(idx / (static_cast<long long>(WZ) * H));
    __align__(16) T16 hi[64];
    __align__(16) T16 lo[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) { hi[k] = T16(0.f); lo[k] = T16(0.f); }
    for (int s = 0; s < S; ++s) {
      const int w = u - s;
      if (w < 0 || w >= W) continue;
      for (int co = 0; co < CO; ++co) {
        const float v = __ldg(dzz + (static_cast<long long>(n) * CO + co) * plane + static_cast<long long>(h) * W + w);
        const T16 hh = T16(v);
        hi[s * 4 + co] = hh;
        lo[s * 4 + co] = T16(v - static_cast<float>(hh));
      }
    }
    uint4* dh = reinterpret_cast<uint4*>(out_hi + idx * 64);
    const uint4* sh = reinterpret_cast<const uint4*>(hi);
#pragma unroll
    for (int k = 0; k < 8; ++k) dh[k] = sh[k];
    if (out_lo != nullptr) {
      uint4* dl = reinterpret_cast<uint4*>(out_lo + idx * 64);
      const uint4* sl = reinterpret_cast<const uint4*>(lo);
#pragma unroll
      for (int k = 0; k < 8; ++k) dl[k] = sl[k];
    }
  }
}

// is_empty() support (deepliif/models/__init__.py:391-396, util/__init__.py:478-485): per tile, over the pixels whose
// PIL 'L' luma L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 is neither 0 nor 255 (the reference drops saturated black /
// white before taking the variance): their count, sum of L and sum of L^2 as exact 64-bit integers (order-independent,
// so the atomics are deterministic); the host turns them into the variance.  grid (blocks_per_tile, N), block 256.
__global__ void __launch_bounds__(256) tile_luma_sums_kernel(const uint8_t* __restrict__ img, int HW,
                                                             unsigned long long* __restrict__ sums) {
  const int n = blockIdx.y;
  unsigned long long s0 = 0, s1 = 0, s2 = 0;
  for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < HW; px += gridDim.x * blockDim.x) {
    const uint8_t* q = img + (static_cast<long long>(n) * HW + px) * 3;
    const unsigned int L = (19595u * q[0] + 38470u * q[1] + 7471u * q[2] + 0x8000u) >> 16;
    if (L != 0u && L != 255u) { s0 += 1; s1 += L; s2 += L * L; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sums[n * 3], s0); atomicAdd(&sums[n * 3 + 1], s1); atomicAdd(&sums[n * 3 + 2], s2);
  }
}

int grid1d(long long total) {
  long long g = (total + 255) / 256;
  return static_cast<int>(g < 148 * 8 ? (g < 1 ? 1 : g) : 148 * 8);
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" int dlb_u8_to_f32(const uint8_t* img_nhwc, float* out_nchw, int N, int H, int W, dlb_stream_t stream) {
  u8_to_f32_kernel<<<grid1d(static_cast<long long>(N) * H * W), 256, 0, stream>>>(img_nhwc, out_nchw, N, H, W);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("u8_to_f32_kernel launch");
  return 0;
}

extern "C" int dlb_f32_to_u8(const float* x_nchw, uint8_t* out_nhwc, int N, int H, int W, dlb_stream_t stream) {
  f32_to_u8_kernel<<<grid1d(static_cast<long long>(N) * H * W), 256, 0, stream>>>(x_nchw, out_nhwc, N, H, W);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("f32_to_u8_kernel launch");
  return 0;
}

extern "C" int dlb_seg_finish(const float* const* segs, const float* weights, int nseg, int N, int H, int W, int thresh,
                              float* seg_f32_nchw, uint8_t* seg_u8_nhwc, uint8_t* mask, dlb_stream_t stream) {
  if (nseg < 1 || nseg > 8) return set_error("dlb_seg_finish: 1..8 seg inputs");
  SegParams p;
  memset(&p, 0, sizeof(p));
  for (int k = 0; k < nseg; ++k) { p.segs[k] = segs[k]; p.w[k] = weights[k]; }
  p.nseg = nseg; p.N = N; p.H = H; p.W = W; p.thresh = thresh;
  p.seg_f32 = seg_f32_nchw; p.seg_u8 = seg_u8_nhwc; p.mask = mask;
  seg_finish_kernel<<<grid1d(static_cast<long long>(N) * H * W), 256, 0, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("seg_finish_kernel launch");
  return 0;
}

extern "C" int dlb_head_finish(const float* z, const float* bias, int N, int H, int W, int S, int CO, int act,
                               float* y_nchw, dlb_stream_t stream) {
  if (S < 1 || S > 8 || CO < 1 || CO > 4) return set_error("dlb_head_finish: needs S <= 8 and CO <= 4");
  head_finish_kernel<<<grid1d(static_cast<long long>(N) * H * W), 256, 0, stream>>>(z, bias, N, H, W, S, CO, act, y_nchw);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("head_finish_kernel launch");
  return 0;
}

extern "C" int dlb_head_bwd_pack(const float* dzz_nchw, int N, int H, int W, int S, int CO, int fmt, void* out_hi,
                                 void* out_lo, dlb_stream_t stream) {
  if (S < 1 || S > 8 || CO < 1 || CO > 4) return set_error("dlb_head_bwd_pack: needs S <= 8 and CO <= 4");
  const int g = grid1d(static_cast<long long>(N) * H * (W + S - 1));
  if (fmt == DLB_FMT_BF16)
    head_bwd_pack_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(dzz_nchw, N, H, W, S, CO, reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                             reinterpret_cast<__nv_bfloat16*>(out_lo));
  else if (fmt == DLB_FMT_FP16)
    head_bwd_pack_kernel<__half><<<g, 256, 0, stream>>>(dzz_nchw, N, H, W, S, CO, reinterpret_cast<__half*>(out_hi),
                                                      reinterpret_cast<__half*>(out_lo));
  else return set_error("dlb_head_bwd_pack: bad fmt");
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("head_bwd_pack_kernel launch");
  return 0;
}

// sums: uint64 [N][3] (count, sum L, sum L^2 over the pixels with 0 < L < 255), zeroed by this call.
extern "C" int dlb_tile_luma_sums(const uint8_t* img_nhwc, int N, int H, int W, unsigned long long* sums,
                                  dlb_stream_t stream) {
  if (cudaMemsetAsync(sums, 0, sizeof(unsigned long long) * 3 * N, stream) != cudaSuccess) return set_cuda_error("memset");
  int bx = (H * W + 255) / 256; if (bx > 64) bx = 64;
  tile_luma_sums_kernel<<<dim3(bx, N), 256, 0, stream>>>(img_nhwc, H * W, sums);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("tile_luma_sums_kernel launch");
  return 0;
}

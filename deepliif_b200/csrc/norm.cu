// Normalisation statistics and the fused "normalise + activation (+ residual) + operand split" pass.
//
// Replaces nn.BatchNorm2d (batch statistics, affine) / nn.InstanceNorm2d (affine-free) + nn.ReLU /
// nn.LeakyReLU + the ResnetBlock skip add (networks.py:25-44, 391-404, 490-513, 573-606, 640-656).
// A per-plane reduction sits between every conv and its activation, so the conv epilogue cannot apply
// the norm itself (SURVEY.md §7 hard part 2); these passes are pure HBM streams: 128-bit loads/stores,
// one read of the raw conv output and one write of the operand planes of the next conv.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"

namespace dlb {
namespace {

constexpr int kSlicePixels = 128;   // pixels per partial-statistics slice

// ---- pass 1: per-slice (sum, M2) partials; thread = 4 channels, coalesced float4 rows -------------
// grid (slices, N), block 256.  y: [N][HW][C].  partial: [N][slices][C] float2 (sum, M2 about slice mean).
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ y, int HW, int C,
                                                            int slices, float2* __restrict__ partial) {
  __shared__ float s_sum[256 * 4];
  __shared__ float s_m2[256 * 4];
  __shared__ int s_cnt[256];
  const int n = blockIdx.y, sl = blockIdx.x;
  const int c4n = C / 4;                           // float4 columns
  const int p0 = sl * kSlicePixels;
  const int p1 = min(p0 + kSlicePixels, HW);
  const int tid = threadIdx.x;
  // threads are laid out as (pixel lane, channel quad): lanes = 256 / min(c4n,256)
  for (int cq0 = 0; cq0 < c4n; cq0 += 256) {
    const int cols = min(c4n - cq0, 256);
    const int lanes = 256 / cols;                  // C is a multiple of 4 and cols divides 256 for C in {4..1024}
    const int cq = cq0 + tid % cols;
    const int pl = tid / cols;
    float sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0}, shiftv[4] = {0, 0, 0, 0};
    bool have_shift = false;
    int cnt = 0;
    if (pl < lanes) {
      for (int px = p0 + pl; px < p1; px += lanes) {
        const float4 v = *reinterpret_cast<const float4*>(y + (static_cast<long long>(n) * HW + px) * C + cq * 4);
        const float a[4] = {v.x, v.y, v.z, v.w};
        if (!have_shift) { for (int k = 0; k < 4; ++k) shiftv[k] = a[k]; have_shift = true; }
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float d = a[k] - shiftv[k]; sum[k] += d; sq[k] = fmaf(d, d, sq[k]); }
        ++cnt;
      }
    }
    // per-thread (count, mean, M2) then fixed-order merge over the pixel lanes through shared memory
    float mean[4], m2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float inv = cnt > 0 ? 1.f / cnt : 0.f;
      const float ds = sum[k] * inv;
      mean[k] = shiftv[k] + ds;
      m2[k] = fmaxf(sq[k] - sum[k] * ds, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_sum[tid * 4 + k] = mean[k]; s_m2[tid * 4 + k] = m2[k]; }
    s_cnt[tid] = (pl < lanes) ? cnt : 0;
    __syncthreads();
    if (pl == 0) {
      // Chan merge in lane order (deterministic)
      float cn = 0.f, cm[4] = {0, 0, 0, 0}, cM[4] = {0, 0, 0, 0};
      for (int l = 0; l < lanes; ++l) {
        const int o = l * cols + (tid % cols);
        const float nb = static_cast<float>(s_cnt[o]);
        if (nb == 0.f) continue;
        const float nt = cn + nb;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = s_sum[o * 4 + k] - cm[k];
          cm[k] += d * (nb / nt);
          cM[k] += s_m2[o * 4 + k] + d * d * (cn * nb / nt);
        }
        cn = nt;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        partial[(static_cast<long long>(n) * slices + sl) * C + cq * 4 + k] = make_float2(cm[k] * cn, cM[k]);
    }
    __syncthreads();
  }
}

// ---- pass 2: merge slices (fp64, fixed order) -> scale/shift --------------------------------------
// thread = one (n, c) (or one c when pooled).
__global__ void stats_finalize_kernel(const float2* __restrict__ partial, int N, int HW, int C, int slices,
                                      int pooled, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float eps, float* __restrict__ scale, float* __restrict__ shift) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int groups = pooled ? 1 : N;
  if (idx >= groups * C) return;
  const int c = idx % C, g = idx / C;
  const int n_lo = pooled ? 0 : g, n_hi = pooled ? N : g + 1;
  double cn = 0.0, cm = 0.0, cM = 0.0;
  for (int n = n_lo; n < n_hi; ++n) {
    for (int s = 0; s < slices; ++s) {
      const float2 pr = partial[(static_cast<long long>(n) * slices + s) * C + c];
      const int p0 = s * kSlicePixels;
      const double nb = static_cast<double>(min(kSlicePixels, HW - p0));
      const double mb = static_cast<double>(pr.x) / nb;
      const double nt = cn + nb;
      const double d = mb - cm;
      cm += d * (nb / nt);
      cM += static_cast<double>(pr.y) + d * d * (cn * nb / nt);
      cn = nt;
    }
  }
  const double var = cM / cn;                        // biased variance (PyTorch norm layers)
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float ga = gamma ? gamma[c] : 1.f;
  const float be = beta ? beta[c] : 0.f;
  const float sc = ga * rstd;
  const float sh = be - static_cast<float>(cm) * sc;
  for (int n = n_lo; n < n_hi; ++n) { scale[n * C + c] = sc; shift[n * C + c] = sh; }
}

// ---- apply: act(y*scale+shift) (+residual) -> fp32 and/or split 16-bit planes ---------------------
template <typename T16> struct Cvt;
template <> struct Cvt<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 to(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float from(__nv_bfloat16 v) { return __bfloat162float(v); }
};
template <> struct Cvt<__half> {
  static __device__ __forceinline__ __half to(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float from(__half v) { return __half2float(v); }
};

__device__ __forceinline__ float act1(float v, int act) {
  if (act == DLB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLB_ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
  if (act == DLB_ACT_TANH) return tanhf(v);
  return v;
}

struct ApplyParams {
  const float* y; const float* scale; const float* shift; int act; const float* residual;
  float* out_f32; void* out_hi; void* out_lo;
  int N, H, W, C, pad, pad_mode;
};

// One thread = one output pixel-quad (4 channels) of the (possibly padded) output grid.
template <typename T16>
__global__ void __launch_bounds__(256) norm_apply_kernel(const ApplyParams p) {
  const int c4n = p.C / 4;
  const int HP = p.H + 2 * p.pad, WP = p.W + 2 * p.pad;
  const long long total = static_cast<long long>(p.N) * HP * WP * c4n;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(idx % c4n);
    long long r = idx / c4n;
    const int wp = static_cast<int>(r % WP); r /= WP;
    const int hp = static_cast<int>(r % HP);
    const int n = static_cast<int>(r / HP);
    int h = hp - p.pad, w = wp - p.pad;
    bool border = (h < 0) || (h >= p.H) || (w < 0) || (w >= p.W);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    bool zero = false;
    if (border) {
      if (p.pad_mode == DLB_PAD_REFLECT) {
        if (h < 0) h = -h; if (h >= p.H) h = 2 * p.H - 2 - h;
        if (w < 0) w = -w; if (w >= p.W) w = 2 * p.W - 2 - w;
      } else {
        zero = true;
      }
    }
    if (!zero) {
      const long long src = ((static_cast<long long>(n) * p.H + h) * p.W + w) * p.C + cq * 4;
      const float4 v = *reinterpret_cast<const float4*>(p.y + src);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      if (p.scale != nullptr) {
        const float4 sc = *reinterpret_cast<const float4*>(p.scale + n * p.C + cq * 4);
        const float4 sh = *reinterpret_cast<const float4*>(p.shift + n * p.C + cq * 4);
        o[0] = fmaf(o[0], sc.x, sh.x); o[1] = fmaf(o[1], sc.y, sh.y);
        o[2] = fmaf(o[2], sc.z, sh.z); o[3] = fmaf(o[3], sc.w, sh.w);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = act1(o[k], p.act);
      if (p.residual != nullptr) {
        const float4 rv = *reinterpret_cast<const float4*>(p.residual + src);
        o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
      }
      if (p.out_f32 != nullptr && !border)
        *reinterpret_cast<float4*>(p.out_f32 + src) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (p.out_hi != nullptr) {
      const long long dst = ((static_cast<long long>(n) * HP + hp) * WP + wp) * p.C + cq * 4;
      T16 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hi[k] = Cvt<T16>::to(o[k]);
        lo[k] = Cvt<T16>::to(o[k] - Cvt<T16>::from(hi[k]));
      }
      *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(p.out_hi) + dst) = *reinterpret_cast<uint2*>(hi);
      if (p.out_lo != nullptr)
        *reinterpret_cast<uint2*>(reinterpret_cast<T16*>(p.out_lo) + dst) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

// ---- weight repacking -------------------------------------------------------------------------------
// tc: [tap][Cout][Cin] 16-bit hi/lo;  direct: [tap][Cin][Cout] fp32.
template <typename T16>
__global__ void pack_w_tc_kernel(const float* __restrict__ w, int cout, int cin, int taps, int transposed,
                                 T16* __restrict__ hi, T16* __restrict__ lo) {
  const long long total = static_cast<long long>(taps) * cout * cin;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ci = static_cast<int>(idx % cin);
    const int co = static_cast<int>((idx / cin) % cout);
    const int t = static_cast<int>(idx / (static_cast<long long>(cin) * cout));
    const long long src = transposed ? (static_cast<long long>(ci) * cout + co) * taps + t
                                     : (static_cast<long long>(co) * cin + ci) * taps + t;
    const float v = w[src];
    const T16 h = Cvt<T16>::to(v);
    hi[idx] = h;
    if (lo != nullptr) lo[idx] = Cvt<T16>::to(v - Cvt<T16>::from(h));
  }
}

__global__ void pack_w_direct_kernel(const float* __restrict__ w, int cout, int cin, int taps, int transposed,
                                     float* __restrict__ out) {
  const long long total = static_cast<long long>(taps) * cout * cin;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % cout);
    const int ci = static_cast<int>((idx / cout) % cin);
    const int t = static_cast<int>(idx / (static_cast<long long>(cin) * cout));
    const long long src = transposed ? (static_cast<long long>(ci) * cout + co) * taps + t
                                     : (static_cast<long long>(co) * cin + ci) * taps + t;
    out[idx] = w[src];
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" size_t dlb_norm_stats_workspace(int N, int HW, int C) {
  const int slices = (HW + kSlicePixels - 1) / kSlicePixels;
  return static_cast<size_t>(N) * slices * C * sizeof(float2);
}

extern "C" int dlb_norm_stats(const float* y, int N, int HW, int C, int pooled, const float* gamma, const float* beta,
                              float eps, float* scale, float* shift, void* workspace, size_t workspace_bytes,
                              dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_stats: C % 4 != 0");
  const int c4n = C / 4;
  if (c4n < 256 && 256 % c4n != 0) return set_error("dlb_norm_stats: C/4 must divide 256 (or be a multiple of 256)");
  if (c4n > 256 && c4n % 256 != 0) return set_error("dlb_norm_stats: C/4 must divide 256 (or be a multiple of 256)");
  if (workspace_bytes < dlb_norm_stats_workspace(N, HW, C)) return set_error("dlb_norm_stats: workspace too small");
  const int slices = (HW + kSlicePixels - 1) / kSlicePixels;
  float2* partial = reinterpret_cast<float2*>(workspace);
  stats_partial_kernel<<<dim3(slices, N), 256, 0, stream>>>(y, HW, C, slices, partial);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stats_partial_kernel launch");
  const int groups = pooled ? 1 : N;
  stats_finalize_kernel<<<(groups * C + 127) / 128, 128, 0, stream>>>(partial, N, HW, C, slices, pooled, gamma, beta,
                                                                    eps, scale, shift);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stats_finalize_kernel launch");
  return 0;
}

extern "C" int dlb_norm_apply(const float* y, const float* scale, const float* shift, int act, const float* residual,
                              float* out_f32, void* out_hi, void* out_lo, int fmt, int N, int H, int W, int C, int pad,
                              int pad_mode, dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_apply: C % 4 != 0");
  if (pad < 0 || (pad_mode == DLB_PAD_REFLECT && (pad >= H || pad >= W))) return set_error("dlb_norm_apply: bad pad");
  if (out_hi == nullptr && out_f32 == nullptr) return set_error("dlb_norm_apply: no output");
  ApplyParams p{y, scale, shift, act, residual, out_f32, out_hi, out_lo, N, H, W, C, pad, pad_mode};
  const long long total = static_cast<long long>(N) * (H + 2 * pad) * (W + 2 * pad) * (C / 4);
  const int grid = grid_for(total, 256);
  if (fmt == DLB_FMT_BF16) norm_apply_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  else if (fmt == DLB_FMT_FP16) norm_apply_kernel<__half><<<grid, 256, 0, stream>>>(p);
  else return set_error("dlb_norm_apply: bad fmt");
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("norm_apply_kernel launch");
  return 0;
}

extern "C" int dlb_pack_weights_tc(const dlb_conv_desc* d, const float* w, int fmt, void* w_hi, void* w_lo,
                                   dlb_stream_t stream) {
  int cin = 0;
  for (int s = 0; s < d->nsrc; ++s) cin += d->Cin[s];
  const int taps = d->R * d->S;
  const long long total = static_cast<long long>(taps) * d->Cout * cin;
  const int grid = grid_for(total, 256);
  if (fmt == DLB_FMT_BF16)
    pack_w_tc_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w, d->Cout, cin, taps, d->transposed,
                                                            reinterpret_cast<__nv_bfloat16*>(w_hi),
                                                            reinterpret_cast<__nv_bfloat16*>(w_lo));
  else if (fmt == DLB_FMT_FP16)
    pack_w_tc_kernel<__half><<<grid, 256, 0, stream>>>(w, d->Cout, cin, taps, d->transposed,
                                                     reinterpret_cast<__half*>(w_hi), reinterpret_cast<__half*>(w_lo));
  else return set_error("dlb_pack_weights_tc: bad fmt");
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("pack_w_tc_kernel launch");
  return 0;
}

extern "C" int dlb_pack_weights_direct(const dlb_conv_desc* d, const float* w, float* w_packed, dlb_stream_t stream) {
  int cin = 0;
  for (int s = 0; s < d->nsrc; ++s) cin += d->Cin[s];
  const int taps = d->R * d->S;
  const long long total = static_cast<long long>(taps) * d->Cout * cin;
  pack_w_direct_kernel<<<grid_for(total, 256), 256, 0, stream>>>(w, d->Cout, cin, taps, d->transposed, w_packed);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("pack_w_direct_kernel launch");
  return 0;
}

// Backward of "normalise (+affine) -> activation" (BatchNorm2d batch statistics / InstanceNorm2d + ReLU / LeakyReLU),
// the training-time counterpart of norm.cu (reference: autograd through networks.py:25-44, 391-404, 490-513, 640-656).
//
// With n = y*scale + shift (scale = gamma*rstd), yhat = (y - mean)*rstd, a = act(n) and the incoming gradient
// dOut (= dL/da, optionally the sum of two tensors for the ResNet skip):
//     dn    = dOut * act'(n)
//     dbeta = sum dn,  dgamma = sum dn*yhat                    (per channel, over the statistics group(s))
//     dy    = scale * (dn - mean_g(dn) - yhat * mean_g(dn*yhat))   (g = the (n,c) plane, or the whole batch if pooled)
// Three passes, all HBM streams over NHWC fp32: reduce (per-128-pixel slice partials) -> finalize (fixed-order, fp64
// combine => deterministic) -> apply (writes fp32 dy and/or the hi/lo operand planes of the dgrad / wgrad GEMMs).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "internal.h"
#include "rng.cuh"
#include "stats_ws.h"

namespace dlb {
namespace {

constexpr int kSlicePx = 128;

__device__ __forceinline__ float dact(float n, int act) {
  if (act == DLB_ACT_RELU) return n > 0.f ? 1.f : 0.f;
  if (act == DLB_ACT_LRELU02) return n > 0.f ? 1.f : 0.2f;
  return 1.f;
}

struct BwdParams {
  const float* dout; const float* dout2; const float* y;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  int act, N, HW, C;
  int act2;                 // activation of the branch dout2 flows through (UNet: LeakyReLU down path + ReLU skip path)
  float drop_p; unsigned long long drop_seed;   // dropout that followed the activation on the dout branch (0 = off)
  const unsigned long long* drop_epoch;
};

// grid (slices, N), block 256: thread = (pixel lane, channel quad); fixed-order merge over the pixel lanes.
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const BwdParams p, int slices, StatsPtrs ws) {
  __shared__ float s1s[256 * 4];
  __shared__ float s2s[256 * 4];
  const int n = blockIdx.y, sl = blockIdx.x, tid = threadIdx.x;
  const int c4n = p.C / 4;
  const int p0 = sl * kSlicePx, p1 = min(p0 + kSlicePx, p.HW);
  if (tid == 0 && n == 0 && sl == 0) *ws.S = slices;
  for (int cq0 = 0; cq0 < c4n; cq0 += 256) {
    const int cols = min(c4n - cq0, 256);
    const int lanes = 256 / cols;
    const int cq = cq0 + tid % cols, pl = tid / cols;
    float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
    if (pl < lanes) {
      const float4 sc = *reinterpret_cast<const float4*>(p.scale + n * p.C + cq * 4);
      const float4 sh = *reinterpret_cast<const float4*>(p.shift + n * p.C + cq * 4);
      const float4 mu = *reinterpret_cast<const float4*>(p.mean + n * p.C + cq * 4);
      const float4 rs = *reinterpret_cast<const float4*>(p.rstd + n * p.C + cq * 4);
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
      const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
      for (int px = p0 + pl; px < p1; px += lanes) {
        const long long off = (static_cast<long long>(n) * p.HW + px) * p.C + cq * 4;
        const float4 yv = *reinterpret_cast<const float4*>(p.y + off);
        const float4 dv = *reinterpret_cast<const float4*>(p.dout + off);
        float4 d2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dout2 != nullptr) d2 = *reinterpret_cast<const float4*>(p.dout2 + off);
        const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, d2v[4] = {d2.x, d2.y, d2.z, d2.w};
        float dd[4] = {dv.x, dv.y, dv.z, dv.w};
        if (p.drop_p > 0.f) {
#pragma unroll
          for (int k = 0; k < 4; ++k) dd[k] *= dropout_scale(effective_seed(p.drop_seed, p.drop_epoch), static_cast<unsigned long long>(off) + k, p.drop_p);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float nn = fmaf(yy[k], scv[k], shv[k]);
          const float dn = dd[k] * dact(nn, p.act) + d2v[k] * dact(nn, p.act2);
          a1[k] += dn;
          a2[k] = fmaf(dn, (yy[k] - muv[k]) * rsv[k], a2[k]);
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) { s1s[tid * 4 + k] = a1[k]; s2s[tid * 4 + k] = a2[k]; }
    __syncthreads();
    if (pl == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t1 = 0.f, t2 = 0.f;
        for (int l = 0; l < lanes; ++l) { const int o = l * cols + (tid % cols); t1 += s1s[o * 4 + k]; t2 += s2s[o * 4 + k]; }
        ws.partial[(static_cast<long long>(n) * ws.S_cap + sl) * p.C + cq * 4 + k] = make_float2(t1, t2);
      }
    }
  }
}

// grid (C/32, groups), block 1024 (32 warps x 32 channels): c1 = mean(dn), c2 = mean(dn*yhat) per group; the
// parameter gradients dgamma/dbeta (sums over ALL samples) are produced by the blocks of group 0.
__global__ void __launch_bounds__(1024) norm_bwd_finalize_kernel(StatsPtrs ws, int N, int HW, int C, int pooled,
                                                                 float* __restrict__ c1, float* __restrict__ c2,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 int accumulate) {
  __shared__ double a[32][33];
  __shared__ double b[32][33];
  const int S = *ws.S;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool cok = c < C;
  const int n_lo = pooled ? 0 : blockIdx.y, n_hi = pooled ? N : blockIdx.y + 1;
  float t1 = 0.f, t2 = 0.f;
  if (cok)
    for (int n = n_lo; n < n_hi; ++n)
      for (int s = w; s < S; s += 32) {
        const float2 pr = ws.partial[(static_cast<long long>(n) * ws.S_cap + s) * C + c];
        t1 += pr.x; t2 += pr.y;
      }
  a[w][lane] = t1; b[w][lane] = t2;
  __syncthreads();
  if (w == 0 && cok) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) { s1 += a[k][lane]; s2 += b[k][lane]; }
    const double cnt = static_cast<double>(HW) * (n_hi - n_lo);
    for (int n = n_lo; n < n_hi; ++n) { c1[n * C + c] = static_cast<float>(s1 / cnt); c2[n * C + c] = static_cast<float>(s2 / cnt); }
  }
  // parameter gradients: total over every sample (per-sample statistics still share gamma/beta)
  if (dgamma == nullptr || blockIdx.y != 0) return;
  __syncthreads();
  float g1 = 0.f, g2 = 0.f;
  if (cok)
    for (int n = 0; n < N; ++n)
      for (int s = w; s < S; s += 32) {
        const float2 pr = ws.partial[(static_cast<long long>(n) * ws.S_cap + s) * C + c];
        g1 += pr.x; g2 += pr.y;
      }
  a[w][lane] = g1; b[w][lane] = g2;
  __syncthreads();
  if (w == 0 && cok) {
    double s1 = 0.0, s2 = 0.0;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) { s1 += a[k][lane]; s2 += b[k][lane]; }
    if (accumulate) { dbeta[c] += static_cast<float>(s1); dgamma[c] += static_cast<float>(s2); }
    else { dbeta[c] = static_cast<float>(s1); dgamma[c] = static_cast<float>(s2); }
  }
}

template <typename T16> struct Cv;
template <> struct Cv<__nv_bfloat16> {
  static __device__ __forceinline__ __nv_bfloat16 to(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ float from(__nv_bfloat16 v) { return __bfloat162float(v); }
};
template <> struct Cv<__half> {
  static __device__ __forceinline__ __half to(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float from(__half v) { return __half2float(v); }
};

// dy = scale*(dn - c1 - yhat*c2)  (or dy = dn when the layer has no norm: scale == nullptr, then n = y)
template <typename T16>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const BwdParams p, const float* __restrict__ c1,
                                                             const float* __restrict__ c2, float* __restrict__ dy_f32,
                                                             T16* __restrict__ dy_hi, T16* __restrict__ dy_lo) {
  const int c4n = p.C / 4;
  const long long total = static_cast<long long>(p.N) * p.HW * c4n;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(q % c4n);
    const int n = static_cast<int>(q / (static_cast<long long>(p.HW) * c4n));
    const long long off = q * 4;
    const float4 yv = __ldcs(reinterpret_cast<const float4*>(p.y + off));
    const float4 dv = __ldcs(reinterpret_cast<const float4*>(p.dout + off));
    float4 d2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.dout2 != nullptr) d2 = __ldcs(reinterpret_cast<const float4*>(p.dout2 + off));
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, d2v[4] = {d2.x, d2.y, d2.z, d2.w};
    float dd[4] = {dv.x, dv.y, dv.z, dv.w};
    if (p.drop_p > 0.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) dd[k] *= dropout_scale(effective_seed(p.drop_seed, p.drop_epoch), static_cast<unsigned long long>(off) + k, p.drop_p);
    }
    float o[4];
    if (p.scale != nullptr) {
      const int b = n * p.C + cq * 4;
      const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + b));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + b));
      const float4 mu = __ldg(reinterpret_cast<const float4*>(p.mean + b));
      const float4 rs = __ldg(reinterpret_cast<const float4*>(p.rstd + b));
      const float4 k1 = __ldg(reinterpret_cast<const float4*>(c1 + b));
      const float4 k2 = __ldg(reinterpret_cast<const float4*>(c2 + b));
      const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
      const float muv[4] = {mu.x, mu.y, mu.z, mu.w}, rsv[4] = {rs.x, rs.y, rs.z, rs.w};
      const float k1v[4] = {k1.x, k1.y, k1.z, k1.w}, k2v[4] = {k2.x, k2.y, k2.z, k2.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float nn = fmaf(yy[k], scv[k], shv[k]);
        const float dn = dd[k] * dact(nn, p.act) + d2v[k] * dact(nn, p.act2);
        const float yh = (yy[k] - muv[k]) * rsv[k];
        o[k] = scv[k] * (dn - k1v[k] - yh * k2v[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = dd[k] * dact(yy[k], p.act) + d2v[k] * dact(yy[k], p.act2);
    }
    if (dy_f32 != nullptr) *reinterpret_cast<float4*>(dy_f32 + off) = make_float4(o[0], o[1], o[2], o[3]);
    if (dy_hi != nullptr) {
      T16 hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { hi[k] = Cv<T16>::to(o[k]); lo[k] = Cv<T16>::to(o[k] - Cv<T16>::from(hi[k])); }
      *reinterpret_cast<uint2*>(dy_hi + off) = *reinterpret_cast<uint2*>(hi);
      if (dy_lo != nullptr) *reinterpret_cast<uint2*>(dy_lo + off) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

// ---- per-channel sum over (n, pixel): bias gradient of a convolution ------------------------------------------------
// grid (slices), block 256: slice partials [slices][C]; then one block per 32 channels sums the slices in order.
__global__ void __launch_bounds__(256) chsum_partial_kernel(const float* __restrict__ x, long long rows, int C, int rows_per,
                                                            float* __restrict__ part) {
  const int sl = blockIdx.x, tid = threadIdx.x;
  const long long r0 = static_cast<long long>(sl) * rows_per, r1 = min(r0 + rows_per, rows);
  for (int c = tid; c < C; c += 256) {
    float a = 0.f;
    for (long long r = r0; r < r1; ++r) a += x[r * C + c];
    part[static_cast<long long>(sl) * C + c] = a;
  }
}
__global__ void chsum_final_kernel(const float* __restrict__ part, int slices, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0;
  for (int s = 0; s < slices; ++s) a += part[static_cast<long long>(s) * C + c];
  out[c] = accumulate ? out[c] + static_cast<float>(a) : static_cast<float>(a);
}


// ---- padding backward ------------------------------------------------------------------------------------------------
// Reflection padding copies interior pixels into the border, so its backward sums every padded position that reads a
// pixel:  dx[h] = dpad[h + p] + (1 <= h <= p ? dpad[p - h] : 0) + (H-1-p <= h <= H-2 ? dpad[2(H-1) - h + p] : 0), same in w.
__device__ __forceinline__ int fold_sources(int h, int H, int p, int out[3]) {
  int n = 0;
  out[n++] = h + p;
  if (h >= 1 && h <= p) out[n++] = p - h;
  if (h >= H - 1 - p && h <= H - 2) out[n++] = 2 * (H - 1) - h + p;
  return n;
}

__global__ void __launch_bounds__(256) reflect_fold_kernel(const float* __restrict__ dpad, const float* __restrict__ add,
                                                           float* __restrict__ dx, int N, int H, int W, int C, int p) {
  const int c4n = C / 4, HP = H + 2 * p, WP = W + 2 * p;
  const long long total = static_cast<long long>(N) * H * W * c4n;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cq = static_cast<int>(i % c4n);
    long long r = i / c4n;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    int hs[3], ws[3];
    const int nh = fold_sources(h, H, p, hs), nw = fold_sources(w, W, p, ws);
    float4 acc = add ? *reinterpret_cast<const float4*>(add + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        const float4 v = *reinterpret_cast<const float4*>(dpad + ((static_cast<long long>(n) * HP + hs[a]) * WP + ws[b]) * C + cq * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    *reinterpret_cast<float4*>(dx + i * 4) = acc;
  }
}

// Backward of dlb_stem_window_pack: dXw fp32 [N, H+2p, W, 64] (lane s*8 + c) -> dx fp32 NCHW [N, C, H, W].
//   Xw[n, hp, w, s*8+c] = xpad[n, c, hp, w + s]  =>  dxpad[n, c, hp, wp] = sum_s dXw[n, hp, wp - s, s*8 + c]
// then the padding backward (crop for zero padding, the reflection sum above otherwise).
__global__ void __launch_bounds__(256) stem_window_bwd_kernel(const float* __restrict__ dxw, int N, int C, int H, int W, int p,
                                                              int S, int pad_mode, float* __restrict__ dx) {
  const long long total = static_cast<long long>(N) * H * W;
  const int HP = H + 2 * p;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    const int h = static_cast<int>((i / W) % H);
    const int n = static_cast<int>(i / (static_cast<long long>(W) * H));
    int hs[3], ws[3], nh = 1, nw = 1;
    hs[0] = h + p; ws[0] = w + p;
    if (pad_mode == DLB_PAD_REFLECT) { nh = fold_sources(h, H, p, hs); nw = fold_sources(w, W, p, ws); }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b)
        for (int s = 0; s < S; ++s) {
          const int wq = ws[b] - s;
          if (wq < 0 || wq >= W) continue;
          const float* q = dxw + ((static_cast<long long>(n) * HP + hs[a]) * W + wq) * 64 + s * 8;
          for (int c = 0; c < C; ++c) acc[c] += q[c];
        }
    for (int c = 0; c < C; ++c) dx[((static_cast<long long>(n) * C + c) * H + h) * W + w] = acc[c];
  }
}

}  // namespace
}  // namespace dlb

using namespace dlb;

// out[c] (+)= sum over rows of x[rows][C]  (bias gradient: rows = N*OH*OW of dy).  workspace: >= 1024*C floats.
extern "C" int dlb_channel_sum(const float* x, long long rows, int C, float* out, int accumulate, void* workspace,
                               size_t workspace_bytes, dlb_stream_t stream) {
  int slices = static_cast<int>((rows + 255) / 256);
  if (slices > 1024) slices = 1024;
  if (slices < 1) slices = 1;
  const int rows_per = static_cast<int>((rows + slices - 1) / slices);
  if (workspace_bytes < static_cast<size_t>(slices) * C * sizeof(float)) return set_error("dlb_channel_sum: workspace too small");
  float* part = reinterpret_cast<float*>(workspace);
  chsum_partial_kernel<<<slices, 256, 0, stream>>>(x, rows, C, rows_per, part);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("chsum_partial_kernel launch");
  chsum_final_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, slices, C, out, accumulate);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("chsum_final_kernel launch");
  return 0;
}

// dout2 (nullable): second addend of the incoming gradient.  scale == NULL: the layer has no norm (then only `apply`
// is meaningful: dy = dOut * act'(y)).  c1/c2: fp32 [N,C] scratch produced by reduce+finalize, consumed by apply.
extern "C" int dlb_norm_bwd(const float* dout, const float* dout2, const float* y, const float* scale, const float* shift,
                            const float* mean, const float* rstd, int act, int act2, int N, int HW, int C, int pooled,
                            float* c1, float* c2, float* dgamma, float* dbeta, int accumulate_param_grads,
                            float* dy_f32, void* dy_hi, void* dy_lo, int fmt, float drop_p, unsigned long long drop_seed,
                            const unsigned long long* drop_epoch, void* workspace, size_t workspace_bytes, dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_bwd: C % 4 != 0");
  BwdParams p{dout, dout2, y, scale, shift, mean, rstd, act, N, HW, C, act2, drop_p, drop_seed, drop_epoch};
  if (scale != nullptr) {
    const int c4n = C / 4;
    if ((c4n < 256 && 256 % c4n != 0) || (c4n > 256 && c4n % 256 != 0)) return set_error("dlb_norm_bwd: C/4 must divide 256");
    const StatsLayout L = stats_layout(N, HW, C);
    if (workspace_bytes < L.total) return set_error("dlb_norm_bwd: workspace too small");
    const StatsPtrs ws = stats_ptrs(workspace, L);
    const int slices = (HW + kSlicePx - 1) / kSlicePx;
    norm_bwd_reduce_kernel<<<dim3(slices, N), 256, 0, stream>>>(p, slices, ws);
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("norm_bwd_reduce_kernel launch");
    norm_bwd_finalize_kernel<<<dim3((C + 31) / 32, pooled ? 1 : N), 1024, 0, stream>>>(ws, N, HW, C, pooled, c1, c2, dgamma,
                                                                                     dbeta, accumulate_param_grads);
    if (cudaGetLastError() != cudaSuccess) return set_cuda_error("norm_bwd_finalize_kernel launch");
  }
  const long long total = static_cast<long long>(N) * HW * (C / 4);
  long long g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1;
  if (fmt == DLB_FMT_BF16)
    norm_bwd_apply_kernel<__nv_bfloat16><<<static_cast<int>(g), 256, 0, stream>>>(p, c1, c2, dy_f32,
        reinterpret_cast<__nv_bfloat16*>(dy_hi), reinterpret_cast<__nv_bfloat16*>(dy_lo));
  else if (fmt == DLB_FMT_FP16)
    norm_bwd_apply_kernel<__half><<<static_cast<int>(g), 256, 0, stream>>>(p, c1, c2, dy_f32,
        reinterpret_cast<__half*>(dy_hi), reinterpret_cast<__half*>(dy_lo));
  else return set_error("dlb_norm_bwd: bad fmt");
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("norm_bwd_apply_kernel launch");
  return 0;
}

extern "C" int dlb_reflect_fold(const float* dpad_nhwc, const float* add_nhwc, int N, int H, int W, int C, int pad,
                                float* dx_nhwc, dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_reflect_fold: C must be a multiple of 4");
  if (pad < 0 || pad >= H || pad >= W) return set_error("dlb_reflect_fold: need 0 <= pad < H, W");
  const long long total = static_cast<long long>(N) * H * W * (C / 4);
  long long g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1;
  reflect_fold_kernel<<<static_cast<int>(g), 256, 0, stream>>>(dpad_nhwc, add_nhwc, dx_nhwc, N, H, W, C, pad);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("reflect_fold_kernel launch");
  return 0;
}

extern "C" int dlb_stem_window_bwd(const float* dxw, int N, int C, int H, int W, int pad, int S, int pad_mode,
                                   float* dx_nchw, dlb_stream_t stream) {
  if (C < 1 || C > 8 || S < 1 || S > 8) return set_error("dlb_stem_window_bwd: needs C <= 8 and S <= 8");
  if (S != 2 * pad + 1) return set_error("dlb_stem_window_bwd: S must equal 2*pad + 1");
  if (pad_mode == DLB_PAD_REFLECT && (pad >= H || pad >= W)) return set_error("dlb_stem_window_bwd: bad pad");
  const long long total = static_cast<long long>(N) * H * W;
  long long g = (total + 255) / 256; if (g > 148 * 16) g = 148 * 16; if (g < 1) g = 1;
  stem_window_bwd_kernel<<<static_cast<int>(g), 256, 0, stream>>>(dxw, N, C, H, W, pad, S, pad_mode, dx_nchw);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stem_window_bwd_kernel launch");
  return 0;
}

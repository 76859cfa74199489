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
#include "rng.cuh"
#include "stats_ws.h"

namespace dlb {
namespace {

constexpr int kSlicePixels = 128;   // pixels per slice of the standalone statistics kernel

// ---- standalone pass 1: per-slice (sum, M2) partials; thread = 4 channels, coalesced float4 rows ----------
// grid (slices, N), block 256.  y: [N][HW][C].
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ y, int HW, int C, int slices,
                                                            StatsPtrs ws) {
  __shared__ float s_sum[256 * 4];
  __shared__ float s_m2[256 * 4];
  __shared__ int s_cnt[256];
  const int n = blockIdx.y, sl = blockIdx.x;
  const int c4n = C / 4;                           // float4 columns
  const int p0 = sl * kSlicePixels;
  const int p1 = min(p0 + kSlicePixels, HW);
  const int tid = threadIdx.x;
  if (tid == 0) {
    ws.cnt[n * ws.S_cap + sl] = static_cast<float>(p1 - p0);
    if (n == 0 && sl == 0) *ws.S = slices;
  }
  for (int cq0 = 0; cq0 < c4n; cq0 += 256) {
    const int cols = min(c4n - cq0, 256);
    const int lanes = 256 / cols;
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
      float cn = 0.f, cm[4] = {0, 0, 0, 0}, cM[4] = {0, 0, 0, 0};
      for (int l = 0; l < lanes; ++l) {            // Chan merge in lane order (deterministic)
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
        ws.partial[(static_cast<long long>(n) * ws.S_cap + sl) * C + cq * 4 + k] = make_float2(cm[k] * cn, cM[k]);
    }
    __syncthreads();
  }
}

// ---- finalize: merge slices -> scale/shift.  grid (C/32, groups), block 1024 (32 warps, lane = channel). -------
// Two independent-load passes over the slice partials instead of a serial Chan chain:
//   pass 1: n = sum n_i, mean = sum(sum_i) / n;   pass 2: M2 = sum(M2_i + n_i * (mean_i - mean)^2).
// Every thread accumulates its slices in index order (fp32), the 32 warps are combined in warp order (fp64):
// fixed orders -> deterministic.  groups = N (per-sample statistics) or 1 (pooled over the batch).
struct BnRunning {              // nn.BatchNorm2d(track_running_stats=True) buffers, updated in training (networks.py:34-37)
  float* mean; float* var; long long* num_batches_tracked; float momentum;
};

__global__ void __launch_bounds__(1024) stats_finalize_kernel(StatsPtrs ws, int N, int C, int pooled,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps,
                                                              float* __restrict__ scale, float* __restrict__ shift,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                              BnRunning run) {
  __shared__ double sm[32][33];
  __shared__ double sm2[32][33];
  const int S = *ws.S;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool cok = c < C;
  const int n_lo = pooled ? 0 : blockIdx.y, n_hi = pooled ? N : blockIdx.y + 1;
  // pass 1
  // The slice loops are latency-bound (one L2 round trip per slice): four slices are fetched before any is consumed, and
  // added in slice order afterwards, so the result does not depend on the unrolling.
  float cnt = 0.f, sum = 0.f;
  if (cok)
    for (int n = n_lo; n < n_hi; ++n) {
      const float* cntp = ws.cnt + static_cast<long long>(n) * ws.S_cap;
      const float2* prp = ws.partial + static_cast<long long>(n) * ws.S_cap * C + c;
      int s = w;
      for (; s + 96 < S; s += 128) {
        float nb[4]; float2 pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { nb[u] = cntp[s + 32 * u]; pr[u] = prp[static_cast<long long>(s + 32 * u) * C]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { cnt += nb[u]; sum += pr[u].x; }
      }
      for (; s < S; s += 32) { cnt += cntp[s]; sum += prp[static_cast<long long>(s) * C].x; }
    }
  sm[w][lane] = cnt; sm2[w][lane] = sum;
  __syncthreads();
  double tn = 0.0, ts = 0.0;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) { tn += sm[k][lane]; ts += sm2[k][lane]; }
  const float mean = static_cast<float>(ts / tn);
  __syncthreads();
  // pass 2
  float m2 = 0.f;
  if (cok)
    for (int n = n_lo; n < n_hi; ++n) {
      const float* cntp = ws.cnt + static_cast<long long>(n) * ws.S_cap;
      const float2* prp = ws.partial + static_cast<long long>(n) * ws.S_cap * C + c;
      int s = w;
      for (; s + 96 < S; s += 128) {
        float nb[4]; float2 pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { nb[u] = cntp[s + 32 * u]; pr[u] = prp[static_cast<long long>(s + 32 * u) * C]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (nb[u] > 0.f) { const float d = pr[u].x / nb[u] - mean; m2 += pr[u].y + nb[u] * d * d; }
      }
      for (; s < S; s += 32) {
        const float nb = cntp[s];
        const float2 pr = prp[static_cast<long long>(s) * C];
        if (nb > 0.f) { const float d = pr.x / nb - mean; m2 += pr.y + nb * d * d; }
      }
    }
  sm[w][lane] = m2;
  __syncthreads();
  if (w != 0 || !cok) return;
  double tm = 0.0;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) tm += sm[k][lane];
  const double var = tm / tn;                         // biased variance (PyTorch norm layers)
  const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float ga = gamma ? gamma[c] : 1.f;
  const float be = beta ? beta[c] : 0.f;
  const float sc = ga * rstd;
  const float sh = be - mean * sc;
  if (run.mean != nullptr && pooled) {
    // torch.nn.functional.batch_norm in training mode: running <- (1 - momentum) * running + momentum * batch statistic,
    // the variance one being the UNBIASED batch variance (n / (n - 1)); num_batches_tracked += 1 (once per call)
    const double unbias = tn > 1.0 ? tn / (tn - 1.0) : 1.0;
    run.mean[c] = (1.f - run.momentum) * run.mean[c] + run.momentum * mean;
    run.var[c] = (1.f - run.momentum) * run.var[c] + run.momentum * static_cast<float>(var * unbias);
    if (c == 0 && run.num_batches_tracked != nullptr) *run.num_batches_tracked += 1;
  }
  for (int n = n_lo; n < n_hi; ++n) {
    scale[n * C + c] = sc; shift[n * C + c] = sh;
    if (mean_out != nullptr) { mean_out[n * C + c] = mean; rstd_out[n * C + c] = rstd; }
  }
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
  float drop_p; unsigned long long drop_seed;     // training-time nn.Dropout after the activation (0 = off)
  const unsigned long long* drop_epoch;           // optional device counter mixed into the seed (CUDA-graph replays)
};

// grid (x: quads of one padded output row, y: rows n*HP + hp).  One thread = 4 channels of one output pixel;
// 32-bit index math only; every access is a 128-bit (fp32) or 64-bit (4 x 16-bit) vector.
template <typename T16>
__global__ void __launch_bounds__(256) norm_apply_kernel(const ApplyParams p) {
  const int c4n = p.C / 4;
  const int HP = p.H + 2 * p.pad, WP = p.W + 2 * p.pad;
  const int row_quads = WP * c4n;
  const int rows = p.N * HP;
  for (int row = blockIdx.y; row < rows; row += gridDim.y) {
    const int n = row / HP, hp = row - n * HP;
    int h = hp - p.pad;
    const bool hborder = (h < 0) || (h >= p.H);
    if (hborder && p.pad_mode == DLB_PAD_REFLECT) { if (h < 0) h = -h; if (h >= p.H) h = 2 * p.H - 2 - h; }
    const float* __restrict__ yrow = p.y + (static_cast<long long>(n) * p.H + h) * p.W * p.C;
    const float* __restrict__ rrow = p.residual ? p.residual + (static_cast<long long>(n) * p.H + h) * p.W * p.C : nullptr;
    float* __restrict__ frow = p.out_f32 ? p.out_f32 + (static_cast<long long>(n) * p.H + h) * p.W * p.C : nullptr;
    const long long drow = (static_cast<long long>(n) * HP + hp) * WP * p.C;
    constexpr int U = 4;                                 // quads per thread: all loads are issued before any use
    for (int q0 = blockIdx.x * blockDim.x * U + threadIdx.x; q0 < row_quads; q0 += gridDim.x * blockDim.x * U) {
      float4 v[U], rv[U];
      int srcs[U];
      bool zero[U], border[U], live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int q = q0 + u * blockDim.x;
        live[u] = q < row_quads;
        const int wp = q / c4n, cq = q - wp * c4n;
        int w = wp - p.pad;
        border[u] = hborder || (w < 0) || (w >= p.W);
        zero[u] = false;
        if (border[u]) {
          if (p.pad_mode == DLB_PAD_REFLECT) { if (w < 0) w = -w; if (w >= p.W) w = 2 * p.W - 2 - w; }
          else zero[u] = true;
        }
        srcs[u] = w * p.C + cq * 4;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f); rv[u] = v[u];
        if (live[u] && !zero[u]) {
          v[u] = __ldcs(reinterpret_cast<const float4*>(yrow + srcs[u]));
          if (rrow != nullptr) rv[u] = __ldcs(reinterpret_cast<const float4*>(rrow + srcs[u]));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
        const int q = q0 + u * blockDim.x;
        const int cq = q % c4n;
        float o[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        if (!zero[u]) {
          if (p.scale != nullptr) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n * p.C + cq * 4));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + n * p.C + cq * 4));
            o[0] = fmaf(o[0], sc.x, sh.x); o[1] = fmaf(o[1], sc.y, sh.y);
            o[2] = fmaf(o[2], sc.z, sh.z); o[3] = fmaf(o[3], sc.w, sh.w);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = act1(o[k], p.act);
          if (p.drop_p > 0.f) {
            const unsigned long long e = (static_cast<unsigned long long>(n) * p.H + h) * p.W * p.C + srcs[u];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] *= dropout_scale(effective_seed(p.drop_seed, p.drop_epoch), e + k, p.drop_p);
          }
          o[0] += rv[u].x; o[1] += rv[u].y; o[2] += rv[u].z; o[3] += rv[u].w;
          if (frow != nullptr && !border[u])
            *reinterpret_cast<float4*>(frow + srcs[u]) = make_float4(o[0], o[1], o[2], o[3]);
        }
        if (p.out_hi != nullptr) {
          const long long dst = drow + static_cast<long long>(q) * 4;
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
  }
}

// ---- stem operand: 7-wide horizontal windows of the padded input, packed into the channel dimension ------
// Xw[n, hp, w, s*8 + c] = xpad[n, hp, w + s, c]  (s < S taps, c < C <= 8; other lanes 0), xpad = pad-`pad` border of x.
// With this operand the Cin=3 7x7 stem (networks.py:386-397) is a 7-tap (vertical) conv with Cin = 64 on the
// tensor cores: K = (s, c) sits in one 128-byte swizzle row.  One thread = one (n, hp, w) pixel = 64 channels.
template <typename T16>
__global__ void __launch_bounds__(256) stem_window_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                                          int pad, int S, int pad_mode, T16* __restrict__ out_hi,
                                                          T16* __restrict__ out_lo) {
  const int HP = H + 2 * pad;
  const long long total = static_cast<long long>(N) * HP * W;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(idx % W);
    const int hp = static_cast<int>((idx / W) % HP);
    const int n = static_cast<int>(idx / (static_cast<long long>(W) * HP));
    int h = hp - pad;
    bool hout = (h < 0) || (h >= H);
    if (hout && pad_mode == DLB_PAD_REFLECT) { if (h < 0) h = -h; if (h >= H) h = 2 * H - 2 - h; hout = false; }
    __align__(16) T16 hi[64];
    __align__(16) T16 lo[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) { hi[k] = Cvt<T16>::to(0.f); lo[k] = Cvt<T16>::to(0.f); }
    if (!hout) {
      for (int s = 0; s < S; ++s) {
        int ww = w + s - pad;
        bool wout = (ww < 0) || (ww >= W);
        if (wout && pad_mode == DLB_PAD_REFLECT) { if (ww < 0) ww = -ww; if (ww >= W) ww = 2 * W - 2 - ww; wout = false; }
        if (wout) continue;
        for (int c = 0; c < C; ++c) {
          const float v = __ldg(x + ((static_cast<long long>(n) * C + c) * H + h) * W + ww);
          const T16 hh = Cvt<T16>::to(v);
          hi[s * 8 + c] = hh;
          lo[s * 8 + c] = Cvt<T16>::to(v - Cvt<T16>::from(hh));
        }
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

extern "C" size_t dlb_norm_stats_workspace(int N, int HW, int C) { return stats_layout(N, HW, C).total; }

static int launch_finalize(void* workspace, int N, int HW, int C, int pooled, const float* gamma, const float* beta,
                           float eps, float* scale, float* shift, float* mean, float* rstd, cudaStream_t stream,
                           BnRunning run = BnRunning{nullptr, nullptr, nullptr, 0.f}) {
  const StatsLayout L = stats_layout(N, HW, C);
  const StatsPtrs ws = stats_ptrs(workspace, L);
  dim3 grid(L.cchunks, pooled ? 1 : N);
  stats_finalize_kernel<<<grid, 1024, 0, stream>>>(ws, N, C, pooled, gamma, beta, eps, scale, shift, mean, rstd, run);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stats_finalize_kernel launch");
  return 0;
}

extern "C" int dlb_norm_finalize(void* workspace, size_t workspace_bytes, int N, int HW, int C, int pooled,
                                 const float* gamma, const float* beta, float eps, float* scale, float* shift,
                                 float* mean, float* rstd, dlb_stream_t stream) {
  if (workspace_bytes < stats_layout(N, HW, C).total) return set_error("dlb_norm_finalize: workspace too small");
  return launch_finalize(workspace, N, HW, C, pooled, gamma, beta, eps, scale, shift, mean, rstd,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int dlb_norm_stats(const float* y, int N, int HW, int C, int pooled, const float* gamma, const float* beta,
                              float eps, float* scale, float* shift, float* mean, float* rstd, void* workspace,
                              size_t workspace_bytes, dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_stats: C % 4 != 0");
  const int c4n = C / 4;
  if (c4n < 256 && 256 % c4n != 0) return set_error("dlb_norm_stats: C/4 must divide 256 (or be a multiple of 256)");
  if (c4n > 256 && c4n % 256 != 0) return set_error("dlb_norm_stats: C/4 must divide 256 (or be a multiple of 256)");
  const StatsLayout L = stats_layout(N, HW, C);
  if (workspace_bytes < L.total) return set_error("dlb_norm_stats: workspace too small");
  const int slices = (HW + kSlicePixels - 1) / kSlicePixels;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  stats_partial_kernel<<<dim3(slices, N), 256, 0, st>>>(y, HW, C, slices, stats_ptrs(workspace, L));
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stats_partial_kernel launch");
  return launch_finalize(workspace, N, HW, C, pooled, gamma, beta, eps, scale, shift, mean, rstd, st);
}

extern "C" int dlb_norm_finalize_bn(void* workspace, size_t workspace_bytes, int N, int HW, int C, const float* gamma,
                                    const float* beta, float eps, float* scale, float* shift, float* mean, float* rstd,
                                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                                    dlb_stream_t stream) {
  if (workspace_bytes < stats_layout(N, HW, C).total) return set_error("dlb_norm_finalize_bn: workspace too small");
  if (running_mean == nullptr || running_var == nullptr) return set_error("dlb_norm_finalize_bn: running buffers are null");
  return launch_finalize(workspace, N, HW, C, 1, gamma, beta, eps, scale, shift, mean, rstd,
                         reinterpret_cast<cudaStream_t>(stream), BnRunning{running_mean, running_var, num_batches_tracked, momentum});
}

extern "C" int dlb_norm_stats_bn(const float* y, int N, int HW, int C, const float* gamma, const float* beta, float eps,
                                 float* scale, float* shift, float* mean, float* rstd, float* running_mean, float* running_var,
                                 long long* num_batches_tracked, float momentum, void* workspace, size_t workspace_bytes,
                                 dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_stats_bn: C % 4 != 0");
  const int c4n = C / 4;
  if ((c4n < 256 && 256 % c4n != 0) || (c4n > 256 && c4n % 256 != 0))
    return set_error("dlb_norm_stats_bn: C/4 must divide 256 (or be a multiple of 256)");
  if (running_mean == nullptr || running_var == nullptr) return set_error("dlb_norm_stats_bn: running buffers are null");
  const StatsLayout L = stats_layout(N, HW, C);
  if (workspace_bytes < L.total) return set_error("dlb_norm_stats_bn: workspace too small");
  const int slices = (HW + kSlicePixels - 1) / kSlicePixels;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  stats_partial_kernel<<<dim3(slices, N), 256, 0, st>>>(y, HW, C, slices, stats_ptrs(workspace, L));
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stats_partial_kernel launch");
  return launch_finalize(workspace, N, HW, C, 1, gamma, beta, eps, scale, shift, mean, rstd, st,
                         BnRunning{running_mean, running_var, num_batches_tracked, momentum});
}

extern "C" int dlb_norm_apply(const float* y, const float* scale, const float* shift, int act, const float* residual,
                              float* out_f32, void* out_hi, void* out_lo, int fmt, int N, int H, int W, int C, int pad,
                              int pad_mode, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_epoch,
                              dlb_stream_t stream) {
  if (C % 4 != 0) return set_error("dlb_norm_apply: C % 4 != 0");
  if (pad < 0 || (pad_mode == DLB_PAD_REFLECT && (pad >= H || pad >= W))) return set_error("dlb_norm_apply: bad pad");
  if (out_hi == nullptr && out_f32 == nullptr) return set_error("dlb_norm_apply: no output");
  if (drop_p < 0.f || drop_p >= 1.f) return set_error("dlb_norm_apply: dropout p must be in [0, 1)");
  ApplyParams p{y, scale, shift, act, residual, out_f32, out_hi, out_lo, N, H, W, C, pad, pad_mode, drop_p, drop_seed, drop_epoch};
  const int row_quads = (W + 2 * pad) * (C / 4);
  const int rows = N * (H + 2 * pad);
  int gx = (row_quads + 1023) / 1024; if (gx > 64) gx = 64;
  dim3 grid(gx, rows < 65535 ? rows : 65535);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (fmt == DLB_FMT_BF16) norm_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (fmt == DLB_FMT_FP16) norm_apply_kernel<__half><<<grid, 256, 0, st>>>(p);
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

extern "C" int dlb_stem_window_pack(const float* x_nchw, int N, int C, int H, int W, int pad, int S, int pad_mode,
                                    int fmt, void* out_hi, void* out_lo, dlb_stream_t stream) {
  if (C < 1 || C > 8 || S < 1 || S > 8) return set_error("dlb_stem_window_pack: needs C <= 8 and S <= 8");
  if (pad_mode == DLB_PAD_REFLECT && (pad >= H || pad >= W)) return set_error("dlb_stem_window_pack: bad pad");
  const long long total = static_cast<long long>(N) * (H + 2 * pad) * W;
  const int grid = grid_for(total, 256);
  if (fmt == DLB_FMT_BF16)
    stem_window_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x_nchw, N, C, H, W, pad, S, pad_mode,
                                                              reinterpret_cast<__nv_bfloat16*>(out_hi),
                                                              reinterpret_cast<__nv_bfloat16*>(out_lo));
  else if (fmt == DLB_FMT_FP16)
    stem_window_kernel<__half><<<grid, 256, 0, stream>>>(x_nchw, N, C, H, W, pad, S, pad_mode,
                                                       reinterpret_cast<__half*>(out_hi), reinterpret_cast<__half*>(out_lo));
  else return set_error("dlb_stem_window_pack: bad fmt");
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("stem_window_kernel launch");
  return 0;
}

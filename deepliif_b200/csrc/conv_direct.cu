// fp32 CUDA-core convolution in tap-list form, for the layers that tensor cores cannot tile:
// Cin = 3 stem (networks.py:386-397), Cout = 3 head + Tanh (:438-444), UNet outermost convs (:576, :584),
// PatchGAN first (6->64) and last (512->1) convs (:638, :659).  These layers are HBM-/latency-bound
// (SURVEY.md §8a), so the kernel is organised around coalesced NHWC traffic and shared-memory reuse:
// a CTA stages an input halo patch (with the producer's norm+activation fused into the load) and the
// weight slab of one channel chunk in shared memory, and each thread keeps a PX-pixel x 4-channel
// register micro-tile.
#include "internal.h"

namespace dlb {
namespace {

struct DirectParams {
  int N, H, W, OH, OW, stride, ntaps;
  int tap_dh[64], tap_dw[64], tap_widx[64];
  int dh_min, dw_min, PH, PW;         // halo patch extents
  int cin, cin_pad, cout, CC;         // CC = channels per chunk (multiple of 4)
  const float* x; int in_nchw;
  const float* in_scale; const float* in_shift; int in_act; int pad_mode;
  const float* w; const float* bias; float* y; int out_act;
  long long ys_n, ys_h, ys_w, ys_c, y_base;
  int tiles_w, tiles_h;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == DLB_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DLB_ACT_LRELU02) return v > 0.f ? v : 0.2f * v;
  if (act == DLB_ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// TH x TW output pixels per CTA, PX consecutive pixels (along W) x 4 output channels per thread,
// NCG channel groups per CTA (CTA covers NCG*4 output channels).  256 threads.
template <int TH, int TW, int PX, int NCG>
__global__ void __launch_bounds__(256) conv_direct_kernel(const DirectParams p) {
  extern __shared__ float smem_f[];
  constexpr int COT = NCG * 4;
  constexpr int GPR = TW / PX;               // pixel groups per tile row
  static_assert(TH * GPR * NCG == 256, "256 threads");
  float* patch = smem_f;                                   // [CC/4][PH][PW] float4 (channel quads)
  float* wsm = smem_f + p.PH * p.PW * p.CC;                // [ntaps][CC][COT]

  const int tid = threadIdx.x;
  const int cg = tid % NCG;
  const int pg = tid / NCG;
  const int row = pg / GPR;
  const int wbase = pg % GPR;                // thread's pixels: wbase + j*GPR (lane-adjacent pixels => conflict-free LDS.128)

  int bt = blockIdx.x;
  const int tw = bt % p.tiles_w; bt /= p.tiles_w;
  const int th = bt % p.tiles_h; bt /= p.tiles_h;
  const int n = bt;
  const int oh0 = th * TH, ow0 = tw * TW;
  const int co0 = blockIdx.y * COT;

  float acc[PX][4];
#pragma unroll
  for (int j = 0; j < PX; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }

  const int ih0 = oh0 * p.stride + p.dh_min;
  const int iw0 = ow0 * p.stride + p.dw_min;

  for (int c0 = 0; c0 < p.cin_pad; c0 += p.CC) {
    __syncthreads();
    // ---- stage the input halo patch (fused producer transform) -------------------------------------
    const int patch_elems = p.PH * p.PW * p.CC;
    for (int idx = tid; idx < patch_elems; idx += 256) {
      const int c = idx % p.CC;
      const int pw = (idx / p.CC) % p.PW;
      const int ph = idx / (p.CC * p.PW);
      int ih = ih0 + ph, iw = iw0 + pw;
      const int ci = c0 + c;
      float v = 0.f;
      bool inside = (ih >= 0) && (ih < p.H) && (iw >= 0) && (iw < p.W);
      if (!inside && p.pad_mode == DLB_PAD_REFLECT) {
        ih = reflect_idx(ih, p.H); iw = reflect_idx(iw, p.W);
        inside = (ih >= 0) && (ih < p.H) && (iw >= 0) && (iw < p.W);
      }
      if (inside && ci < p.cin) {
        const long long off = p.in_nchw ? ((static_cast<long long>(n) * p.cin + ci) * p.H + ih) * p.W + iw
                                        : ((static_cast<long long>(n) * p.H + ih) * p.W + iw) * p.cin + ci;
        v = __ldg(p.x + off);
        if (p.in_scale != nullptr) v = fmaf(v, __ldg(p.in_scale + n * p.cin + ci), __ldg(p.in_shift + n * p.cin + ci));
        v = act_apply(v, p.in_act);
      }
      patch[(((c >> 2) * p.PH + ph) * p.PW + pw) * 4 + (c & 3)] = v;
    }
    // ---- stage the weight slab [tap][CC][COT] ------------------------------------------------------
    const int w_elems = p.ntaps * p.CC * COT;
    for (int idx = tid; idx < w_elems; idx += 256) {
      const int co = idx % COT;
      const int c = (idx / COT) % p.CC;
      const int t = idx / (COT * p.CC);
      const int ci = c0 + c;
      float v = 0.f;
      if (ci < p.cin && (co0 + co) < p.cout)
        v = __ldg(p.w + (static_cast<long long>(p.tap_widx[t]) * p.cin + ci) * p.cout + co0 + co);
      wsm[idx] = v;
    }
    __syncthreads();
    // ---- micro-tile FMAs ---------------------------------------------------------------------------
    for (int t = 0; t < p.ntaps; ++t) {
      const int pr = row * p.stride + p.tap_dh[t] - p.dh_min;
      const int pc = wbase * p.stride + p.tap_dw[t] - p.dw_min;
      const float* prow = patch + (pr * p.PW + pc) * 4;
      const int quad_stride = p.PH * p.PW * 4;
      const float* wt = wsm + t * p.CC * COT + cg * 4;
      for (int c = 0; c < p.CC; c += 4) {
        float4 wv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wv[k] = *reinterpret_cast<const float4*>(wt + (c + k) * COT);
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const float4 xv = *reinterpret_cast<const float4*>(prow + (c >> 2) * quad_stride + j * GPR * p.stride * 4);
          acc[j][0] = fmaf(xv.x, wv[0].x, acc[j][0]); acc[j][1] = fmaf(xv.x, wv[0].y, acc[j][1]);
          acc[j][2] = fmaf(xv.x, wv[0].z, acc[j][2]); acc[j][3] = fmaf(xv.x, wv[0].w, acc[j][3]);
          acc[j][0] = fmaf(xv.y, wv[1].x, acc[j][0]); acc[j][1] = fmaf(xv.y, wv[1].y, acc[j][1]);
          acc[j][2] = fmaf(xv.y, wv[1].z, acc[j][2]); acc[j][3] = fmaf(xv.y, wv[1].w, acc[j][3]);
          acc[j][0] = fmaf(xv.z, wv[2].x, acc[j][0]); acc[j][1] = fmaf(xv.z, wv[2].y, acc[j][1]);
          acc[j][2] = fmaf(xv.z, wv[2].z, acc[j][2]); acc[j][3] = fmaf(xv.z, wv[2].w, acc[j][3]);
          acc[j][0] = fmaf(xv.w, wv[3].x, acc[j][0]); acc[j][1] = fmaf(xv.w, wv[3].y, acc[j][1]);
          acc[j][2] = fmaf(xv.w, wv[3].z, acc[j][2]); acc[j][3] = fmaf(xv.w, wv[3].w, acc[j][3]);
        }
      }
    }
  }

  // ---- epilogue: bias, activation, store -----------------------------------------------------------
  const int oh = oh0 + row;
  if (oh >= p.OH) return;
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    const int ow = ow0 + wbase + j * GPR;
    if (ow >= p.OW) continue;
    float* yp = p.y + p.y_base + n * p.ys_n + oh * p.ys_h + ow * p.ys_w;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int co = co0 + cg * 4 + k;
      if (co < p.cout) {
        float v = acc[j][k];
        if (p.bias != nullptr) v += __ldg(p.bias + co);
        yp[co * p.ys_c] = act_apply(v, p.out_act);
      }
    }
  }
}

template <int TH, int TW, int PX, int NCG>
int launch_variant(DirectParams& p, cudaStream_t stream) {
  constexpr int COT = NCG * 4;
  p.tiles_w = (p.OW + TW - 1) / TW;
  p.tiles_h = (p.OH + TH - 1) / TH;
  p.PH = (TH - 1) * p.stride + 1;
  p.PW = (TW - 1) * p.stride + 1;
  int dh_min = p.tap_dh[0], dh_max = p.tap_dh[0], dw_min = p.tap_dw[0], dw_max = p.tap_dw[0];
  for (int t = 1; t < p.ntaps; ++t) {
    dh_min = min(dh_min, p.tap_dh[t]); dh_max = max(dh_max, p.tap_dh[t]);
    dw_min = min(dw_min, p.tap_dw[t]); dw_max = max(dw_max, p.tap_dw[t]);
  }
  p.dh_min = dh_min; p.dw_min = dw_min;
  p.PH += dh_max - dh_min; p.PW += dw_max - dw_min;
  // channel chunk: largest multiple of 4 (<= 16) whose patch + weights fit ~150 KB
  p.cin_pad = (p.cin + 3) & ~3;
  int cc = p.cin_pad < 16 ? p.cin_pad : 16;
  while (cc > 4 && (size_t)(p.PH * p.PW * cc + p.ntaps * cc * COT) * 4 > 150 * 1024) cc -= 4;
  p.CC = cc;
  p.cin_pad = ((p.cin + cc - 1) / cc) * cc;
  const size_t smem = (size_t)(p.PH * p.PW * cc + p.ntaps * cc * COT) * 4;
  if (smem > 200 * 1024) return set_error("conv_direct: tile does not fit shared memory");
  auto kern = conv_direct_kernel<TH, TW, PX, NCG>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
    return set_cuda_error("cudaFuncSetAttribute(conv_direct_kernel)");
  dim3 grid(p.tiles_w * p.tiles_h * p.N, (p.cout + COT - 1) / COT);
  kern<<<grid, 256, smem, stream>>>(p);
  if (cudaGetLastError() != cudaSuccess) return set_cuda_error("conv_direct_kernel launch");
  return 0;
}

}  // namespace

int launch_conv_direct_phase(const DirectPhase& ph, cudaStream_t stream) {
  DirectParams p;
  memset(&p, 0, sizeof(p));
  if (ph.ntaps < 1 || ph.ntaps > 64) return set_error("conv_direct: bad tap count");
  p.N = ph.N; p.H = ph.H; p.W = ph.W; p.OH = ph.OH; p.OW = ph.OW; p.stride = ph.stride; p.ntaps = ph.ntaps;
  for (int t = 0; t < ph.ntaps; ++t) { p.tap_dh[t] = ph.tap_dh[t]; p.tap_dw[t] = ph.tap_dw[t]; p.tap_widx[t] = ph.tap_widx[t]; }
  p.cin = ph.cin; p.cout = ph.cout;
  p.x = ph.x; p.in_nchw = ph.in_nchw; p.in_scale = ph.in_scale; p.in_shift = ph.in_shift; p.in_act = ph.in_act;
  p.pad_mode = ph.pad_mode; p.w = ph.w; p.bias = ph.bias; p.y = ph.y; p.out_act = ph.out_act;
  p.ys_n = ph.ys_n; p.ys_h = ph.ys_h; p.ys_w = ph.ys_w; p.ys_c = ph.ys_c; p.y_base = ph.y_base;
  if (ph.cout <= 4) return launch_variant<16, 64, 4, 1>(p, stream);
  if (ph.cout <= 32) return launch_variant<8, 32, 4, 4>(p, stream);
  return launch_variant<8, 16, 8, 16>(p, stream);
}

}  // namespace dlb

/* deepliif_b200 — C ABI of the B200-native (sm_100a) kernels behind the DeepLIIF tile-parallel cGAN path.
 *
 * The reference (nadeemlab/DeepLIIF) has no FFI / plugin registry: its hot path is
 * nn.Sequential.forward over nn.Conv2d / nn.ConvTranspose2d / BatchNorm2d|InstanceNorm2d / ReLU /
 * LeakyReLU / Tanh modules (deepliif/models/networks.py:386-446, 479-513, 576-615, 638-660).  The seam
 * a maintainer would bind is therefore "one library call per nn.Module group"; every entry point below
 * names the reference module(s) whose forward it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch caching allocator in practice);
 *     no entry point allocates device memory, frees or synchronises; every call only enqueues work on `stream`;
 *   - library state: (i) a per-device, mutex-guarded table of launch facts (SM count, "shared-memory limit raised" flags);
 *     (ii) per (host thread, device) three helper streams + events, created on first use by dlb_conv_tc_fwd[_fused] to
 *     run the independent output-parity phases of a small ConvTranspose2d side by side (forked from / joined back into
 *     `stream` with events, valid under stream capture) and destroyed by dlb_release_thread_resources(); nothing else;
 *   - activations are NHWC.  fp32 tensors are plain float; "split" tensors are two 16-bit planes
 *     (hi, lo) with hi = round16(x), lo = round16(x - hi), format DLB_FMT_BF16 or DLB_FMT_FP16;
 *   - return value 0 = ok, negative = error (message via dlb_last_error(), thread-local);
 *   - re-entrant across streams; one process per GPU is the expected deployment.
 */
#ifndef DEEPLIIF_B200_H_
#define DEEPLIIF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dlb_stream_t;

#define DLB_OK 0
#define DLB_ERR_INVALID (-1)
#define DLB_ERR_CUDA (-2)

enum { DLB_FMT_BF16 = 0, DLB_FMT_FP16 = 1 };
enum { DLB_ACT_NONE = 0, DLB_ACT_RELU = 1, DLB_ACT_LRELU02 = 2, DLB_ACT_TANH = 3 };
enum { DLB_PAD_ZERO = 0, DLB_PAD_REFLECT = 1 };

/* Geometry of one nn.Conv2d (transposed = 0) or nn.ConvTranspose2d (transposed = 1).
 * `nsrc`/`Cin[]`: the input may be given as up to two tensors concatenated along channels (the UNet
 * skip connection torch.cat([x, model(x)], 1), networks.py:615, is consumed without materialising it). */
typedef struct dlb_conv_desc {
  int N, H, W;
  int nsrc;
  int Cin[2];
  int Cout;
  int R, S;
  int stride;          /* 1 or 2 */
  int pad;             /* Conv2d/ConvTranspose2d `padding` */
  int transposed;
  int output_padding;  /* ConvTranspose2d only */
  int pad_mode;        /* DLB_PAD_ZERO, or DLB_PAD_REFLECT (= nn.ReflectionPad2d(pad) + conv padding 0) */
} dlb_conv_desc;

const char* dlb_last_error(void);
int dlb_version(void);
/* Destroys the helper streams / events the calling host thread created (see "library state" above). */
int dlb_release_thread_resources(void);

/* Output extent of the layer (PyTorch formulas). */
int dlb_conv_out_shape(const dlb_conv_desc* d, int* OH, int* OW);

/* ---- weight repacking (once per checkpoint load) ---------------------------------------------------
 * w: fp32, PyTorch layout: Conv2d (Cout, Cin_total, R, S); ConvTranspose2d (Cin_total, Cout, R, S).
 * tc:     two 16-bit planes [R*S][Cout][Cin_total]  (K-major B operand of tcgen05.mma)
 * direct: fp32 [R*S][Cin_total][Cout] */
int dlb_pack_weights_tc(const dlb_conv_desc* d, const float* w, int fmt, void* w_hi, void* w_lo, dlb_stream_t stream);
int dlb_pack_weights_direct(const dlb_conv_desc* d, const float* w, float* w_packed, dlb_stream_t stream);

/* ---- convolution forward ---------------------------------------------------------------------------
 * Replaces nn.Conv2d.forward / nn.ConvTranspose2d.forward (+bias) of networks.py:399-404, 425-430,
 * 490, 505, 576-600, 638-659 on the tcgen05 tensor cores.  Requires Cin[i] % 64 == 0, Cout % 32 == 0,
 * zero padding (reflect padding is produced by dlb_norm_apply into a padded operand buffer).
 * x_hi/x_lo: per source 16-bit NHWC planes; split != 0 selects the 3-MMA hi/lo scheme (fp32-grade
 * products), split == 0 a single pass on the hi planes.  y: fp32 NHWC [N, OH, OW, Cout].
 * n_tile: 0 = auto, or 64/128/256 (UMMA N).
 * stats_ws (nullable): when given, the epilogue also writes per-slice partial statistics of y into the
 * workspace (dlb_norm_stats_workspace(N, OH*OW, Cout) bytes, zero-initialised once at allocation), to be
 * reduced by dlb_norm_finalize — this replaces the separate statistics read of y.  Needs OH*OW >= 128. */
int dlb_conv_tc_fwd(const dlb_conv_desc* d, const void* const* x_hi, const void* const* x_lo, const void* w_hi,
                    const void* w_lo, const float* bias, float* y, int fmt, int split, int n_tile,
                    void* stats_ws, size_t stats_ws_bytes, dlb_stream_t stream);

/* Fused operand load: the same convolution, but the input is given as the PRODUCER's raw fp32 output plus the
 * normalisation / activation / skip-add that sits between the two layers in the reference
 * (ResnetBlock: Conv -> Norm -> ReLU -> Conv -> Norm -> +x, networks.py:490-513; ResnetGenerator down/up stages :399-436;
 * the Pad(3) in front of the head :438-444).  Converter warps inside the kernel evaluate
 *     a[n,h,w,c] = act(x[n,h,w,c] * scale[n,c] + shift[n,c]) + residual[n,h,w,c]
 * on the fly, split it into hi/lo 16-bit values and write them into the shared-memory operand tile the tensor core
 * reads — the operand planes never exist in HBM and dlb_norm_apply is not launched.
 *   x, residual (nullable), out (nullable): fp32 NHWC [N, H-2*border, W-2*border, Cin[s]];
 *   scale/shift (nullable together): fp32 [N, Cin[s]] from dlb_norm_finalize / dlb_norm_stats;
 *   border / border_mode: the conv input is the source behind `border` pixels of zeros or reflection
 *     (nn.ReflectionPad2d(border) + conv padding 0); d->H, d->W are the extents INCLUDING the border;
 *   out: when given, the evaluated operand `a` is also written there once (fp32) — the ResnetBlock residual stream
 *     for the next block; needs a stride-1 convolution with OH == H-2*border, OW == W-2*border.
 * One dlb_fused_src per K-source (d->nsrc).  Everything else as dlb_conv_tc_fwd. */
typedef struct dlb_fused_src {
  const float* x;
  const float* scale;
  const float* shift;
  int act;
  const float* residual;
  float* out;
  int border;
  int border_mode;
} dlb_fused_src;
/* Number of kernel launches dlb_conv_tc_fwd / dlb_conv_tc_fwd_fused make for this layer: 1 for a Conv2d and for a
 * ConvTranspose2d whose output-parity phases run merged (one input strip, one TMEM accumulator per phase), stride^2 otherwise. */
int dlb_conv_tc_launches(const dlb_conv_desc* d, int split, int n_tile, int fused);
/* How dlb_conv_tc_fwd_fused would run this layer (callers use it to choose between the fused call and
 * dlb_norm_apply + dlb_conv_tc_fwd; every mode computes the same function):
 *   2  every phase in halo-strip mode AND enough tensor-core work per converted strip (taps x N) that the converter warps
 *      stay hidden behind the MMAs — the 256 -> 256 3x3 ResNet-block convolutions in split precision: use the fused call;
 *   4  halo-strip mode with little MMA work per strip (ConvTranspose phases, narrow layers, single-pass precision) whose
 *      fp32 source can be staged through shared memory by TMA: efficient for a PLAIN source (one K-source, no residual, no
 *      write-back, no or zero border) — converters then read shared memory, not HBM; otherwise as 3;
 *   3  the same without room for the staging ring: the converters set the pace, the unfused pair is faster;
 *   1  vertical-strip mode with resident weights (R x 1 filters: the head) — converter-bound as well;
 *   0  at least one phase would convert per tap (stride-2 layers, maps below 16 x 8): slowest. */
int dlb_conv_tc_fused_mode(const dlb_conv_desc* d, int split, int n_tile);
int dlb_conv_tc_fwd_fused(const dlb_conv_desc* d, const dlb_fused_src* src, const void* w_hi, const void* w_lo,
                          const float* bias, float* y, int fmt, int split, int n_tile, void* stats_ws,
                          size_t stats_ws_bytes, dlb_stream_t stream);

/* ---- generator head in one kernel --------------------------------------------------------------------------
 * Replaces nn.ReflectionPad2d(3) | nn.ZeroPad2d(3) + nn.Conv2d(64, CO, 7) + nn.Tanh (reference networks.py:438-444)
 * together with the producer's normalisation + ReLU (networks.py:434-436), on the tensor cores:
 *   y[n, co, h, w] = out_act(bias[co] + sum_{ci,kh,kw} a[n, pad(h+kh-3), pad(w+kw-3), ci] * w[co, ci, kh, kw]),
 *   a = act(x * scale[n, ci] + shift[n, ci])          (scale / shift NULL: a = act(x))
 * x: the producer's raw fp32 NHWC [N, H, W, 64] output; y: fp32 NCHW [N, CO, H, W]; border_mode DLB_PAD_ZERO / _REFLECT
 * pads `a` (not x).  Split precision (bf16 hi + lo, three MMA passes), fp32 accumulation.  Row-streaming kernel: the
 * horizontal taps are the GEMM's K (shifted operand windows), the vertical taps are summed in the epilogue's registers;
 * nothing but x is read and nothing but y is written.  Needs C == 64, CO <= 3, H, W >= 8.
 * dlb_head_conv_pack_weights: w fp32 [CO][64][7][7] (PyTorch layout) -> dlb_head_conv_weight_bytes() bytes, once per model. */
size_t dlb_head_conv_weight_bytes(void);
int dlb_head_conv_pack_weights(const float* w, int CO, int C, int R, int S, void* out, dlb_stream_t stream);
int dlb_head_conv_fwd(const float* x, const float* scale, const float* shift, int act, int N, int H, int W, int C,
                      const void* w_packed, const float* bias, int CO, int border_mode, int out_act, float* y_nchw,
                      dlb_stream_t stream);

/* ---- generator stem in one kernel --------------------------------------------------------------------------
 * Replaces nn.ReflectionPad2d(3) | nn.ZeroPad2d(3) + nn.Conv2d(C, 64, 7) (reference networks.py:386-397) on the tensor
 * cores, from the fp32 NCHW network input [N, C <= 4, H, W]:
 *   y[n, h, w, co] = bias[co] + sum_{c,kh,kw} pad(x)[n, c, h+kh-3, w+kw-3] * w[co, c, kh, kw]        fp32 NHWC [N, H, W, 64]
 * and, when stats_ws is given (dlb_norm_stats_workspace bytes for (N, H*W, 64)), the partial normalisation statistics of y
 * for dlb_norm_finalize (one slice per 128-pixel row tile).  Row-streaming kernel without an im2col operand: every input
 * pixel is one 16-byte slot [hi(c), lo(c)] in shared memory and the MMA reads the horizontal filter window of an output
 * pixel in place through a non-swizzled K-major descriptor with overlapping core matrices (LBO 16 B, SBO 128 B);
 * split precision in one pass (128 accumulator columns, halves added in the epilogue).  Needs H, W >= 8.
 * dlb_stem_conv_pack_weights: w fp32 [64][C][7][7] (PyTorch layout) -> dlb_stem_conv_weight_bytes() bytes, once per model. */
size_t dlb_stem_conv_weight_bytes(void);
int dlb_stem_conv_pack_weights(const float* w, int Cout, int C, int R, int S, void* out, dlb_stream_t stream);
int dlb_stem_conv_fwd(const float* x_nchw, int N, int C, int H, int W, const void* w_packed, const float* bias, int Cout,
                      int border_mode, float* y, void* stats_ws, size_t stats_ws_bytes, dlb_stream_t stream);

/* fp32 CUDA-core convolution for the layers tensor cores cannot tile (Cin = 3 stem, networks.py:386-397;
 * Cout = 3 head + Tanh, :438-444; PatchGAN first/last convs, :638, :659).  Fuses the producer's
 * normalisation + activation on the input side: x' = act_in(x * in_scale[n,c] + in_shift[n,c]) (zero /
 * reflect padding applies to x'), bias, optional output activation, optional NCHW input / output so the
 * network boundary needs no layout pass.  Single source (nsrc == 1). */
int dlb_conv_direct_fwd(const dlb_conv_desc* d, const float* x, int in_nchw, const float* in_scale,
                        const float* in_shift, int in_act, const float* w_packed, const float* bias, float* y,
                        int out_act, int out_nchw, dlb_stream_t stream);

/* ---- normalisation -----------------------------------------------------------------------------------
 * Replaces BatchNorm2d (batch statistics) / InstanceNorm2d of networks.py:25-44.
 * dlb_norm_stats: y fp32 NHWC [N, HW, C] -> per-(n,c) scale/shift such that norm(y) = y*scale + shift,
 *   scale = gamma * rstd, shift = beta - mean * scale (gamma/beta may be NULL = 1/0), biased variance.
 *   pooled != 0: statistics over N*HW (training-mode BatchNorm2d, N > 1), replicated for every n.
 *   Deterministic (fixed merge orders; the only atomic is a completion ticket).  workspace:
 *   dlb_norm_stats_workspace bytes, zero-initialised once at allocation (calls leave it clean).
 * dlb_norm_finalize: the reduction half of dlb_norm_stats, for partials written by dlb_conv_tc_fwd.
 *   mean/rstd (nullable, fp32 [N,C]): the statistics themselves, kept for the backward pass.
 * dlb_norm_apply: out = act(y*scale + shift) (+ residual), written as fp32 (out_f32) and/or as split
 *   16-bit planes (out_hi/out_lo) for the next tensor-core conv; `pad` > 0 writes the planes into a
 *   [N, H+2pad, W+2pad, C] buffer with a reflected (DLB_PAD_REFLECT) or zero border.
 *   drop_epoch (device uint64, may be NULL): added into the seed so a captured CUDA graph draws new masks per replay.
 *   drop_p > 0: training-time nn.Dropout(drop_p) right after the activation (networks.py:493-494, 604-605) with the
 *   counter-based mask keep(seed, element index) of csrc/rng.cuh; dlb_norm_bwd regenerates the same mask. */
size_t dlb_norm_stats_workspace(int N, int HW, int C);
int dlb_norm_finalize(void* workspace, size_t workspace_bytes, int N, int HW, int C, int pooled, const float* gamma,
                      const float* beta, float eps, float* scale, float* shift, float* mean, float* rstd,
                      dlb_stream_t stream);
int dlb_norm_stats(const float* y, int N, int HW, int C, int pooled, const float* gamma, const float* beta,
                   float eps, float* scale, float* shift, float* mean, float* rstd, void* workspace,
                   size_t workspace_bytes, dlb_stream_t stream);
/* Training-mode nn.BatchNorm2d (track_running_stats=True, networks.py:34-37): the pooled (batch) forms of the two calls
 * above that also update the module's buffers the way torch does — running_mean / running_var <- (1-momentum)*running +
 * momentum*(batch mean / UNBIASED batch variance), num_batches_tracked (int64 scalar, nullable) += 1 — inside the same
 * finalize kernel, so checkpoints written after training carry the statistics the reference would have written
 * (base_model.py:190-212). */
int dlb_norm_finalize_bn(void* workspace, size_t workspace_bytes, int N, int HW, int C, const float* gamma, const float* beta,
                         float eps, float* scale, float* shift, float* mean, float* rstd, float* running_mean,
                         float* running_var, long long* num_batches_tracked, float momentum, dlb_stream_t stream);
int dlb_norm_stats_bn(const float* y, int N, int HW, int C, const float* gamma, const float* beta, float eps, float* scale,
                      float* shift, float* mean, float* rstd, float* running_mean, float* running_var,
                      long long* num_batches_tracked, float momentum, void* workspace, size_t workspace_bytes,
                      dlb_stream_t stream);
int dlb_norm_apply(const float* y, const float* scale, const float* shift, int act, const float* residual,
                   float* out_f32, void* out_hi, void* out_lo, int fmt, int N, int H, int W, int C, int pad,
                   int pad_mode, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_epoch,
                   dlb_stream_t stream);

/* ---- training: backward of norm + activation -------------------------------------------------------------------
 * Autograd of BatchNorm2d (batch statistics) / InstanceNorm2d + ReLU / LeakyReLU(0.2) (networks.py:25-44, 391-404,
 * 490-513, 640-656).  With n = y*scale + shift, yhat = (y-mean)*rstd, dn = dout*act'(n) [+ dout2*act2'(n)] (the second
 * term: a tensor consumed through two activations, e.g. the UNet skip: LeakyReLU into the down conv, ReLU into the up conv):
 *   dbeta = sum dn, dgamma = sum dn*yhat (nullable; += when accumulate_param_grads),
 *   dy = scale * (dn - mean_g(dn) - yhat * mean_g(dn*yhat)),  g = (n,c) plane, or the batch when pooled.
 * scale == NULL: layer without norm, dy = dn with n = y.  c1/c2: fp32 [N,C] scratch.  dy is written as fp32 and/or
 * as hi/lo planes (operands of the dgrad / wgrad GEMMs).  workspace: dlb_norm_stats_workspace(N, HW, C) bytes. */
int dlb_norm_bwd(const float* dout, const float* dout2, const float* y, const float* scale, const float* shift,
                 const float* mean, const float* rstd, int act, int act2, int N, int HW, int C, int pooled, float* c1, float* c2,
                 float* dgamma, float* dbeta, int accumulate_param_grads, float* dy_f32, void* dy_hi, void* dy_lo,
                 int fmt, float drop_p, unsigned long long drop_seed, const unsigned long long* drop_epoch, void* workspace,
                 size_t workspace_bytes, dlb_stream_t stream);

/* Fused Adam on a flat fp32 bucket (torch.optim.Adam semantics, DeepLIIF_model.py:133-146); g is scaled by grad_scale
 * first (1/world_size after a sum all-reduce).  step counts from 1. */
int dlb_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                  int step, float grad_scale, dlb_stream_t stream);
/* CUDA-graph form: the step-dependent scalars {lr, 1-beta1^t, sqrt(1-beta2^t), grad_scale} are read from device memory
 * (hyper4_dev), so one captured launch serves every replay; dlb_adam_hyper fills the same four floats on the host with
 * the arithmetic dlb_adam_step uses (bit-identical updates). */
int dlb_adam_hyper(float lr, float beta1, float beta2, int step, float grad_scale, float* hyper4_host);
int dlb_adam_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper4_dev, float beta1,
                      float beta2, float eps, dlb_stream_t stream);

/* Bias gradient: out[c] (+)= sum_rows x[rows][C] (x = dy as fp32 NHWC, rows = N*OH*OW).  workspace >= 1024*C floats. */
int dlb_channel_sum(const float* x, long long rows, int C, float* out, int accumulate, void* workspace,
                    size_t workspace_bytes, dlb_stream_t stream);

/* Weight gradient of a convolution on the tensor cores (autograd of nn.Conv2d / nn.ConvTranspose2d wrt weight):
 *   Conv2d:          dW[co][ci][r][s] = sum_{n,oh,ow} dy[n,oh,ow,co] * x[n, oh*st + r - pad, ow*st + s - pad, ci]
 *   ConvTranspose2d: dW[ci][co][r][s] = sum_{n,ih,iw} x[n,ih,iw,ci] * dy[n, ih*st + r - pad, iw*st + s - pad, co]
 * i.e. per tap a [P-channels x Q-channels] GEMM whose K dimension runs over the pixels of the low-resolution
 * ("anchor") tensor P (dy for Conv2d, x for ConvTranspose2d); Q is read through the same shifted / stride-2 TMA views
 * as the forward operand.  d describes the FORWARD layer (nsrc == 1).  x/dy: hi/lo NHWC planes.  dw: fp32 in the
 * PyTorch weight layout; accumulate != 0 adds to it.  Split-K partials go to `workspace`
 * (dlb_conv_wgrad_workspace bytes) and are reduced in a fixed order (deterministic). */
size_t dlb_conv_wgrad_workspace(const dlb_conv_desc* d);
int dlb_conv_wgrad(const dlb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo,
                   float* dw, int accumulate, int fmt, int split, void* workspace, size_t workspace_bytes,
                   dlb_stream_t stream);

/* Stem operand for the tensor cores.  The 7x7 Cin=3 stem conv (ReflectionPad2d/ZeroPad2d(3) + Conv2d(3, ngf, 7),
 * networks.py:386-397) has K = 147, too ragged for 64-channel K chunks; this pass writes
 *   Xw[n, hp, w, s*8 + c] = pad(x)[n, c, hp, w + s]   (s < S, c < C <= 8, zero elsewhere; [N, H+2pad, W, 64] planes)
 * so that the stem becomes a vertical S=1, R=7, Cin=64 convolution for dlb_conv_tc_fwd with weights
 * wk[co][s*8 + c][r] = w[co][c][r][s].  x: fp32 NCHW. */
int dlb_stem_window_pack(const float* x_nchw, int N, int C, int H, int W, int pad, int S, int pad_mode, int fmt,
                         void* out_hi, void* out_lo, dlb_stream_t stream);

/* The stem in one kernel: ReflectionPad2d/ZeroPad2d(pad) + Conv2d(C <= 4, Cout, S x S) (networks.py:386-397) straight from the
 * fp32 NCHW network input.  Converter warps build the window operand of dlb_stem_window_pack inside shared memory (that
 * tensor, 21x the input, is never written to HBM) and the S x 1 vertical convolution over it runs on the tensor cores in
 * vertical-strip mode; w_hi / w_lo are the planes of wk[co][s*8 + c][r] packed by dlb_pack_weights_tc (R = S, S = 1,
 * Cin = 64).  y: fp32 NHWC [N, H, W, Cout]; stats_ws as in dlb_conv_tc_fwd.  Needs H >= 16, W >= 8, S = 2 * pad + 1 <= 8. */
int dlb_conv_tc_fwd_stem(const float* x_nchw, int N, int C, int H, int W, int pad, int S, int pad_mode, int Cout,
                         const void* w_hi, const void* w_lo, const float* bias, float* y, int fmt, int split, int n_tile,
                         void* stats_ws, size_t stats_ws_bytes, dlb_stream_t stream);

/* Head finish.  The 7x7 Cout=3 head (Pad(3) + Conv2d(ngf, 3, 7) + Tanh, networks.py:438-444) has N = 3, far too
 * narrow for a tensor-core tile; it is run as a vertical R=7, S=1 convolution with the S horizontal taps moved
 * into 32 virtual output channels j = s*4 + co (weights wv[j][c][r] = w[co][c][r][s]) over the operand padded
 * by 3 (dlb_norm_apply pad=3), z = [N, H, W+S-1, 32]; this pass then forms
 *   y[n, co, h, w] = act(bias[co] + sum_s z[n, h, w + s, s*4 + co])     (fp32 NCHW out). */
int dlb_head_finish(const float* z, const float* bias, int N, int H, int W, int S, int CO, int act, float* y_nchw,
                    dlb_stream_t stream);

/* Backward of dlb_head_finish: scatters dzz = dL/d(pre-activation) (fp32 NCHW [N,CO,H,W]) into the 64-lane virtual
 * channel planes dz[n, h, u, s*4+co] = dzz[n, co, h, u-s], [N, H, W+S-1, 64] hi/lo (operand of the head's dgrad/wgrad). */
int dlb_head_bwd_pack(const float* dzz_nchw, int N, int H, int W, int S, int CO, int fmt, void* out_hi, void* out_lo,
                      dlb_stream_t stream);

/* ---- pixel ends ------------------------------------------------------------------------------------
 * dlb_u8_to_f32: deepliif.data.transform (data/__init__.py:133-138): uint8 HWC -> fp32 NCHW in [-1,1].
 * dlb_seg_finish: run_dask seg aggregation (models/__init__.py:338) + tensor2im quantisation
 *   (util/util.py:130-135) + create_posneg_mask (postprocessing.py:163-190) in one pass:
 *   seg = sum_k w_k * segs[k] (fp32, list order); u8 = trunc((seg+1)/2*255); mask from u8.
 *   segs: nseg device pointers to fp32 NCHW [N,3,H,W]; seg_f32 / seg_u8 (NHWC) / mask may be NULL. */
/* is_empty() support (models/__init__.py:391-396, util/__init__.py:478-485): per tile, over the pixels whose PIL 'L' luma
 * is neither 0 nor 255 (the reference drops saturated pixels first): sums = uint64 [N][3] = count n, sum, sum of squares
 * (exact integers); variance = s2/n - (s1/n)^2 on the host, 0 when n = 0 (tiles below 9 skip the networks). */
int dlb_tile_luma_sums(const uint8_t* img_nhwc, int N, int H, int W, unsigned long long* sums, dlb_stream_t stream);
int dlb_u8_to_f32(const uint8_t* img_nhwc, float* out_nchw, int N, int H, int W, dlb_stream_t stream);
int dlb_f32_to_u8(const float* x_nchw, uint8_t* out_nhwc, int N, int H, int W, dlb_stream_t stream);
int dlb_seg_finish(const float* const* segs, const float* weights, int nseg, int N, int H, int W, int thresh,
                   float* seg_f32_nchw, uint8_t* seg_u8_nhwc, uint8_t* mask, dlb_stream_t stream);

/* ---- padding backward (training with padding_type='reflect', generator input gradients) ----------------------------
 * dlb_reflect_fold: backward of nn.ReflectionPad2d(pad) (networks.py:386, 438, 481-499) on fp32 NHWC:
 *   dx[n,h,w,c] = (add ? add[n,h,w,c] : 0) + sum of dpad over every padded position that mirrors (h,w);
 *   dpad [N,H+2pad,W+2pad,C], dx / add [N,H,W,C]; pad = 0 is a plain (add +) copy.
 * dlb_stem_window_bwd: backward of dlb_stem_window_pack incl. its zero / reflect padding:
 *   dxw fp32 [N,H+2pad,W,64] (lane s*8+c, S = 2*pad+1) -> dx fp32 NCHW [N,C,H,W] (the generator's input gradient:
 *   in the DeepLIIF cascade the seg generators sit behind the modality generators, DeepLIIF_model.py:175-203). */
int dlb_reflect_fold(const float* dpad_nhwc, const float* add_nhwc, int N, int H, int W, int C, int pad, float* dx_nhwc,
                     dlb_stream_t stream);
int dlb_stem_window_bwd(const float* dxw, int N, int C, int H, int W, int pad, int S, int pad_mode, float* dx_nchw,
                        dlb_stream_t stream);

/* ---- cell post-processing on the stitched uint8 images ----------------------------------------
 * The integer graph work of deepliif/postprocessing.py, each call = one reference function, results bit-exact:
 *   dlb_cells_posneg_mask      create_posneg_mask (:163-190): seg uint8 [H,W,3] -> mask uint8 [H,W] (50/150/200).
 *   dlb_cells_marker_plane     mode 0: to_array(marker, grayscale=True) (:98-120) = max over channels, plus the 256-bin
 *                              histogram of the non-zero values (calculate_stain_range :450-469; hist256 may be NULL);
 *                              mode 1: create_od_image (:123-138) with the caller's float64 LUT [256]
 *                              (lut[0] = lut[1] = log10(255), lut[i] = log10(255 / i)).  plane: uint16 [H,W].
 *   dlb_cells_mark_background  mark_background (:193-232), in place; labels_ws: int32 [H*W] scratch.
 *   dlb_cells_label            the cell search of compute_cell_mapping (:235-308): 8-connected components of the pixels
 *                              that are neither BACKGROUND (0) nor CELL (100), numbered in raster order of their first
 *                              pixel.  labels: int32 [H*W] (component index, -1 elsewhere); roots: int32 [cap] first
 *                              pixel (y*W+x) per component, cap >= ceil(H/2)*ceil(W/2); n_cells: device int32.
 *   dlb_cells_stats            per component int64 [n][8] = count, count_pos, count_neg, marker (max, or sum when
 *                              use_avg), x0, y0, sum_x, sum_y (the host applies the noise filter and the rounding).
 *   dlb_cells_classify         create_cell_classification (:923-1000): cls uint8 [n] = 0 (left as LABEL_CELL 100),
 *                              1 (negative), 2 (positive) -> mask_out uint8 [H,W] with 200/150 cells and 220/170 borders.
 *   dlb_cells_enlarge          one enlarge_cell_boundaries pass (:1003-1030), out of place.
 *   dlb_cells_final_images     create_final_images (:1033-1071): overlay (copy of orig with coloured borders), refined.
 * All pointers are device pointers; H*W < 2^31. */
int dlb_cells_posneg_mask(const uint8_t* seg_hwc, int H, int W, int thresh, uint8_t* mask, dlb_stream_t stream);
int dlb_cells_marker_plane(const uint8_t* img_hwc, int H, int W, int mode, const double* od_lut, uint16_t* plane,
                           unsigned int* hist256, dlb_stream_t stream);
int dlb_cells_mark_background(uint8_t* mask, int H, int W, int* labels_ws, dlb_stream_t stream);
size_t dlb_cells_label_workspace(int H, int W);
int dlb_cells_label(const uint8_t* mask, int H, int W, int* labels, int* roots, int* n_cells, void* ws, size_t ws_bytes,
                    dlb_stream_t stream);
int dlb_cells_stats(const uint8_t* mask, const uint16_t* marker, const int* labels, const int* roots, int n, int H, int W,
                    int use_avg, long long* table, dlb_stream_t stream);
int dlb_cells_classify(const int* labels, const int* roots, const uint8_t* cls, int H, int W, uint8_t* mask_out,
                       dlb_stream_t stream);
int dlb_cells_enlarge(const uint8_t* mask_in, uint8_t* mask_out, int H, int W, dlb_stream_t stream);
int dlb_cells_final_images(const uint8_t* orig_hwc, const uint8_t* mask, int H, int W, uint8_t* overlay_hwc,
                           uint8_t* refined_hwc, dlb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DEEPLIIF_B200_H_ */

// Internal (non-ABI) declarations shared by the translation units of libdeepliif_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/deepliif_b200.h"

namespace dlb {

int set_error(const char* msg);            // records msg for dlb_last_error(); returns DLB_ERR_INVALID
int set_cuda_error(const char* where);     // records where + cudaGetErrorString; returns DLB_ERR_CUDA

// Per-device, mutex-guarded launch facts (the only process-wide mutable state of the library).
//   ensure_dyn_smem: raise `func`'s dynamic shared-memory limit once per device (slot = small id of the kernel);
//   device_num_sms:  multiprocessor count of the current device.
enum { kSlotConvTc = 0, kSlotWgradMt = 1, kSlotWgrad = 2, kSlotConvTc2 = 3, kSlotHeadConv = 4, kSlotStemConv = 5, kNumSmemSlots = 8 };
int ensure_dyn_smem(const void* func, int bytes, int slot);
int device_num_sms(int* num_sms);

// One phase of a convolution in tap-list form:
//   y[n,i,j,co] = sum_t sum_ci x[n, i*stride + dh_t, j*stride + dw_t, ci] * w[widx_t][co][ci]
struct PhaseGeom {
  int N, H, W;                 // input extents (per source; sources share N,H,W)
  int OH, OW;                  // logical output extents of this phase
  int stride;                  // 1 or 2 (input step per output pixel)
  int ntaps;
  int tap_dh[64], tap_dw[64], tap_widx[64];   // tensor-core path uses <= 16
  int w_taps;                  // taps in the packed weight tensor (R*S)
  long long ys_n, ys_h, ys_w, y_base;   // output element strides / base (NHWC, possibly strided phase)
  int cout;
};

struct TcPhase : PhaseGeom {
  int nsrc;
  int cin[2];
  const void* x_hi[2];
  const void* x_lo[2];
  const void* w_hi;
  const void* w_lo;
  const float* bias;
  float* y;
  int fmt;                     // DLB_FMT_BF16 / DLB_FMT_FP16
  int split;                   // 1: hi+lo planes, 3 MMAs per K step; 0: single pass
  int n_tile;                  // 0 = auto
  int max_stages;              // 0 = as many as fit
  int max_ctas;                // 0 = #SMs
  int no_vs;                   // 1 = never use the vertical-strip mode (A/B testing)
  int no_hs;                   // 1 = never use the plane-fed halo-strip mode (A/B testing)
  int no_cta2;                 // 1 = never pair CTAs (cta_group::2)
  // merged output-parity phases of a ConvTranspose2d (nacc > 1): tap t accumulates into accumulator tap_acc[t]; accumulator a
  // is written at output base acc_ybase[a] with statistics slice base acc_slice[a].  nacc = 0 / 1: a single accumulator.
  int nacc;
  int tap_acc[64];
  long long acc_ybase[4];
  int acc_slice[4];
  // fused normalisation statistics (see stats_ws.h); st_partial == nullptr disables
  float2* st_partial;
  float* st_cnt;
  int* st_S;
  int st_S_cap, st_slice_base, st_S_total;
  // fused-operand mode (fa != 0): x_hi/x_lo are unused; the kernel builds the operand planes itself from the producer's
  // raw fp32 output: act(x*scale + shift) + residual, seen behind a zero / reflected border of fa_border pixels
  // (H, W of the phase are the extents INCLUDING the border; the source tensors are [N, H-2b, W-2b, cin]).
  int fa;
  const float* fa_x[2];
  const float* fa_scale[2];
  const float* fa_shift[2];
  const float* fa_res[2];
  float* fa_out[2];
  int fa_act[2];
  int fa_border, fa_border_mode;
  int stem_C, stem_S, stem_pad;   // fa == 2: fa_x[0] is the fp32 NCHW network input, see TcParams
};

// cuTensorMapEncodeTiled (16-bit elements, 128-byte swizzle, zero OOB fill) through the runtime's driver entry point.
bool encode_tiled_map(CUtensorMap* map, const void* base, int is_bf16, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box);

// fp32 tiled map (zero OOB fill / clipped stores); swizzle128: 128-byte swizzle (inner box extent 32 floats), else none.
bool encode_f32_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle128 = 0);

// Returns the phase's execution mode: 0 tap mode, 1 vertical strip (resident weights), 2 halo strip (fa only).
int tc_plan_tiles(const PhaseGeom& g, int nsrc, const int* cin, int cout, int split, int n_tile_req, int* tile_w,
                  int* tile_h, int* tile_n, int* n_tile_out, int fa = 0, int allow_hp = 1);

bool hs_staging_fits(const PhaseGeom& g, int split, int n_tile);

int launch_conv_tc_phase(const TcPhase& ph, cudaStream_t stream);

struct DirectPhase : PhaseGeom {
  int cin;
  const float* x;              // fp32 NHWC (or NCHW if in_nchw)
  int in_nchw;
  const float* in_scale;       // [N][cin] or null: fused input transform act(x*scale+shift)
  const float* in_shift;
  int in_act;
  int pad_mode;                // DLB_PAD_ZERO / DLB_PAD_REFLECT
  const float* w;              // fp32 [tap][cin][cout]
  const float* bias;
  float* y;
  int out_act;
  int out_nchw;                // y is NCHW [N][cout][OHf][OWf]; ys_* then address the (h,w) plane
  long long ys_c;              // channel stride in elements (1 for NHWC)
};

int launch_conv_direct_phase(const DirectPhase& ph, cudaStream_t stream);

}  // namespace dlb

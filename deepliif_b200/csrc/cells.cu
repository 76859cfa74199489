// Cell post-processing on the stitched uint8 images: the integer graph work the reference runs
// in numba on the host (deepliif/postprocessing.py).
//   dlb_cells_posneg_mask     create_posneg_mask            postprocessing.py:163-190
//   dlb_cells_marker_plane    to_array(grayscale) / create_od_image + the histogram calculate_stain_range needs
//                                                           :98-120, 123-138, 450-469
//   dlb_cells_mark_background mark_background               :193-232
//   dlb_cells_label           the component search of compute_cell_mapping   :235-308
//   dlb_cells_stats           its per-cell counts / marker / centroid sums
//   dlb_cells_classify        create_cell_classification    :923-1000
//   dlb_cells_enlarge         enlarge_cell_boundaries       :1003-1030
//   dlb_cells_final_images    create_final_images           :1033-1071
//
// The reference's sequential sweeps and stack floods have order-independent results, which is what is computed here:
//   * mark_background = closure of BACKGROUND over 4-connected UNKNOWN pixels -> union-find components of the UNKNOWN
//     pixels, components that touch the image border (or an existing BACKGROUND pixel) become BACKGROUND;
//   * a cell = 8-connected component of the non-background pixels; the reference lists cells in raster order of their
//     first pixel, and a union-find that always links the larger root under the smaller one ends with root == first
//     pixel, so compacting the roots in index order gives the reference's cell order;
//   * classification / boundary growth are "first writer in list (raster) order wins" -> min over the candidates.
// All of it is byte/int traffic bound by HBM and atomics latency; no tensor-core work.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "internal.h"

namespace dlb {
namespace {

constexpr uint8_t kUnknown = 50, kPositive = 200, kNegative = 150, kBackground = 0, kCell = 100, kBorderPos = 220,
                  kBorderNeg = 170;

template <int MODE> __device__ __forceinline__ bool member(uint8_t v) {
  return MODE == 0 ? (v == kUnknown) : (v != kBackground && v != kCell);   // 0: flood region, 1: cell pixels
}

__device__ __forceinline__ int ld(const int* L, long long i) { return __ldcg(L + i); }

__device__ __forceinline__ int find_root(const int* L, int p) {
  int q = ld(L, p);
  while (q != p) { p = q; q = ld(L, p); }
  return p;
}

// Lock-free union (Komura-style): the larger root is linked under the smaller with atomicMin; when another thread got
// there first the displaced parent is carried on, so no equivalence is lost.
__device__ void unite(int* L, int a, int b) {
  bool done;
  do {
    a = find_root(L, a);
    b = find_root(L, b);
    if (a < b) { const int old = atomicMin(L + b, a); done = (old == b); b = old; }
    else if (b < a) { const int old = atomicMin(L + a, b); done = (old == a); a = old; }
    else done = true;
  } while (!done);
}

// block (32, 8): one warp = 32 consecutive pixels of one row.  Every member pixel starts linked to the first pixel of
// its run inside the warp's 32-pixel segment (chains of length 1, built from one ballot).
template <int MODE>
__global__ void __launch_bounds__(256) ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W) {
  const int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H) return;
  const bool m = x < W && member<MODE>(mask[static_cast<long long>(y) * W + x]);
  const unsigned b = __ballot_sync(0xffffffffu, m);
  if (x >= W) return;
  const long long p = static_cast<long long>(y) * W + x;
  if (!m) { L[p] = -1; return; }
  const unsigned zeros = ~b & ((1u << lane) - 1u);
  const int start = zeros ? 32 - __clz(zeros) : 0;
  L[p] = static_cast<int>(p - lane + start);
}

template <int MODE, int CONN>
__global__ void __launch_bounds__(256) ccl_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W) {
  const int lane = threadIdx.x, x = blockIdx.x * 32 + lane, y = blockIdx.y * 8 + threadIdx.y;
  if (y >= H || x >= W) return;
  const long long p = static_cast<long long>(y) * W + x;
  if (!member<MODE>(mask[p])) return;
  const bool w = x > 0 && member<MODE>(mask[p - 1]);
  if (lane == 0 && w) unite(L, static_cast<int>(p), static_cast<int>(p - 1));      // runs continue across segments
  if (y == 0) return;
  const bool n = member<MODE>(mask[p - W]);
  const bool nw = x > 0 && member<MODE>(mask[p - W - 1]);
  if (n) {
    if (!(w && nw)) unite(L, static_cast<int>(p), static_cast<int>(p - W));        // else W already joined NW == N's run
  } else if (CONN == 8) {
    if (nw && !w) unite(L, static_cast<int>(p), static_cast<int>(p - W - 1));      // with W present, W joins its own N
    if (x + 1 < W && member<MODE>(mask[p - W + 1])) unite(L, static_cast<int>(p), static_cast<int>(p - W + 1));
  }
}

__global__ void ccl_compress_kernel(int* __restrict__ L, unsigned total) {
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int v = L[p];
    if (v >= 0 && v != static_cast<int>(p)) L[p] = find_root(L, v);
  }
}

// ---- mark_background --------------------------------------------------------------------------------------------
// Roots of components that touch the border or an existing BACKGROUND pixel are overwritten with -2 (only root
// entries are ever rewritten, so the non-root entries other threads read stay valid).
__global__ void flood_flag_kernel(const uint8_t* __restrict__ mask, int* __restrict__ L, int H, int W) {
  const unsigned total = static_cast<unsigned>(H) * W;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    if (mask[p] != kUnknown) continue;
    const int y = static_cast<int>(p / static_cast<unsigned>(W)), x = static_cast<int>(p - static_cast<unsigned>(y) * W);
    bool seed = (y == 0 || x == 0 || y == H - 1 || x == W - 1);
    if (!seed) seed = mask[p - W] == kBackground || mask[p + W] == kBackground || mask[p - 1] == kBackground ||
                      mask[p + 1] == kBackground;
    if (!seed) continue;
    const int r = ld(L, p);
    if (r >= 0) L[r] = -2;
  }
}

__global__ void flood_apply_kernel(uint8_t* __restrict__ mask, const int* __restrict__ L, unsigned total) {
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int v = L[p];
    if (v == -1) continue;
    if (v == -2 || L[v] == -2) mask[p] = kBackground;
  }
}

// ---- cells --------------------------------------------------------------------------------------------------------
struct IsRoot {
  const int* L;
  __device__ bool operator()(const int& i) const { return L[i] == i; }
};

__global__ void root_code_kernel(int* __restrict__ L, const int* __restrict__ roots, const int* __restrict__ n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < *n; i += gridDim.x * blockDim.x) L[roots[i]] = -(i + 2);
}

__global__ void relabel_kernel(int* __restrict__ L, unsigned total) {          // -> component index, -1 elsewhere
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int v = L[p];
    if (v >= 0) L[p] = -(L[v] + 2);            // L[v] is a coded root entry: never rewritten by this kernel's non-roots
  }
}
__global__ void relabel_roots_kernel(int* __restrict__ L, const int* __restrict__ roots, const int* __restrict__ n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < *n; i += gridDim.x * blockDim.x) L[roots[i]] = i;
}

// table: int64 [n][8] = count, count_pos, count_neg, marker (max or sum), x0, y0, sum_x, sum_y.
// Pixels of one warp that belong to the same cell are combined first (match_any + redux), so a large cell costs one
// set of atomics per warp instead of one per pixel.
__global__ void __launch_bounds__(256) cell_stats_kernel(const uint8_t* __restrict__ mask, const uint16_t* __restrict__ marker,
                                                         const int* __restrict__ lab, int H, int W, int use_avg,
                                                         unsigned long long* __restrict__ table) {
  const unsigned total = static_cast<unsigned>(H) * W;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned base = blockIdx.x * blockDim.x; base < total; base += stride) {
    const unsigned p = base + threadIdx.x;
    const int c = p < total ? lab[p] : -1;
    const unsigned active = __ballot_sync(0xffffffffu, c >= 0);
    if (c < 0) continue;
    const unsigned peers = __match_any_sync(active, c);
    const uint8_t v = mask[p];
    const unsigned y = p / static_cast<unsigned>(W), x = p - y * W;
    const unsigned mv = marker ? marker[p] : 0u;
    const unsigned cnt = __popc(peers);
    const unsigned pos = __reduce_add_sync(peers, v == kPositive ? 1u : 0u);
    const unsigned neg = __reduce_add_sync(peers, v == kNegative ? 1u : 0u);
    const unsigned sx = __reduce_add_sync(peers, x);
    const unsigned sy = __reduce_add_sync(peers, y);
    const unsigned mk = use_avg ? __reduce_add_sync(peers, mv) : __reduce_max_sync(peers, mv);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
      unsigned long long* t = table + static_cast<long long>(c) * 8;
      atomicAdd(t + 0, static_cast<unsigned long long>(cnt));
      if (pos) atomicAdd(t + 1, static_cast<unsigned long long>(pos));
      if (neg) atomicAdd(t + 2, static_cast<unsigned long long>(neg));
      if (use_avg) atomicAdd(t + 3, static_cast<unsigned long long>(mk));
      else if (mk) atomicMax(t + 3, static_cast<unsigned long long>(mk));
      atomicAdd(t + 6, static_cast<unsigned long long>(sx));
      atomicAdd(t + 7, static_cast<unsigned long long>(sy));
    }
  }
}

__global__ void cell_first_kernel(const int* __restrict__ roots, int n, int W, unsigned long long* __restrict__ table) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    table[static_cast<long long>(i) * 8 + 4] = static_cast<unsigned long long>(roots[i] % W);
    table[static_cast<long long>(i) * 8 + 5] = static_cast<unsigned long long>(roots[i] / W);
  }
}

// cls[c]: 0 = cell not counted (stays LABEL_CELL), 1 = negative, 2 = positive.
__global__ void classify_kernel(const int* __restrict__ lab, const int* __restrict__ roots, const uint8_t* __restrict__ cls,
                                int H, int W, uint8_t* __restrict__ out) {
  const unsigned total = static_cast<unsigned>(H) * W;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int c = lab[p];
    if (c >= 0) {
      const uint8_t k = cls[c];
      if (k == 0) out[p] = kCell;
      else if (roots[c] == static_cast<int>(p)) out[p] = (k == 2) ? kBorderPos : kBorderNeg;     // the flood's start pixel keeps the border label
      else out[p] = (k == 2) ? kPositive : kNegative;
      continue;
    }
    const int y = static_cast<int>(p / static_cast<unsigned>(W)), x = static_cast<int>(p - static_cast<unsigned>(y) * W);
    int best = 0x7fffffff;
    auto look = [&](unsigned q) {
      const int cq = lab[q];
      if (cq >= 0 && cq < best && cls[cq] != 0 && roots[cq] != static_cast<int>(q)) best = cq;
    };
    if (y > 0) look(p - W);
    if (y + 1 < H) look(p + W);
    if (x > 0) look(p - 1);
    if (x + 1 < W) look(p + 1);
    out[p] = best == 0x7fffffff ? kBackground : (cls[best] == 2 ? kBorderPos : kBorderNeg);
  }
}

__global__ void enlarge_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W) {
  const unsigned total = static_cast<unsigned>(H) * W;
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    uint8_t v = in[p];
    if (v == kBackground) {
      const int y = static_cast<int>(p / static_cast<unsigned>(W)), x = static_cast<int>(p - static_cast<unsigned>(y) * W);
      bool found = false;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          if ((dy == 0 && dx == 0) || found) continue;
          const int yy = y + dy, xx = x + dx;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          const uint8_t nb = in[static_cast<unsigned>(yy) * W + xx];
          if (nb == kBorderPos || nb == kBorderNeg) { v = nb; found = true; }
        }
    }
    out[p] = v;
  }
}

__device__ __forceinline__ void final_pixel(uint8_t m, uint8_t* o, uint8_t* r) {
  r[0] = r[1] = r[2] = 0;
  if (m == kBorderPos) { o[0] = 255; o[1] = 0; o[2] = 0; r[1] = 255; }
  else if (m == kBorderNeg) { o[0] = 0; o[1] = 0; o[2] = 255; r[1] = 255; }
  else if (m == kPositive) r[0] = 255;
  else if (m == kNegative) r[2] = 255;
}

// 4 pixels per thread: one 32-bit mask word, three 32-bit words of each RGB image (all buffers 4-byte aligned).
__global__ void final_images_kernel(const uint8_t* __restrict__ orig, const uint8_t* __restrict__ mask, unsigned total,
                                    uint8_t* __restrict__ overlay, uint8_t* __restrict__ refined) {
  const unsigned quads = total / 4;
  for (unsigned q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
    union { uint32_t w; uint8_t b[4]; } m;
    union { uint32_t w[3]; uint8_t b[12]; } o, r;
    m.w = reinterpret_cast<const uint32_t*>(mask)[q];
    const uint32_t* src = reinterpret_cast<const uint32_t*>(orig) + static_cast<size_t>(q) * 3;
    o.w[0] = src[0]; o.w[1] = src[1]; o.w[2] = src[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) final_pixel(m.b[k], o.b + 3 * k, r.b + 3 * k);
    uint32_t* po = reinterpret_cast<uint32_t*>(overlay) + static_cast<size_t>(q) * 3;
    uint32_t* pr = reinterpret_cast<uint32_t*>(refined) + static_cast<size_t>(q) * 3;
    po[0] = o.w[0]; po[1] = o.w[1]; po[2] = o.w[2];
    pr[0] = r.w[0]; pr[1] = r.w[1]; pr[2] = r.w[2];
  }
  if (blockIdx.x == 0 && threadIdx.x < total % 4) {                       // tail pixels
    const size_t p = static_cast<size_t>(quads) * 4 + threadIdx.x;
    uint8_t o[3] = {orig[p * 3], orig[p * 3 + 1], orig[p * 3 + 2]}, r[3];
    final_pixel(mask[p], o, r);
    for (int c = 0; c < 3; ++c) { overlay[p * 3 + c] = o[c]; refined[p * 3 + c] = r[c]; }
  }
}

__global__ void posneg_mask_kernel(const uint8_t* __restrict__ seg, unsigned total, int thresh, uint8_t* __restrict__ mask) {
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const size_t o = static_cast<size_t>(p) * 3;
    const int r = seg[o], g = seg[o + 1], b = seg[o + 2];
    uint8_t m = kUnknown;
    if (r + b > thresh && g <= 80) m = (r >= b) ? kPositive : kNegative;
    mask[p] = m;
  }
}

// mode 0: max over the channels (+ histogram of the non-zero values); mode 1: optical density, the reference's
// round(100 * (lut[r] + lut[g] + lut[b])) with the caller's 256-entry float64 LUT (left-to-right adds, ties to even).
__global__ void __launch_bounds__(256) marker_plane_kernel(const uint8_t* __restrict__ img, unsigned total, int mode,
                                                           const double* __restrict__ lut, uint16_t* __restrict__ out,
                                                           unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[256];
  if (hist) { sh[threadIdx.x] = 0; __syncthreads(); }
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const size_t o = static_cast<size_t>(p) * 3;
    const int r = img[o], g = img[o + 1], b = img[o + 2];
    if (mode == 0) {
      const int v = max(r, max(g, b));
      out[p] = static_cast<uint16_t>(v);
      if (hist && v) atomicAdd(&sh[v], 1u);
    } else {
      const double val = __dadd_rn(__dadd_rn(lut[r], lut[g]), lut[b]);
      out[p] = static_cast<uint16_t>(rint(__dmul_rn(val, 100.0)));
    }
  }
  if (hist) {
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
  }
}

int grid_for(long long total) {
  long long g = (total + 255) / 256;
  return static_cast<int>(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

int check_dims(const char* who, int H, int W) {
  if (H < 1 || W < 1 || static_cast<long long>(H) * W >= (1ll << 31)) {
    static thread_local char buf[128];
    snprintf(buf, sizeof(buf), "%s: need 1 <= H*W < 2^31", who);
    return set_error(buf);
  }
  return 0;
}

template <int MODE, int CONN> int run_ccl(const uint8_t* mask, int* L, int H, int W, cudaStream_t s) {
  const dim3 grid((W + 31) / 32, (H + 7) / 8), block(32, 8);
  ccl_init_kernel<MODE><<<grid, block, 0, s>>>(mask, L, H, W);
  ccl_merge_kernel<MODE, CONN><<<grid, block, 0, s>>>(mask, L, H, W);
  ccl_compress_kernel<<<grid_for(static_cast<long long>(H) * W), 256, 0, s>>>(L, static_cast<long long>(H) * W);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("ccl kernels launch");
}

}  // namespace
}  // namespace dlb

using namespace dlb;

extern "C" int dlb_cells_posneg_mask(const uint8_t* seg_hwc, int H, int W, int thresh, uint8_t* mask, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_posneg_mask", H, W)) return rc;
  const long long total = static_cast<long long>(H) * W;
  posneg_mask_kernel<<<grid_for(total), 256, 0, stream>>>(seg_hwc, total, thresh, mask);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("posneg_mask_kernel launch");
}

extern "C" int dlb_cells_marker_plane(const uint8_t* img_hwc, int H, int W, int mode, const double* od_lut,
                                      uint16_t* plane, unsigned int* hist256, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_marker_plane", H, W)) return rc;
  if (mode != 0 && mode != 1) return set_error("dlb_cells_marker_plane: mode 0 (max) or 1 (optical density)");
  if (mode == 1 && !od_lut) return set_error("dlb_cells_marker_plane: optical density needs the 256-entry LUT");
  if (hist256 && cudaMemsetAsync(hist256, 0, 256 * sizeof(unsigned int), stream) != cudaSuccess) return set_cuda_error("memset");
  const long long total = static_cast<long long>(H) * W;
  marker_plane_kernel<<<grid_for(total), 256, 0, stream>>>(img_hwc, total, mode, od_lut, plane, mode == 0 ? hist256 : nullptr);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("marker_plane_kernel launch");
}

extern "C" int dlb_cells_mark_background(uint8_t* mask, int H, int W, int* labels_ws, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_mark_background", H, W)) return rc;
  if (int rc = run_ccl<0, 4>(mask, labels_ws, H, W, stream)) return rc;
  const long long total = static_cast<long long>(H) * W;
  flood_flag_kernel<<<grid_for(total), 256, 0, stream>>>(mask, labels_ws, H, W);
  flood_apply_kernel<<<grid_for(total), 256, 0, stream>>>(mask, labels_ws, total);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("flood kernels launch");
}

extern "C" size_t dlb_cells_label_workspace(int H, int W) {
  size_t bytes = 0;
  thrust::counting_iterator<int> it(0);
  cub::DeviceSelect::If(nullptr, bytes, it, static_cast<int*>(nullptr), static_cast<int*>(nullptr), H * W, IsRoot{nullptr});
  return bytes + 256;
}

extern "C" int dlb_cells_label(const uint8_t* mask, int H, int W, int* labels, int* roots, int* n_cells, void* ws,
                               size_t ws_bytes, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_label", H, W)) return rc;
  if (int rc = run_ccl<1, 8>(mask, labels, H, W, stream)) return rc;
  size_t need = 0;
  thrust::counting_iterator<int> it(0);
  cub::DeviceSelect::If(nullptr, need, it, roots, n_cells, H * W, IsRoot{labels}, stream);
  if (ws_bytes < need) return set_error("dlb_cells_label: workspace too small (dlb_cells_label_workspace)");
  if (cub::DeviceSelect::If(ws, need, it, roots, n_cells, H * W, IsRoot{labels}, stream) != cudaSuccess)
    return set_cuda_error("cub::DeviceSelect::If");
  const long long total = static_cast<long long>(H) * W;
  root_code_kernel<<<148, 256, 0, stream>>>(labels, roots, n_cells);
  relabel_kernel<<<grid_for(total), 256, 0, stream>>>(labels, total);
  relabel_roots_kernel<<<148, 256, 0, stream>>>(labels, roots, n_cells);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("relabel kernels launch");
}

extern "C" int dlb_cells_stats(const uint8_t* mask, const uint16_t* marker, const int* labels, const int* roots, int n,
                               int H, int W, int use_avg, long long* table, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_stats", H, W)) return rc;
  if (n < 0) return set_error("dlb_cells_stats: negative cell count");
  if (n == 0) return 0;
  if (cudaMemsetAsync(table, 0, sizeof(long long) * 8 * n, stream) != cudaSuccess) return set_cuda_error("memset");
  auto* t = reinterpret_cast<unsigned long long*>(table);
  cell_stats_kernel<<<grid_for(static_cast<long long>(H) * W), 256, 0, stream>>>(mask, marker, labels, H, W, use_avg, t);
  cell_first_kernel<<<grid_for(n), 256, 0, stream>>>(roots, n, W, t);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("cell stats kernels launch");
}

extern "C" int dlb_cells_classify(const int* labels, const int* roots, const uint8_t* cls, int H, int W, uint8_t* mask_out,
                                  dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_classify", H, W)) return rc;
  classify_kernel<<<grid_for(static_cast<long long>(H) * W), 256, 0, stream>>>(labels, roots, cls, H, W, mask_out);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("classify_kernel launch");
}

extern "C" int dlb_cells_enlarge(const uint8_t* mask_in, uint8_t* mask_out, int H, int W, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_enlarge", H, W)) return rc;
  if (mask_in == mask_out) return set_error("dlb_cells_enlarge: out of place only");
  enlarge_kernel<<<grid_for(static_cast<long long>(H) * W), 256, 0, stream>>>(mask_in, mask_out, H, W);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("enlarge_kernel launch");
}

extern "C" int dlb_cells_final_images(const uint8_t* orig_hwc, const uint8_t* mask, int H, int W, uint8_t* overlay_hwc,
                                      uint8_t* refined_hwc, dlb_stream_t stream) {
  if (int rc = check_dims("dlb_cells_final_images", H, W)) return rc;
  const long long total = static_cast<long long>(H) * W;
  final_images_kernel<<<grid_for(total), 256, 0, stream>>>(orig_hwc, mask, total, overlay_hwc, refined_hwc);
  return cudaGetLastError() == cudaSuccess ? 0 : set_cuda_error("final_images_kernel launch");
}

"""Cell post-processing of the stitched Seg / Marker images on the GPU — host mirror of the reference's
deepliif/postprocessing.py for the path `infer_modalities -> postprocess -> compute_final_results`
(models/__init__.py:582-611, postprocessing.py:1223-1304).  Same function names, arguments, return values and scoring
keys; the pixel/graph work runs in libdeepliif_b200.so (csrc/cells.cu), the per-cell threshold logic (a few hundred
cells) stays on the host exactly as the reference has it.  There is no CPU fallback: without the CUDA library these
functions raise.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import check
from .ops import LAUNCHES, _p, _stream

DEFAULT_SEG_THRESH = 120
DEFAULT_NOISE_THRESH = 4
LABEL_UNKNOWN, LABEL_POSITIVE, LABEL_NEGATIVE, LABEL_BACKGROUND, LABEL_CELL = 50, 200, 150, 0, 100
LABEL_BORDER_POS, LABEL_BORDER_NEG = 220, 170


def to_array(img, grayscale=False):
    """postprocessing.py:98-120."""
    if not isinstance(img, np.ndarray):
        img = np.asarray(img) if img.mode == "RGB" else np.asarray(img.convert("RGB"))
    if grayscale and img.ndim == 3:
        img = img.max(axis=-1)
    return img


def _dev(a, device):
    import warnings
    with warnings.catch_warnings():                 # np.asarray(PIL image) is read-only; it is only read here
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=False)


def _to_host(t):
    """D2H through pinned memory (torch's caching host allocator keeps the buffer for the next call)."""
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return out.numpy()


def _device():
    if not torch.cuda.is_available():
        raise _lib.DeepliifB200Error("deepliif_b200.postprocessing needs a CUDA device (no CPU path exists)")
    return torch.device("cuda", torch.cuda.current_device())


# ---- device stages (thin wrappers over the C ABI) ---------------------------------------------------------------
def create_posneg_mask_gpu(seg_d, thresh):
    H, W, _ = seg_d.shape
    mask = torch.empty((H, W), dtype=torch.uint8, device=seg_d.device)
    check(_lib.load().dlb_cells_posneg_mask(_p(seg_d), H, W, int(thresh), _p(mask), _stream()), "dlb_cells_posneg_mask")
    LAUNCHES["count"] += 1
    return mask


_OD_LUT = None


def od_lut():
    """create_od_image's LUT (postprocessing.py:125-129) as float64."""
    global _OD_LUT
    if _OD_LUT is None:
        lut = [0.0] + [math.log10(255 / i) for i in range(1, 256)]
        lut[0] = lut[1]
        _OD_LUT = np.asarray(lut, dtype=np.float64)
    return _OD_LUT


def marker_plane_gpu(img_d, use_od=False, want_hist=False):
    """uint8 [H,W,3] -> uint16 [H,W] (max over channels, or optical density) [+ histogram of non-zero values]."""
    H, W, _ = img_d.shape
    plane = torch.empty((H, W), dtype=torch.uint16, device=img_d.device)
    hist = torch.empty(256, dtype=torch.int32, device=img_d.device) if (want_hist and not use_od) else None
    lut = _dev(od_lut(), img_d.device) if use_od else None
    check(_lib.load().dlb_cells_marker_plane(_p(img_d), H, W, 1 if use_od else 0, _p(lut), _p(plane), _p(hist), _stream()),
          "dlb_cells_marker_plane")
    LAUNCHES["count"] += 1
    return plane, hist


def mark_background_gpu(mask_d, labels_ws=None):
    H, W = mask_d.shape
    if labels_ws is None:
        labels_ws = torch.empty(H * W, dtype=torch.int32, device=mask_d.device)
    check(_lib.load().dlb_cells_mark_background(_p(mask_d), H, W, _p(labels_ws), _stream()), "dlb_cells_mark_background")
    LAUNCHES["count"] += 5
    return mask_d


def label_cells_gpu(mask_d, labels=None):
    """-> (labels int32 [H,W], roots int32 [n], n)."""
    H, W = mask_d.shape
    lib = _lib.load()
    if labels is None:
        labels = torch.empty(H * W, dtype=torch.int32, device=mask_d.device)
    cap = ((H + 1) // 2) * ((W + 1) // 2)
    roots = torch.empty(cap, dtype=torch.int32, device=mask_d.device)
    n_d = torch.zeros(1, dtype=torch.int32, device=mask_d.device)
    wsb = lib.dlb_cells_label_workspace(H, W)
    ws = torch.empty(wsb, dtype=torch.uint8, device=mask_d.device)
    check(lib.dlb_cells_label(_p(mask_d), H, W, _p(labels), _p(roots), _p(n_d), _p(ws), C.c_size_t(wsb), _stream()),
          "dlb_cells_label")
    LAUNCHES["count"] += 8
    n = int(n_d.item())
    return labels.view(H, W), roots[:n], n


def cell_stats_gpu(mask_d, marker_plane, labels, roots, use_avg=False):
    H, W = mask_d.shape
    n = roots.numel()
    table = torch.empty((n, 8), dtype=torch.int64, device=mask_d.device)
    check(_lib.load().dlb_cells_stats(_p(mask_d), _p(marker_plane), _p(labels), _p(roots), n, H, W, int(bool(use_avg)),
                                      _p(table), _stream()), "dlb_cells_stats")
    LAUNCHES["count"] += 2
    return table


def classify_gpu(labels, roots, cls_d, H, W):
    out = torch.empty((H, W), dtype=torch.uint8, device=labels.device)
    check(_lib.load().dlb_cells_classify(_p(labels), _p(roots), _p(cls_d), H, W, _p(out), _stream()), "dlb_cells_classify")
    LAUNCHES["count"] += 1
    return out


def enlarge_cell_boundaries_gpu(mask_d):
    H, W = mask_d.shape
    out = torch.empty_like(mask_d)
    check(_lib.load().dlb_cells_enlarge(_p(mask_d), _p(out), H, W, _stream()), "dlb_cells_enlarge")
    LAUNCHES["count"] += 1
    return out


def create_final_images_gpu(orig_d, mask_d):
    H, W = mask_d.shape
    overlay = torch.empty_like(orig_d)
    refined = torch.empty_like(orig_d)
    check(_lib.load().dlb_cells_final_images(_p(orig_d), _p(mask_d), H, W, _p(overlay), _p(refined), _stream()),
          "dlb_cells_final_images")
    LAUNCHES["count"] += 1
    return overlay, refined


# ---- host logic over the cell list (postprocessing.py:365-488, 1125-1133) --------------------------------------
def _round_div_half_even(num, den):
    """int(round(num / den)) for non-negative int64 arrays with python3 / numba semantics (ties to even)."""
    q, r = np.divmod(num, np.maximum(den, 1))
    return q + ((2 * r > den) | ((2 * r == den) & (q % 2 == 1)))


class CellTable:
    """The reference's cellsinfo list (7-tuples: count, positive, marker, x0, y0, cx, cy) held as columns.  The
    centroids are only materialised when the list is asked for (compute_final_results never reads them)."""
    __slots__ = ("count", "positive", "marker", "x0", "y0", "sum_x", "sum_y")

    def __len__(self):
        return int(self.count.shape[0])

    @property
    def cx(self):
        return _round_div_half_even(self.sum_x, self.count)

    @property
    def cy(self):
        return _round_div_half_even(self.sum_y, self.count)

    def as_tuples(self):
        return list(zip(self.count.tolist(), self.positive.tolist(), self.marker.tolist(), self.x0.tolist(),
                        self.y0.tolist(), self.cx.tolist(), self.cy.tolist()))


def _kde_bins(uniq, mult, n, step, b0, b1, bandwidth=1.0):
    """Bins b0..b1-1 of the reference's Gaussian KDE (postprocessing.py:365-403; float64 sums stored as float32).  Cell
    sizes repeat, so the kernel is evaluated once per distinct value and weighted by its multiplicity."""
    c = 1 / math.sqrt(2 * math.pi)
    x = (np.arange(b0, b1, dtype=np.float64) * step)[:, None]
    val = (x - uniq[None, :]) * (1 / bandwidth)
    total = (np.exp(-(val * val / 2)) * c) @ mult
    return (total / (n * bandwidth)).astype(np.float32)


def create_kde(values, count, bandwidth=1.0):
    """postprocessing.py:365-403 -> (kde float32 [count], step)."""
    step = (float(values.max()) + 1) / count
    uniq, mult = np.unique(values, return_counts=True)
    return _kde_bins(uniq, mult.astype(np.float64), values.shape[0], step, 0, count, bandwidth), step


def calculate_default_size_threshold(cell_sizes, resolution="40x"):
    """postprocessing.py:406-447.  Only the first interior local minimum of the KDE is used, so the bins are evaluated in
    blocks from the left and the evaluation stops at the block that contains it (same index as scanning the full KDE)."""
    cell_sizes = np.asarray(cell_sizes, dtype=np.int64)
    if cell_sizes.shape[0] <= 1:
        return 0
    values = np.sqrt(cell_sizes)
    count = 500
    step = (float(values.max()) + 1) / count
    uniq, mult = np.unique(values, return_counts=True)
    mult = mult.astype(np.float64)
    idx, kde, block = 1, np.zeros(0, np.float32), 64
    for b0 in range(0, count, block):
        kde = np.concatenate([kde, _kde_bins(uniq, mult, values.shape[0], step, b0, min(count, b0 + block))])
        m = kde.shape[0]                                   # interior bins 1 .. m-2 are decidable now
        interior = (kde[1:m - 1] < kde[:m - 2]) & (kde[1:m - 1] < kde[2:m])
        hits = np.flatnonzero(interior)
        if hits.size:
            idx = int(hits[0]) + 1
            break
    thresh_sqrt = (idx - 1) * step
    lo, default, hi = {"20x": (3, 4, 6), "10x": (2, 2, 3)}.get(resolution, (4, 7, 10))
    if thresh_sqrt < lo:
        thresh_sqrt = lo
    elif thresh_sqrt > hi:
        thresh_sqrt = default
    return int(round(thresh_sqrt * thresh_sqrt))


def _percentile_from_hist(hist, q):
    """np.percentile(values, q) (method 'linear') for uint8 values given as a histogram: same virtual index and
    the same two-sided lerp numpy uses."""
    n = int(hist.sum())
    cum = np.cumsum(hist)
    vi = n * (q / 100.0) + (1 + (q / 100.0) * (1 - 1 - 1)) - 1
    prev = int(math.floor(vi))
    nxt = min(prev + 1, n - 1)
    prev = max(prev, 0)
    g = vi - math.floor(vi)
    a = float(np.searchsorted(cum, prev + 1, side="left"))          # value of the prev-th order statistic (0-based)
    b = float(np.searchsorted(cum, nxt + 1, side="left"))
    d = b - a
    return (a + d * g) if g < 0.5 else (b - d * (1 - g))


def calculate_stain_range_from_hist(hist):
    """postprocessing.py:450-469 on the histogram of the non-zero marker values."""
    hist = np.asarray(hist, dtype=np.int64).copy()
    hist[0] = 0
    if hist.sum() > 0:
        return (round(_percentile_from_hist(hist, 0.1)), round(_percentile_from_hist(hist, 99.9)))
    return (0, 0)


def calculate_default_marker_threshold_from_hist(hist):
    """postprocessing.py:472-488."""
    lo, hi = calculate_stain_range_from_hist(hist)
    return round((hi - lo) * 0.9) + lo


def calculate_large_noise_thresh(large_noise_thresh, resolution):
    if not (isinstance(large_noise_thresh, str) and large_noise_thresh == "default"):
        return large_noise_thresh
    return {"10x": 1000, "20x": 4000}.get(resolution, 16000)


# ---- the reference's public functions ----------------------------------------------------------------------------
class CellState:
    """What get_cells_info leaves on the device for create_cell_classification (the reference keeps x0,y0 per cell
    and re-floods; here the label image does that job)."""
    __slots__ = ("mask", "labels", "roots", "kept", "H", "W")


class StageClock:
    """Optional per-stage device timing (CUDA events on the current stream) for bench.py."""

    def __init__(self):
        self.marks = [("start", self._ev())]

    @staticmethod
    def _ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def mark(self, name):
        self.marks.append((name, self._ev()))

    def ms(self):
        torch.cuda.synchronize()
        return {n: self.marks[i][1].elapsed_time(e) for i, (n, e) in enumerate(self.marks[1:])}


def _cells_device(seg_d, marker_d, resolution, noise_thresh, seg_thresh, large_noise_thresh, use_od, st, clock=None):
    """Device part of get_cells_info: uint8 [H,W,3] device tensors -> (cells, defaults); label state left in `st`."""
    mark = clock.mark if clock is not None else (lambda name: None)
    H, W, _ = seg_d.shape
    plane = hist = None
    if marker_d is not None:
        plane, hist = marker_plane_gpu(marker_d, use_od=use_od, want_hist=not use_od)
    mask_d = create_posneg_mask_gpu(seg_d, seg_thresh)
    mark("mask+marker")
    labels = torch.empty(H * W, dtype=torch.int32, device=seg_d.device)
    mark_background_gpu(mask_d, labels)
    mark("mark_background")
    labels, roots, n = label_cells_gpu(mask_d, labels)
    mark("label")
    table = cell_stats_gpu(mask_d, plane, labels, roots, use_avg=use_od).cpu().numpy()
    mark("stats+d2h")
    keep = table[:, 0] > noise_thresh
    if large_noise_thresh is not None:
        keep &= table[:, 0] < large_noise_thresh
    kept = np.flatnonzero(keep)
    t = table[kept]
    cnt = t[:, 0]
    cells = CellTable()
    cells.count, cells.positive = cnt, t[:, 1] >= t[:, 2]
    cells.marker = _round_div_half_even(t[:, 3], cnt) if use_od else t[:, 3]
    cells.x0, cells.y0 = t[:, 4], t[:, 5]
    cells.sum_x, cells.sum_y = t[:, 6], t[:, 7]
    defaults = {"size_thresh": calculate_default_size_threshold(cells.count, resolution)}
    if marker_d is not None and not use_od:
        defaults["marker_thresh"] = calculate_default_marker_threshold_from_hist(hist.cpu().numpy())
    mark("host thresholds")
    st.mask, st.labels, st.roots, st.kept, st.H, st.W = mask_d, labels, roots, kept, H, W
    return cells, defaults


def get_cells_info(seg, marker, resolution, noise_thresh, seg_thresh, large_noise_thresh, use_od=False):
    """postprocessing.py:311-362.  Returns (mask uint8 [H,W] with cells = 100, cellsinfo list of 7-tuples, defaults)."""
    dev = _device()
    st = CellState()
    cells, defaults = _cells_device(_dev(to_array(seg), dev), _dev(to_array(marker), dev) if marker is not None else None,
                                    resolution, noise_thresh, seg_thresh, large_noise_thresh, use_od, st)
    mask = torch.where(st.labels >= 0, torch.full_like(st.mask, LABEL_CELL), st.mask)
    return mask.cpu().numpy(), cells.as_tuples(), defaults


def _classes(cells, kept, n, size_thresh, marker_thresh, size_thresh_upper, od_thresh_lower, od_thresh_upper):
    """The per-cell decisions of create_cell_classification (postprocessing.py:958-980), on the columns."""
    cls = np.zeros(max(n, 1), dtype=np.uint8)
    sel = cells.count > size_thresh
    if size_thresh_upper is not None:
        sel &= cells.count < size_thresh_upper
    is_pos = cells.positive.copy()
    if marker_thresh is not None:
        is_pos |= cells.marker > marker_thresh
    low = cells.marker < od_thresh_lower if od_thresh_lower is not None else np.zeros(len(cells), dtype=bool)
    up = (cells.marker > od_thresh_upper) & ~low if od_thresh_upper is not None else np.zeros(len(cells), dtype=bool)
    is_pos &= ~(low | up)
    cls[kept[sel]] = np.where(is_pos[sel], 2, 1)
    num_pos, num_neg = int((sel & is_pos).sum()), int((sel & ~is_pos).sum())
    return cls, {"num_total": num_pos + num_neg, "num_pos": num_pos, "num_neg": num_neg}


def compute_final_results_device(orig_d, seg_d, marker_d, resolution, size_thresh="default", marker_thresh=None,
                                 size_thresh_upper=None, seg_thresh=DEFAULT_SEG_THRESH, noise_thresh=DEFAULT_NOISE_THRESH,
                                 large_noise_thresh=None, od_thresh_lower=None, od_thresh_upper=None, clock=None):
    """compute_final_results on uint8 [H,W,3] CUDA tensors -> (overlay_d, refined_d, scoring, mask_d, cells)."""
    large_noise_thresh = calculate_large_noise_thresh(large_noise_thresh, resolution)
    use_od = od_thresh_lower is not None or od_thresh_upper is not None
    st = CellState()
    cells, defaults = _cells_device(seg_d, orig_d if use_od else marker_d, resolution, noise_thresh, seg_thresh,
                                    large_noise_thresh, use_od, st, clock)
    if size_thresh is None:
        size_thresh = 0
    elif isinstance(size_thresh, str) and size_thresh == "default":
        size_thresh = defaults["size_thresh"]
    if isinstance(marker_thresh, str) and marker_thresh == "default":
        marker_thresh = defaults["marker_thresh"]
    cls, counts = _classes(cells, st.kept, st.roots.numel(), size_thresh, marker_thresh, size_thresh_upper,
                           od_thresh_lower, od_thresh_upper)
    mask_d = classify_gpu(st.labels, st.roots, _dev(cls, seg_d.device), st.H, st.W)
    mask_d = enlarge_cell_boundaries_gpu(enlarge_cell_boundaries_gpu(mask_d))
    overlay, refined = create_final_images_gpu(orig_d, mask_d)
    if clock is not None:
        clock.mark("classify+enlarge+images")
    scoring = {
        "num_total": counts["num_total"], "num_pos": counts["num_pos"], "num_neg": counts["num_neg"],
        "percent_pos": round(counts["num_pos"] / counts["num_total"] * 100, 1) if counts["num_pos"] > 0 else 0,
        "seg_thresh": seg_thresh, "size_thresh": size_thresh, "size_thresh_upper": size_thresh_upper,
        "marker_thresh": marker_thresh if marker_d is not None else None,
    }
    return overlay, refined, scoring, mask_d, cells


def compute_final_results(orig, seg, marker, resolution, size_thresh="default", marker_thresh=None,
                          size_thresh_upper=None, seg_thresh=DEFAULT_SEG_THRESH, noise_thresh=DEFAULT_NOISE_THRESH,
                          large_noise_thresh=None, od_thresh_lower=None, od_thresh_upper=None, return_mask=False):
    """postprocessing.py:1223-1304 -> (overlay uint8 [H,W,3], refined uint8 [H,W,3], scoring dict)."""
    dev = _device()
    overlay, refined, scoring, mask_d, cells = compute_final_results_device(
        _dev(to_array(orig), dev), _dev(to_array(seg), dev), _dev(to_array(marker), dev) if marker is not None else None,
        resolution, size_thresh, marker_thresh, size_thresh_upper, seg_thresh, noise_thresh, large_noise_thresh,
        od_thresh_lower, od_thresh_upper)
    out = (_to_host(overlay), _to_host(refined), scoring)
    return out + (mask_d.cpu().numpy(), cells.as_tuples()) if return_mask else out

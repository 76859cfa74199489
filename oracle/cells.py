"""Oracle: numpy/scipy restatement of the reference's cell post-processing (TEST INFRASTRUCTURE).

Nothing in the product imports this module; tests, `__graft_entry__.smoke()` and bench.py's CPU baseline do.
Parity is pinned: `oracle/gen_golden.py` runs the reference's own numba functions (deepliif/postprocessing.py) in the
build container and commits their outputs as `tests/golden/cells.npz`; `tests/test_oracle_cpu.py` checks this
restatement against them stage by stage.

  mark_background             /root/reference/deepliif/postprocessing.py:193-232
      border UNKNOWN pixels -> BACKGROUND, then BACKGROUND grows through 4-connected UNKNOWN pixels until it
      stops.  The fixed point does not depend on the sweep order, so it is stated as connected components.
  compute_cell_mapping        :235-308   8-connected components of the non-background pixels in raster order of
      their first pixel; per cell (count, positive, marker, x0, y0, cx, cy); noise filter; mask -> LABEL_CELL
  get_cells_info              :311-362
  create_kde / calculate_default_size_threshold   :365-447
  calculate_stain_range / calculate_default_marker_threshold  :450-488
  create_od_image             :123-138
  create_cell_classification  :923-1000  (first pixel of a kept cell gets the border label; BACKGROUND pixels
      4-adjacent to any other pixel of a kept cell get that cell's border label, first cell in list order wins)
  enlarge_cell_boundaries     :1003-1030 (in-place raster scan == "first border pixel in raster order among the 8
      neighbours wins")
  create_final_images         :1033-1071
  compute_final_results       :1223-1304 (scoring dict keys and rounding)
"""
import math

import numpy as np
from scipy import ndimage

from .pixel import create_posneg_mask

DEFAULT_SEG_THRESH = 120
DEFAULT_NOISE_THRESH = 4
LABEL_UNKNOWN = 50
LABEL_POSITIVE = 200
LABEL_NEGATIVE = 150
LABEL_BACKGROUND = 0
LABEL_CELL = 100
LABEL_BORDER_POS = 220
LABEL_BORDER_NEG = 170

_FOUR = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)
_EIGHT = np.ones((3, 3), dtype=bool)


def _shift(a, dy, dx, fill):
    """out[y, x] = a[y + dy, x + dx], `fill` outside the image."""
    H, W = a.shape
    out = np.full_like(a, fill)
    ys, yd = (slice(dy, H), slice(0, H - dy)) if dy >= 0 else (slice(0, H + dy), slice(-dy, H))
    xs, xd = (slice(dx, W), slice(0, W - dx)) if dx >= 0 else (slice(0, W + dx), slice(-dx, W))
    out[yd, xd] = a[ys, xs]
    return out


def mark_background(mask: np.ndarray) -> np.ndarray:
    """Returns a new mask (the reference works in place)."""
    out = mask.copy()
    unk = mask == LABEL_UNKNOWN
    lab, _ = ndimage.label(unk, structure=_FOUR)
    seed = np.zeros_like(unk)
    seed[0, :] = seed[-1, :] = seed[:, 0] = seed[:, -1] = True
    bg = mask == LABEL_BACKGROUND
    for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        seed |= _shift(bg, dy, dx, False)
    ids = np.unique(lab[seed & unk])
    out[np.isin(lab, ids) & unk] = LABEL_BACKGROUND
    return out


def _round_half_even_div(num: np.ndarray, den: np.ndarray) -> np.ndarray:
    """int(round(num / den)) for non-negative integers, python3 / numba semantics (ties to even)."""
    q, r = np.divmod(num, den)
    up = (2 * r > den) | ((2 * r == den) & (q % 2 == 1))
    return q + up


def components(mask: np.ndarray, marker=None, use_avg=False):
    """All 8-connected components of pixels that are neither BACKGROUND nor CELL, in raster order of their first
    pixel.  Returns (lab, table): lab int32 [H,W] 1-based component index in that order (0 elsewhere), table int64
    [n, 8] = count, count_pos, count_neg, marker(max or sum), x0, y0, sum_x, sum_y."""
    H, W = mask.shape
    fg = (mask != LABEL_BACKGROUND) & (mask != LABEL_CELL)
    lab, n = ndimage.label(fg, structure=_EIGHT)
    flat = lab.ravel()
    idx = np.flatnonzero(flat)
    first = np.full(n + 1, H * W, dtype=np.int64)
    np.minimum.at(first, flat[idx], idx)
    order = np.argsort(first[1:], kind="stable")           # scipy already numbers in raster order; do not rely on it
    remap = np.zeros(n + 1, dtype=np.int32)
    remap[order + 1] = np.arange(1, n + 1, dtype=np.int32)
    lab = remap[lab]
    flat = lab.ravel()
    l = flat[idx]
    ys, xs = np.divmod(idx, W)
    t = np.zeros((n + 1, 8), dtype=np.int64)
    t[:, 0] = np.bincount(l, minlength=n + 1)
    m = mask.ravel()[idx]
    t[:, 1] = np.bincount(l, weights=(m == LABEL_POSITIVE), minlength=n + 1)
    t[:, 2] = np.bincount(l, weights=(m == LABEL_NEGATIVE), minlength=n + 1)
    if marker is not None:
        mv = marker.ravel()[idx].astype(np.int64)
        if use_avg:
            np.add.at(t[:, 3], l, mv)
        else:
            np.maximum.at(t[:, 3], l, mv)
    f = np.sort(first[1:])
    t[1:, 5], t[1:, 4] = np.divmod(f, W)
    np.add.at(t[:, 6], l, xs)
    np.add.at(t[:, 7], l, ys)
    return lab, t[1:]


def compute_cell_mapping(mask, marker, noise_thresh, large_noise_thresh, use_avg=False):
    """-> (mask with every cell pixel = LABEL_CELL, list of 7-tuples, lab, kept component indices (0-based))."""
    lab, t = components(mask, marker, use_avg)
    out = mask.copy()
    out[lab > 0] = LABEL_CELL
    cells, kept = [], []
    for i, (cnt, cp, cn, mv, x0, y0, sx, sy) in enumerate(t.tolist()):
        if cnt > noise_thresh and (large_noise_thresh is None or cnt < large_noise_thresh):
            cy = int(_round_half_even_div(np.int64(sy), np.int64(cnt)))
            cx = int(_round_half_even_div(np.int64(sx), np.int64(cnt)))
            if use_avg:
                mv = int(_round_half_even_div(np.int64(mv), np.int64(cnt)))
            cells.append((cnt, cp >= cn, mv, x0, y0, cx, cy))
            kept.append(i)
    return out, cells, lab, np.asarray(kept, dtype=np.int64)


def create_od_image(orig: np.ndarray) -> np.ndarray:
    lut = [0.0] + [math.log10(255 / i) for i in range(1, 256)]
    lut[0] = lut[1]
    lut = np.asarray(lut, dtype=np.float64)
    val = lut[orig[..., 0]] + lut[orig[..., 1]] + lut[orig[..., 2]]     # left-to-right, as the reference adds them
    return np.round(val * 100).astype(np.uint16)                          # round-half-even, like python round()


def create_kde(values, count, bandwidth=1.0):
    c = 1 / math.sqrt(2 * math.pi)
    step = (float(np.max(values)) + 1) / count
    n = values.shape[0]
    kde = np.zeros(count, dtype=np.float32)
    for i in range(count):
        val = (i * step - values) * (1 / bandwidth)
        total = np.cumsum(np.exp(-(val * val / 2)) * c)[-1]                # sequential sum, as the numba loop
        kde[i] = total / (n * bandwidth)
    return kde, step


def calculate_default_size_threshold(cell_sizes, resolution="40x"):
    cell_sizes = np.asarray(cell_sizes, dtype=np.int64)
    if cell_sizes.shape[0] <= 1:
        return 0
    kde, step = create_kde(np.sqrt(cell_sizes), 500)
    idx = 1
    for i in range(1, kde.shape[0] - 1):
        if kde[i] < kde[i - 1] and kde[i] < kde[i + 1]:
            idx = i
            break
    thresh_sqrt = (idx - 1) * step
    lo, default, hi = {"20x": (3, 4, 6), "10x": (2, 2, 3)}.get(resolution, (4, 7, 10))
    if thresh_sqrt < lo:
        thresh_sqrt = lo
    elif thresh_sqrt > hi:
        thresh_sqrt = default
    return int(round(thresh_sqrt * thresh_sqrt))


def calculate_stain_range(stain):
    nz = stain[stain != 0]
    if nz.shape[0] > 0:
        return (round(np.percentile(nz, 0.1)), round(np.percentile(nz, 99.9)))
    return (0, 0)


def calculate_default_marker_threshold(marker):
    lo, hi = calculate_stain_range(marker)
    return round((hi - lo) * 0.9) + lo


def get_cells_info(seg, marker, resolution, noise_thresh, seg_thresh, large_noise_thresh, use_od=False):
    if marker is not None and use_od:
        marker = create_od_image(marker)
    elif marker is not None and marker.ndim == 3:
        marker = marker.max(axis=-1)
    mask = mark_background(create_posneg_mask(seg, seg_thresh))
    mask, cells, lab, kept = compute_cell_mapping(mask, marker, noise_thresh, large_noise_thresh, use_od)
    defaults = {"size_thresh": calculate_default_size_threshold(np.asarray([c[0] for c in cells], np.int64), resolution)}
    if marker is not None and not use_od:
        defaults["marker_thresh"] = calculate_default_marker_threshold(marker)
    return mask, cells, defaults, lab, kept


def create_cell_classification(mask, cells, lab, kept, size_thresh=0, marker_thresh=None, size_thresh_upper=None,
                               od_thresh_lower=None, od_thresh_upper=None):
    """-> (new mask, counts).  `lab`/`kept` locate each listed cell's pixels (the reference re-floods from x0,y0)."""
    H, W = mask.shape
    n = int(lab.max())
    cls = np.zeros(n + 1, dtype=np.uint8)                   # 0 = not classified, 1 = negative, 2 = positive
    num_pos = num_neg = 0
    for cell, ci in zip(cells, kept.tolist()):
        if cell[0] > size_thresh and (size_thresh_upper is None or cell[0] < size_thresh_upper):
            is_pos = bool(cell[1])
            if marker_thresh is not None and cell[2] > marker_thresh:
                is_pos = True
            if od_thresh_lower is not None and cell[2] < od_thresh_lower:
                is_pos = False
            elif od_thresh_upper is not None and cell[2] > od_thresh_upper:
                is_pos = False
            cls[ci + 1] = 2 if is_pos else 1
            num_pos += is_pos
            num_neg += not is_pos
    out = mask.copy()
    pc = cls[lab]                                           # class of the pixel's cell (0 for background / dropped)
    first = np.zeros((H, W), dtype=bool)
    for cell, ci in zip(cells, kept.tolist()):
        if cls[ci + 1]:
            first[cell[4], cell[3]] = True
    out[pc == 2] = LABEL_POSITIVE
    out[pc == 1] = LABEL_NEGATIVE
    out[first & (pc == 2)] = LABEL_BORDER_POS
    out[first & (pc == 1)] = LABEL_BORDER_NEG
    # BACKGROUND pixels next to a non-first pixel of a classified cell: lowest cell index among the 4 neighbours wins
    big = np.int32(n + 1)
    src = np.where((pc > 0) & ~first, lab, big).astype(np.int32)
    best = np.full((H, W), big, dtype=np.int32)
    for dy, dx in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        best = np.minimum(best, _shift(src, dy, dx, big))
    hit = (mask == LABEL_BACKGROUND) & (best < big)
    bcls = cls[np.where(hit, best, 0)]
    out[hit & (bcls == 2)] = LABEL_BORDER_POS
    out[hit & (bcls == 1)] = LABEL_BORDER_NEG
    return out, {"num_total": num_pos + num_neg, "num_pos": num_pos, "num_neg": num_neg}


def enlarge_cell_boundaries(mask):
    out = mask.copy()
    todo = mask == LABEL_BACKGROUND
    for dy in (-1, 0, 1):                                   # raster order of the neighbour that does the writing
        for dx in (-1, 0, 1):
            if dy == 0 and dx == 0:
                continue
            nb = _shift(mask, dy, dx, LABEL_BACKGROUND)
            isb = (nb == LABEL_BORDER_POS) | (nb == LABEL_BORDER_NEG)
            w = todo & isb
            out[w] = nb[w]
            todo &= ~w
    return out


def create_final_images(orig, mask):
    overlay = orig.copy()
    refined = np.zeros_like(orig)
    overlay[mask == LABEL_BORDER_POS] = (255, 0, 0)
    overlay[mask == LABEL_BORDER_NEG] = (0, 0, 255)
    refined[(mask == LABEL_BORDER_POS) | (mask == LABEL_BORDER_NEG), 1] = 255
    refined[mask == LABEL_POSITIVE, 0] = 255
    refined[mask == LABEL_NEGATIVE, 2] = 255
    return overlay, refined


def calculate_large_noise_thresh(large_noise_thresh, resolution):
    if large_noise_thresh != "default":
        return large_noise_thresh
    return {"10x": 1000, "20x": 4000}.get(resolution, 16000)


def compute_final_results(orig, seg, marker, resolution, size_thresh="default", marker_thresh=None,
                          size_thresh_upper=None, seg_thresh=DEFAULT_SEG_THRESH, noise_thresh=DEFAULT_NOISE_THRESH,
                          large_noise_thresh=None, od_thresh_lower=None, od_thresh_upper=None, stages=None):
    large_noise_thresh = calculate_large_noise_thresh(large_noise_thresh, resolution)
    use_od = od_thresh_lower is not None or od_thresh_upper is not None
    mask, cells, defaults, lab, kept = get_cells_info(seg, orig if use_od else marker, resolution, noise_thresh,
                                                      seg_thresh, large_noise_thresh, use_od)
    if size_thresh is None:
        size_thresh = 0
    elif isinstance(size_thresh, str) and size_thresh == "default":
        size_thresh = defaults["size_thresh"]
    if isinstance(marker_thresh, str) and marker_thresh == "default":
        marker_thresh = defaults["marker_thresh"]
    mask, counts = create_cell_classification(mask, cells, lab, kept, size_thresh, marker_thresh, size_thresh_upper,
                                              od_thresh_lower, od_thresh_upper)
    if stages is not None:
        stages.update(cells=cells, defaults=defaults, classified=mask.copy())
    mask = enlarge_cell_boundaries(enlarge_cell_boundaries(mask))
    overlay, refined = create_final_images(np.asarray(orig), mask)
    scoring = {
        "num_total": counts["num_total"], "num_pos": counts["num_pos"], "num_neg": counts["num_neg"],
        "percent_pos": round(counts["num_pos"] / counts["num_total"] * 100, 1) if counts["num_pos"] > 0 else 0,
        "seg_thresh": seg_thresh, "size_thresh": size_thresh, "size_thresh_upper": size_thresh_upper,
        "marker_thresh": marker_thresh if marker is not None else None,
    }
    if stages is not None:
        stages.update(mask=mask)
    return overlay, refined, scoring


def synth_case(H, W, seed, n_cells=None):
    """Seeded synthetic (orig, seg, marker) uint8 images with the features the reference's loops branch on: red and
    blue blobs, rings with enclosed UNKNOWN holes, blobs cut by the image border, abutting cells of both classes,
    isolated specks below the noise threshold, and green-vetoed pixels.  Each blob only touches its bounding box, so
    region-sized cases stay cheap to build."""
    rng = np.random.default_rng(seed)
    seg = rng.integers(0, 40, size=(H, W, 3), dtype=np.int64)
    marker = rng.integers(0, 30, size=(H, W, 3), dtype=np.int64)
    n_cells = n_cells or max(6, H * W // 900)
    for _ in range(n_cells):
        cy, cx = rng.uniform(-4, H + 4), rng.uniform(-4, W + 4)
        a, b = rng.uniform(2.5, 11), rng.uniform(2.5, 11)
        th = rng.uniform(0, math.pi)
        ring = rng.random() < 0.25
        inner = rng.uniform(0.2, 0.5)
        pos = rng.random() < 0.5
        mval = int(rng.integers(40, 256))
        r = int(math.ceil(max(a, b))) + 1
        y0, y1, x0, x1 = max(0, int(cy) - r), min(H, int(cy) + r + 1), max(0, int(cx) - r), min(W, int(cx) + r + 1)
        if y0 >= y1 or x0 >= x1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        u = (xx - cx) * math.cos(th) + (yy - cy) * math.sin(th)
        v = -(xx - cx) * math.sin(th) + (yy - cy) * math.cos(th)
        d = (u / a) ** 2 + (v / b) ** 2
        inside = d <= 1
        if ring:
            inside &= d >= inner                                       # ring: encloses an UNKNOWN hole
        shp = inside.shape
        hi = rng.integers(110, 256, size=shp)
        lo = rng.integers(0, 90, size=shp)
        w = seg[y0:y1, x0:x1]
        w[..., 0] = np.where(inside, hi if pos else lo, w[..., 0])
        w[..., 2] = np.where(inside, lo if pos else hi, w[..., 2])
        w[..., 1] = np.where(inside, rng.integers(0, 95, size=shp), w[..., 1])        # some G > 80 vetoes
        if pos:
            mw = marker[y0:y1, x0:x1]
            mw[...] = np.where(inside[..., None], rng.integers(0, mval + 1, size=shp + (3,)), mw)
    speck = rng.random((H, W)) < 0.004
    seg[..., 0] = np.where(speck, 200, seg[..., 0])
    seg[..., 1] = np.where(speck, 0, seg[..., 1])
    orig = rng.integers(0, 256, size=(H, W, 3), dtype=np.int64)
    return orig.astype(np.uint8), seg.astype(np.uint8), marker.astype(np.uint8)

"""Tiling / stitching and the BatchNorm statistic toggles (reference deepliif/util/__init__.py:129-331, 743-771).

``TileGrid`` restates ``InferenceTiler``'s geometry on arrays so that all tiles of an image are produced as one
uint8 batch (for TilePipeline) and stitched back with the same centre-crop + border-strip rule."""
import os

import numpy as np
import torch


# names a result image ends with (`<stem>_<name>.png`): results written next to the inputs are not inputs of the next run
excluding_names = ["Hema", "DAPI", "DAPILap2", "Ki67", "Seg", "Marked", "SegRefined", "SegOverlaid", "Marker", "Lap2"]
image_extensions = [".png", ".jpg", ".tif", ".jpeg"]


def allowed_file(filename):
    """util/__init__.py:38-48: an image extension, and not one of this tool's own result files."""
    name, extension = os.path.splitext(filename)
    return extension in image_extensions and name.split("_")[-1] not in excluding_names


def chunker(iterable, size):
    for i in range(size):
        yield iterable[i::size]


def disable_batchnorm_tracking_stats(model):
    """Reference: inference uses batch statistics (running stats nulled).  The sm_100a engines always normalise
    with statistics of the data at hand, so this only records the mode on the stock BatchNorm containers."""
    for m in model.modules():
        if type(m) == torch.nn.BatchNorm2d and m.track_running_stats:
            m.track_running_stats = False
            m.running_mean_backup, m.running_var_backup = m.running_mean, m.running_var
            m.running_mean = m.running_var = None
    return model


def enable_batchnorm_tracking_stats(model):
    for m in model.modules():
        if type(m) == torch.nn.BatchNorm2d and hasattr(m, "running_mean_backup"):
            m.track_running_stats = True
            m.running_mean, m.running_var = m.running_mean_backup, m.running_var_backup
    return model


def image_variance_gray(img_u8_hwc):
    """Variance of the ITU-R 601 luma (PIL 'L' conversion) of an RGB tile over the pixels that are neither saturated
    white (255) nor black (0); 0 when no such pixel exists (util/__init__.py:478-485)."""
    a = np.asarray(img_u8_hwc).astype(np.int64)
    gray = ((a[..., 0] * 19595 + a[..., 1] * 38470 + a[..., 2] * 7471 + 0x8000) >> 16).astype(np.uint8)
    val = gray[(gray != 255) & (gray != 0)]
    return float(np.var(val)) if val.shape[0] else 0.0


def is_empty(tile):
    """models/__init__.py:391-396: a tile (or every tile of a list) whose unsaturated gray variance is below 9."""
    if isinstance(tile, list):
        return all(image_variance_gray(t) < 9 for t in tile)
    return image_variance_gray(tile) < 9


def _pad_cols(mod, tile_size):
    w = mod.shape[1]
    return np.pad(mod, ((0, 0), (0, tile_size - w % tile_size), (0, 0))) if w % tile_size else mod


def infer_background_colors_for_img(img, input_no=1, modalities_no=4, seg_no=1, tile_size=32):
    """util/__init__.py:421-470.  `img` = one training row (input | modalities | seg), uint8 [h, num_img*h, 3].  Boxes of
    the seg image(s) that are empty (is_empty) mark background; the mean colour of the same boxes in every modality image
    is that modality's background colour.  None when the row has no empty box."""
    a = np.asarray(img)
    h = a.shape[0]
    num_img = int(a.shape[1] / h)
    if h % tile_size:                      # PIL's crop pads boxes that stick out of the image with black
        a = np.pad(a, ((0, tile_size - h % tile_size), (0, 0), (0, 0)))
    boxes_per_seg = []
    for i in range(num_img - seg_no, num_img):
        mod = _pad_cols(a[:, h * i:h * (i + 1)], tile_size)
        boxes = []
        for x in range(0, h, tile_size):
            for y in range(0, h, tile_size):
                if is_empty(mod[y:y + tile_size, x:x + tile_size]):
                    boxes.append((x, y))
        boxes_per_seg.append(boxes)
    if len(boxes_per_seg) > 1:
        final = set()                      # the reference intersects with an empty set here: no box survives
        for b in boxes_per_seg:
            final = final & set(b)
        final = list(final)
    else:
        final = boxes_per_seg[0]
    if len(final) == 0:
        return None
    colors = {}
    for i in range(input_no, modalities_no + input_no):
        mod = _pad_cols(a[:, h * i:h * (i + 1)], tile_size)
        tiles = np.stack([mod[y:y + tile_size, x:x + tile_size] for x, y in final], axis=0)
        img_avg = np.mean(tiles, axis=0)
        colors[i] = np.mean(img_avg, axis=(0, 1)).astype(np.uint8)
    return colors


def infer_background_colors(dir_data, sample_size=5, input_no=1, modalities_no=4, seg_no=1, tile_size=32, return_list=False):
    """util/__init__.py:380-418: average the per-image background colours of the first `sample_size` training rows that
    have empty boxes (files are taken in sorted order here; the reference takes os.listdir order)."""
    from PIL import Image
    fns = sorted(x for x in os.listdir(dir_data) if x.endswith(".png"))
    sample_size = min(sample_size, len(fns))
    acc, count = {}, 0
    while count < sample_size and len(fns) > 0:
        fn = fns.pop(0)
        img = np.asarray(Image.open(os.path.join(dir_data, fn)).convert("RGB"))
        c = infer_background_colors_for_img(img, input_no, modalities_no, seg_no, tile_size)
        if c is not None:
            count += 1
            for k, v in c.items():
                acc.setdefault(k, []).append(v)
    if count == 0:
        print("None of the images have empty tiles for estimating averge background color. Try with a proper tile size.")
        return None
    print(f"Calculating average color for empty tiles from {count} images..")
    out = {k: np.mean(v, axis=0).astype(np.uint8) for k, v in acc.items()}
    return [tuple(int(x) for x in e) for e in out.values()] if return_list else out


class TileGrid:
    """All tiles of one image as a batch, and their stitching.

    Geometry of InferenceTiler (pad_size = 0): tiles of `tile_size` step by `tile_size - 2*overlap`; the last
    tile of a row/column is clamped to the image edge; a result tile contributes its centre
    [overlap, tile-overlap) plus the border strips on the sides that touch the image edge; later tiles
    overwrite earlier ones (row-major order)."""

    def __init__(self, img_u8_hwc, tile_size, overlap_size=0):
        if tile_size <= 0:
            raise ValueError("tile_size must be positive and non-zero")
        if overlap_size < 0:
            raise ValueError("overlap_size must be positive or zero")
        a = np.asarray(img_u8_hwc)
        self.orig_h, self.orig_w = a.shape[:2]
        # images smaller than a tile are grown by mirroring (util/__init__.py:201-214)
        while a.shape[1] < tile_size:
            a = np.concatenate([a, a[:, ::-1]], axis=1)
        while a.shape[0] < tile_size:
            a = np.concatenate([a, a[::-1]], axis=0)
        if self.orig_w < tile_size:
            a = a[:, :tile_size]
        if self.orig_h < tile_size:
            a = a[:tile_size]
        self.img = np.ascontiguousarray(a)
        self.H, self.W = self.img.shape[:2]
        self.ts = tile_size
        self.ov_w = 0 if tile_size >= self.W else overlap_size
        self.ov_h = 0 if tile_size >= self.H else overlap_size
        cw, ch = tile_size - 2 * self.ov_w, tile_size - 2 * self.ov_h
        if cw <= 0 or ch <= 0:
            raise ValueError("overlap_size is too large for the tile size")
        self.origins = []
        for y in range(0, self.H, ch):
            for x in range(0, self.W, cw):
                self.origins.append((min(x, self.W - tile_size), min(y, self.H - tile_size)))

    def __len__(self):
        return len(self.origins)

    def tiles(self, indices=None, out=None):
        """uint8 [T, ts, ts, 3] (all tiles, or the tiles `indices` only — a rank's shard; `out`: optional destination,
        e.g. a pinned staging buffer, filled in place)."""
        ts = self.ts
        sel = self.origins if indices is None else [self.origins[i] for i in indices]
        if out is None:
            if not sel:
                return np.zeros((0, ts, ts, self.img.shape[2]), self.img.dtype)
            return np.stack([self.img[y:y + ts, x:x + ts] for x, y in sel])
        for k, (x, y) in enumerate(sel):
            out[k] = self.img[y:y + ts, x:x + ts]
        return out

    def stitch(self, result_tiles):
        """result_tiles: uint8 [T, ts, ts, C] -> uint8 [orig_h, orig_w, C]"""
        ts = self.ts
        out = np.zeros((self.H, self.W) + result_tiles.shape[3:], dtype=result_tiles.dtype)
        for (x, y), t in zip(self.origins, result_tiles):
            x0 = 0 if x == 0 else self.ov_w
            x1 = ts if x == self.W - ts else ts - self.ov_w
            y0 = 0 if y == 0 else self.ov_h
            y1 = ts if y == self.H - ts else ts - self.ov_h
            out[y + y0:y + y1, x + x0:x + x1] = t[y0:y1, x0:x1]
        return out[:self.orig_h, :self.orig_w]

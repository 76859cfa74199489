"""Tiling / stitching and the BatchNorm statistic toggles (reference deepliif/util/__init__.py:129-331, 743-771).

``TileGrid`` restates ``InferenceTiler``'s geometry on arrays so that all tiles of an image are produced as one
uint8 batch (for TilePipeline) and stitched back with the same centre-crop + border-strip rule."""
import numpy as np
import torch


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
    """Variance of the ITU-R 601 luma (PIL 'L' conversion) of an RGB tile (util/__init__.py:478-485)."""
    a = np.asarray(img_u8_hwc).astype(np.int64)
    gray = (a[..., 0] * 19595 + a[..., 1] * 38470 + a[..., 2] * 7471 + 0x8000) >> 16
    return float(np.var(gray.astype(np.uint8)))


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

    def tiles(self):
        """uint8 [T, ts, ts, 3]"""
        ts = self.ts
        return np.stack([self.img[y:y + ts, x:x + ts] for x, y in self.origins])

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

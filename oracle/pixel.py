"""Oracle: numpy restatement of the byte/integer ends of the path (TEST INFRASTRUCTURE).

  transform            /root/reference/deepliif/data/__init__.py:133-138
                       ToTensor (/255) then Normalize(0.5, 0.5): uint8 HWC -> fp32 NCHW in [-1,1]
  tensor2im            deepliif/util/util.py:117-135
                       ((x[0] CHW->HWC + 1) / 2.0 * 255.0).astype(uint8)  (float64 promotion does NOT
                       happen: image_numpy is float32 and the python scalars keep it float32;
                       astype truncates toward zero)
  create_posneg_mask   deepliif/postprocessing.py:163-190 with labels :87-95, thresh default 120 (:83)
  seg_aggregate        deepliif/models/__init__.py:338  stack([mul(seg_k, w_k)]).sum(0)

Bit-exact gate: integer/byte outputs of the CUDA path must equal these on identical fp32 inputs.
"""
import numpy as np

LABEL_UNKNOWN = 50
LABEL_POSITIVE = 200
LABEL_NEGATIVE = 150
DEFAULT_SEG_THRESH = 120


def transform(img_u8_hwc: np.ndarray) -> np.ndarray:
    """uint8 [H,W,3] -> float32 [1,3,H,W].  ToTensor: x.float().div(255); Normalize: (x-0.5)/0.5."""
    x = img_u8_hwc.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None]


def tensor2im(t_nchw: np.ndarray) -> np.ndarray:
    """float32 [N,3,H,W] -> uint8 [H,W,3] of batch element 0 (util.py:130-135)."""
    image = t_nchw[0].astype(np.float32)
    image = (np.transpose(image, (1, 2, 0)) + 1) / 2.0 * 255.0
    return image.astype(np.uint8)


def tensor2im_batch(t_nchw: np.ndarray) -> np.ndarray:
    """Same arithmetic for every batch element: [N,3,H,W] -> uint8 [N,H,W,3]."""
    return np.stack([tensor2im(t_nchw[i:i + 1]) for i in range(t_nchw.shape[0])])


def create_posneg_mask(seg_u8_hwc: np.ndarray, thresh: int = DEFAULT_SEG_THRESH) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [H,W].  The reference adds two numba uint8 values; numba promotes
    uint8+uint8 to int64, so R+B does not wrap (postprocessing.py:184)."""
    r = seg_u8_hwc[..., 0].astype(np.int64)
    g = seg_u8_hwc[..., 1].astype(np.int64)
    b = seg_u8_hwc[..., 2].astype(np.int64)
    mask = np.full(seg_u8_hwc.shape[:2], LABEL_UNKNOWN, dtype=np.uint8)
    hit = (r + b > thresh) & (g <= 80)
    mask[hit & (r >= b)] = LABEL_POSITIVE
    mask[hit & (r < b)] = LABEL_NEGATIVE
    return mask


def seg_aggregate(segs, weights) -> np.ndarray:
    """fp32: stack([seg_k * w_k]).sum(0) — sequential fp32 adds in list order (torch sum over a
    leading dim of size <= 5 reduces sequentially per element)."""
    acc = np.zeros_like(segs[0], dtype=np.float32)
    for s, w in zip(segs, weights):
        acc = acc + s.astype(np.float32) * np.float32(w)
    return acc

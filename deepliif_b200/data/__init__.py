"""``transform``: PIL RGB tile -> fp32 NCHW [1,3,H,W] in [-1,1] (reference deepliif/data/__init__.py:133-138:
resize to a multiple of 4, ToTensor (/255), Normalize(0.5, 0.5)).  The host path below is the reference's
arithmetic in numpy (bit-exact, tests/golden/pixel_ends.npz); batched tiles use the device kernel
dlb_u8_to_f32 through deepliif_b200.pipeline."""
import numpy as np
import torch
from PIL import Image


def _make_multiple(img, base=4, method=Image.BICUBIC):
    ow, oh = img.size
    w, h = int(round(ow / base) * base), int(round(oh / base) * base)
    return img if (w, h) == (ow, oh) else img.resize((w, h), method)


def transform_array(img_u8_hwc: np.ndarray) -> np.ndarray:
    x = img_u8_hwc.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))[None]


def transform(img):
    img = _make_multiple(img.convert("RGB") if img.mode != "RGB" else img)
    return torch.from_numpy(transform_array(np.asarray(img)))

"""Boundary glue kept API-compatible with the reference's deepliif/util/util.py: tensor2im / tensor_to_pil
(util.py:117-139) and the model-name id helpers (util.py:208-269)."""
import os

import numpy as np
import torch
from PIL import Image


def mkdirs(paths):
    for p in ([paths] if isinstance(paths, str) else paths):
        os.makedirs(p, exist_ok=True)


def mkdir(path):
    os.makedirs(path, exist_ok=True)


def tensor2im(input_image, imtype=np.uint8):
    """fp32 [N,C,H,W] in [-1,1] -> uint8 [H,W,3] of batch element 0: trunc((x+1)/2*255) (util.py:130-135).
    CUDA tensors are quantised on the device (dlb_f32_to_u8) and only the bytes cross PCIe."""
    if isinstance(input_image, np.ndarray):
        return input_image.astype(imtype)
    if not isinstance(input_image, torch.Tensor):
        return input_image
    t = input_image.data
    if t.is_cuda and t.shape[1] == 3 and imtype == np.uint8:
        from .. import ops
        return ops.f32_to_u8(t[0:1].float().contiguous())[0].cpu().numpy()
    a = t[0].cpu().float().numpy()
    if a.shape[0] == 1:
        a = np.tile(a, (3, 1, 1))
    a = (np.transpose(a, (1, 2, 0)) + 1) / 2.0 * 255.0
    return a.astype(imtype)


def tensor_to_pil(t):
    return Image.fromarray(tensor2im(t))


def _generator_names(dir_model):
    """Model-name suffixes (after the leading 'G') found in a model directory: '1', 'S0', '51', ..."""
    pth = [fn for fn in os.listdir(dir_model) if fn.endswith(".pth") and "net_G" in fn]
    if pth:
        return [fn[:-4].split("_")[2][1:] for fn in pth]
    pt = [fn for fn in os.listdir(dir_model) if fn.endswith(".pt") and fn.startswith("G")]
    if not pt:
        raise Exception("Cannot find any model file ending with .pt or .pth in directory", dir_model)
    return [fn[1:-3] for fn in pt]


def get_mod_id_seg(dir_model):
    """Seg generators carry a two-character id ('S0'.., legacy '51'..): the id's first character."""
    return max(_generator_names(dir_model), key=len)[0]


def get_input_id(dir_model):
    """'0' when the base seg generator is G<seg>0 (new naming), '1' for legacy G51..G55."""
    return "0" if "0" in [n[1:] for n in _generator_names(dir_model)] else "1"


def init_input_and_mod_id(opt, dir_model=None):
    """(mod_id_seg, input_id) for fresh training vs. loading existing nets (util.py:242-269)."""
    d = dir_model if dir_model is not None else os.path.join(getattr(opt, "checkpoints_dir", "."), getattr(opt, "name", ""))
    fresh = opt.is_train and not getattr(opt, "continue_train", False)
    if hasattr(opt, "mod_id_seg"):
        mod_id_seg = opt.mod_id_seg
    elif fresh:
        mod_id_seg = "S" if hasattr(opt, "modalities_names") else opt.modalities_no + 1
    else:
        mod_id_seg = get_mod_id_seg(d)
    input_id = None
    if opt.model in ("DeepLIIF", "DeepLIIFKD"):
        input_id = "0" if fresh else get_input_id(d)
    return mod_id_seg, input_id

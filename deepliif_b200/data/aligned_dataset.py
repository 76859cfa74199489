"""Training data path: the reference's AlignedDataset semantics
(deepliif/data/aligned_dataset.py:36-113, base_dataset.py:62-151) feeding the GPU the B200 way.

What is kept from the reference: one PNG row = `input_no + modalities_no + seg_no` tiles side by side; one set of
random parameters per sample (crop position, flip; same `random` call order as get_params, base_dataset.py:62-78)
applied to every tile; BICUBIC resize to load_size ('resize'), width scaling ('scale_width'), crop, 'none' = round to a
multiple of 4; horizontal flip.
What changes: samples stay **uint8** on the host (3 B/pixel, not 12), batches are collated straight into pinned
memory as one [B, k, H, W, 3] block, copied with one async H2D on a side stream while the previous step computes,
and ToTensor + Normalize(0.5, 0.5) run on the device (dlb_u8_to_f32, bit-exact with the reference's arithmetic).
"""
import os
import random

import numpy as np
import torch
from PIL import Image

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".tif", ".tiff")      # image_folder.py:14-19


def make_dataset(d, max_dataset_size=None):
    """image_folder.py:26-37: every image under `d` (recursive, sorted walk); the caller sorts the result."""
    assert os.path.isdir(d), "%s is not a valid directory" % d
    images = []
    for root, _, fnames in sorted(os.walk(d)):
        for fname in fnames:
            if fname.lower().endswith(IMG_EXTENSIONS):
                images.append(os.path.join(root, fname))
    return images[:max_dataset_size] if max_dataset_size else images


def get_params(preprocess, load_size, crop_size, size):
    """base_dataset.py:62-78 (same `random` calls in the same order)."""
    w, h = size
    new_h, new_w = h, w
    if preprocess == "resize_and_crop":
        new_h = new_w = load_size
    elif preprocess == "scale_width_and_crop":
        new_w = load_size
        new_h = load_size * h // w
    x = random.randint(0, max(0, new_w - crop_size))
    y = random.randint(0, max(0, new_h - crop_size))
    flip = random.random() > 0.5
    return {"crop_pos": (x, y), "flip": flip}


def apply_transform(img, preprocess, load_size, crop_size, no_flip, params, method=Image.BICUBIC):
    """The PIL part of get_transform (base_dataset.py:81-118) for an RGB tile -> uint8 [H,W,3] (ToTensor/Normalize are
    done on the device)."""
    preprocess = preprocess or ""
    if "resize" in preprocess:
        if img.size != (load_size, load_size):
            img = img.resize((load_size, load_size), method)
    elif "scale_width" in preprocess:
        ow, oh = img.size
        if not (ow == load_size and oh >= crop_size):
            img = img.resize((load_size, int(max(load_size * oh / ow, crop_size))), method)
    if "crop" in preprocess:
        ow, oh = img.size
        x1, y1 = params["crop_pos"]
        if ow > crop_size or oh > crop_size:
            img = img.crop((x1, y1, x1 + crop_size, y1 + crop_size))
    if preprocess == "none":
        ow, oh = img.size
        h, w = int(round(oh / 4) * 4), int(round(ow / 4) * 4)
        if (h, w) != (oh, ow):
            img = img.resize((w, h), method)
    if not no_flip and params["flip"]:
        img = img.transpose(Image.FLIP_LEFT_RIGHT)
    return np.asarray(img)


class AlignedDataset(torch.utils.data.Dataset):
    """uint8 version of the reference's AlignedDataset for the DeepLIIF model: item = [k, H, W, 3] uint8 + path."""

    def __init__(self, opt, phase="train"):
        self.dir_AB = os.path.join(opt.dataroot, phase)
        self.AB_paths = sorted(make_dataset(self.dir_AB, opt.max_dataset_size))
        assert opt.load_size >= opt.crop_size
        self.preprocess, self.no_flip = opt.preprocess, opt.no_flip
        self.load_size, self.crop_size = opt.load_size, opt.crop_size
        self.input_no = getattr(opt, "input_no", 1)
        self.num_img = opt.modalities_no + opt.seg_no + self.input_no
        if opt.model not in ("DeepLIIF", "DeepLIIFKD"):
            raise Exception(f"model class {opt.model} does not have corresponding implementation in deepliif_b200/data/aligned_dataset.py")

    def __len__(self):
        return len(self.AB_paths)

    def __getitem__(self, index):
        path = self.AB_paths[index]
        AB = Image.open(path).convert("RGB")
        w, h = AB.size
        w2 = int(w / self.num_img)
        params = get_params(self.preprocess, self.load_size, self.crop_size, (w2, h))
        tiles = [apply_transform(AB.crop((w2 * i, 0, w2 * (i + 1), h)), self.preprocess, self.load_size, self.crop_size,
                                 self.no_flip, params) for i in range(self.num_img)]
        return torch.from_numpy(np.stack(tiles)), path


def collate_u8(items):
    """[B] x ([k,H,W,3] uint8, path) -> ([B,k,H,W,3] uint8, paths).  Plain pageable memory: this runs inside forked
    DataLoader workers, which must not touch CUDA (cudaHostAlloc in a forked child fails); the batch is pinned in the
    main process by DataLoader(pin_memory=True) / DeviceBatches._stage."""
    shape = (len(items),) + tuple(items[0][0].shape)
    out = torch.empty(shape, dtype=torch.uint8)
    for i, (t, _) in enumerate(items):
        out[i].copy_(t)
    return out, [p for _, p in items]


class DeviceBatches:
    """Iterate a DataLoader of uint8 batches as model inputs on the GPU: the H2D copy and the uint8 -> fp32 transform
    of batch i+1 are enqueued on a side stream before batch i is handed out, so they overlap the training step.
    Yields {'A': fp32 [B,3,H,W] (or a list when input_no > 1), 'B': [fp32 [B,3,H,W]] * targets, 'A_paths': [...]}."""

    def __init__(self, loader, device, input_no=1):
        self.loader, self.device, self.input_no = loader, device, input_no
        self.stream = torch.cuda.Stream(device=device)

    def __len__(self):
        return len(self.loader)

    def _stage(self, batch):
        from .. import ops
        u8, paths = batch
        if not u8.is_pinned():                      # loaders built without pin_memory=True: pin here, in the main process
            u8 = u8.pin_memory()
        with torch.cuda.stream(self.stream):
            d = u8.to(self.device, non_blocking=True)                   # one copy: [B, k, H, W, 3]
            B, k, H, W, _ = d.shape
            planes = [ops.u8_to_f32(d[:, j].contiguous()) for j in range(k)]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return planes, paths, ev, d

    def __iter__(self):
        it = iter(self.loader)
        nxt = next(it, None)
        staged = self._stage(nxt) if nxt is not None else None
        while staged is not None:
            planes, paths, ev, keep = staged
            nxt = next(it, None)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for p in planes + [keep]:
                p.record_stream(torch.cuda.current_stream(self.device))
            staged = self._stage(nxt) if nxt is not None else None       # overlaps the consumer's step
            A = planes[0] if self.input_no == 1 else planes[:self.input_no]
            yield {"A": A, "B": planes[self.input_no:], "A_paths": paths}

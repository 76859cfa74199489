"""Training data path (deepliif_b200/data/aligned_dataset.py) against the reference's AlignedDataset outputs
(tests/golden/aligned_dataset.npz, written by oracle/gen_golden.py from deepliif/data/aligned_dataset.py:36-113)."""
import os
import random

import numpy as np
import torch
from PIL import Image

from oracle.gen_golden import DATASET_CASES, checksum, dataset_opt, dataset_rows

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "aligned_dataset.npz"))


def _root(tmp_path):
    (tmp_path / "train").mkdir()
    for i, row in enumerate(dataset_rows()):
        Image.fromarray(row).save(tmp_path / "train" / f"s{i}.png")
    return str(tmp_path)


def test_aligned_dataset_matches_reference_golden(tmp_path):
    from deepliif_b200.data.aligned_dataset import AlignedDataset
    root = _root(tmp_path)
    for ci, (pre, ls, cs, nf) in enumerate(DATASET_CASES):
        ds = AlignedDataset(dataset_opt(root, pre, ls, cs, nf))
        assert len(ds) == 3
        for i in range(3):
            random.seed(100 + i)
            u8, path = ds[i]
            u8 = u8.numpy()
            assert path.endswith(f"s{i}.png")
            assert list(u8.shape) == GOLD[f"c{ci}_i{i}_shape"].tolist()
            assert np.array_equal(checksum(u8), GOLD[f"c{ci}_i{i}_sum"])
            assert np.array_equal(u8[:, ::3, ::3], GOLD[f"c{ci}_i{i}_sub"])


def test_collate_gives_one_block_per_batch(tmp_path):
    from deepliif_b200.data.aligned_dataset import AlignedDataset, collate_u8
    ds = AlignedDataset(dataset_opt(_root(tmp_path), "resize_and_crop", 40, 40, True))
    dl = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, num_workers=0, collate_fn=collate_u8)
    batches = list(dl)
    assert [tuple(b[0].shape) for b in batches] == [(2, 6, 40, 40, 3), (1, 6, 40, 40, 3)]
    assert batches[0][0].dtype == torch.uint8 and len(batches[0][1]) == 2

"""The reference's own DeepLIIFModel.optimize_parameters() (tests/golden/train_step.npz, written by
oracle/gen_golden.py from deepliif/models/DeepLIIF_model.py:205-467) against the oracle restatement of its first-step
losses: the D / L1 losses of the first step depend on the forward pass only, so the functional oracle nets plus the
loss logic the GPU tests use must reproduce them on CPU."""
import json
import os

import numpy as np
import torch

from oracle import nets
from oracle.gen_golden import TRAIN_CASE, train_batch, train_state_dicts

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_step.npz"))


def reference_losses():
    return json.loads(bytes(GOLD["losses"]).decode())


def oracle_first_step_losses(sd, batch, seg_weights):
    """Restatement of backward_D / backward_G's forward-only terms (DeepLIIF_model.py:205-429)."""
    A, Bs = batch["A"], batch["B"]
    n = TRAIN_CASE["modalities_no"]
    cfg = dict(n_blocks=2, norm="batch", use_dropout=False, padding_type="zero", norm_mode="batch")
    with torch.no_grad():
        fk = [nets.resnet_forward(A, sd[f"G{i + 1}"], **cfg) for i in range(n)]
        us = lambda t, s: nets.unet_forward(t, s, num_downs=7, norm="batch", norm_mode="batch")
        parts = [us(A, sd["GS0"])] + [us(fk[i], sd[f"GS{i + 1}"]) for i in range(n)]
        seg = sum(p_ * w_ for p_, w_ in zip(parts, seg_weights))
        D = lambda t, s: nets.nlayer_d_forward(t, s, n_layers=4, norm="batch", norm_mode="batch")
        bce, mse, sl1 = torch.nn.BCEWithLogitsLoss(), torch.nn.MSELoss(), torch.nn.SmoothL1Loss()
        want = {}
        for i in range(n):
            pf = D(torch.cat((A, fk[i]), 1), sd[f"D{i + 1}"]); pr = D(torch.cat((A, Bs[i]), 1), sd[f"D{i + 1}"])
            want[f"D_fake_{i + 1}"] = bce(pf, torch.zeros_like(pf)).item()
            want[f"D_real_{i + 1}"] = bce(pr, torch.ones_like(pr)).item()
            want[f"G_L1_{i + 1}"] = (sl1(fk[i], Bs[i]) * 100).item()
        conds = [A] + [Bs[i] for i in range(n)]
        pf = sum(D(torch.cat((c, seg), 1), sd[f"DS{i}"]) * seg_weights[i] for i, c in enumerate(conds))
        pr = sum(D(torch.cat((c, Bs[n]), 1), sd[f"DS{i}"]) * seg_weights[i] for i, c in enumerate(conds))
        want["D_fake_S"] = mse(pf, torch.zeros_like(pf)).item()
        want["D_real_S"] = mse(pr, torch.ones_like(pr)).item()
        want["G_L1_S"] = (sl1(seg, Bs[n]) * 100).item()
    return want


def test_oracle_first_step_losses_match_the_reference_model():
    torch.set_num_threads(min(16, os.cpu_count()))
    n = TRAIN_CASE["modalities_no"]
    want = oracle_first_step_losses(train_state_dicts(), train_batch(), [1 / (n + 1)] * (n + 1))
    ref = reference_losses()
    assert set(ref) == {f"{a}_{b}" for a in ("G_GAN", "G_L1", "D_real", "D_fake") for b in ("1", "2", "S")}
    for k, v in want.items():
        assert abs(v - ref[k]) <= 2e-5 * max(1.0, abs(ref[k])), (k, v, ref[k])

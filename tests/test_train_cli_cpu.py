"""`deepliif train` option prologue (deepliif_b200.training.prepare_train_params + Options) against the reference's own
command (tests/golden/train_cli_options.npz, captured by oracle/gen_golden.py from cli.py:213-386 at its
print_options(opt, save=True) call): same CLI defaults and overrides on the same synthetic dataroot -> same attributes."""
import json
import os

import numpy as np
import pytest

from oracle.gen_golden import TRAIN_CLI_CASES, options_snapshot, train_cli_dataset

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_cli_options.npz"))


def cli_defaults():
    from deepliif_b200.cli import cli
    return {p.name: p.default for p in cli.commands["train"].params}


@pytest.mark.parametrize("ci", range(len(TRAIN_CLI_CASES)))
def test_train_prologue_matches_reference(ci, tmp_path):
    from deepliif_b200 import training
    ds, over = TRAIN_CLI_CASES[ci]
    root = str(tmp_path)
    train_cli_dataset(root, **ds)
    kw = cli_defaults()
    kw.update(dataroot=root, checkpoints_dir=root, name="exp", gpu_ids=())
    kw.update(over)
    opt = training.build_options(training.prepare_train_params(kw))
    got = json.loads(options_snapshot(opt, root))
    want = json.loads(bytes(GOLD[f"c{ci}"]).decode())
    got.pop("gpu_ids", None); want.pop("gpu_ids", None)          # here: the local rank's GPU; reference on CPU: []
    diff = {k: (got.get(k, "<missing>"), want.get(k, "<missing>")) for k in set(got) | set(want) if got.get(k, "<missing>") != want.get(k, "<missing>")}
    assert not diff, diff


def test_train_command_has_every_reference_option():
    from deepliif_b200.cli import cli
    mine = {p.name for p in cli.commands["train"].params}
    want = set(json.loads(bytes(GOLD["c0"]).decode())) | {"dataroot", "checkpoints_dir", "local_rank", "loss_weights_g", "loss_weights_d"}
    # keys the prologue derives rather than takes from the command line
    derived = {"seg_no", "input_no", "scale_size", "lambda_identity", "pool_size", "loss_G_weights", "loss_D_weights", "netG", "netD",
               "n_layers_D", "lambda_L1", "lambda_feat", "background_colors", "is_train"}
    assert (want - derived) <= mine | {"is_train"}, sorted((want - derived) - mine)

"""Two-GPU checks (skipped on a one-GPU box): data-parallel training over NCCL equals single-GPU training on the
concatenated batch (instance norm => per-sample statistics, so the mean of per-rank gradients IS the full-batch
gradient), and the tile-sharded bench path runs under torchrun."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, torch
sys.path.insert(0, os.environ["DLB_ROOT"])
import torch.distributed as dist
from deepliif_b200 import training
from deepliif_b200.cli import TRAIN_DEFAULTS
from deepliif_b200.models import create_model
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = dict(TRAIN_DEFAULTS, dataroot="/tmp", checkpoints_dir="/tmp/dlb_mg", name="t", gpu_ids=(local,), modalities_no=1, seg_gen=False,
         norm="instance", no_dropout=True, padding="zero", net_g="resnet_2blocks", net_d="basic", batch_size=1)
opt = training.build_options(p)
torch.manual_seed(0)                      # identical initial weights on every rank
model = create_model(opt)
training.make_optimizers(model)
model.train()
g = torch.Generator().manual_seed(5)
A = torch.rand((2, 3, 128, 128), generator=g) * 2 - 1
B = torch.rand((2, 3, 128, 128), generator=g) * 2 - 1
sl = slice(rank, rank + 1) if world > 1 else slice(0, 2)
for _ in range(2):
    model.set_input({"A": A[sl], "B": [B[sl]], "A_paths": []})
    model.optimize_parameters()
torch.cuda.synchronize()
if rank == 0:
    sd = {k: v.detach().cpu() for k, v in model.netG1.module.state_dict().items()}
    torch.save(sd, os.environ["DLB_OUT"])
if world > 1:
    dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_two_ranks_equal_one_rank_full_batch(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DLB_ROOT=ROOT)
    one, two = str(tmp_path / "one.pt"), str(tmp_path / "two.pt")
    r = subprocess.run([sys.executable, str(script)], env=dict(env, DLB_OUT=one), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29577", str(script)], env=dict(env, DLB_OUT=two), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    a, b = torch.load(one), torch.load(two)
    lr = 2e-4
    n_all = n_bad = 0
    for k in a:
        if a[k].dtype.is_floating_point and "running" not in k:
            d = (a[k] - b[k]).abs()
            n_all += d.numel(); n_bad += int((d > 0.5 * lr).sum())
    print(f"2-rank DDP vs 1-rank full batch after 2 steps: {100 * n_bad / n_all:.3f}% of {n_all} weights differ by > lr/2")
    # losses are means over the batch: with a batch split in two the full-batch loss gradient is the rank mean
    assert n_bad / n_all < 0.05


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_under_torchrun_two_gpus():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29578", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
                        "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    print("2-GPU bench:", line["value"], "tiles/s; e2e", line["e2e"]["value"])
    assert line["n_gpus"] == 2 and line["value"] > 0


SHARD_WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["DLB_ROOT"])
from click.testing import CliRunner
from deepliif_b200.cli import cli
r = CliRunner().invoke(cli, ["test", "--input-dir", os.environ["DLB_IN"], "--output-dir", os.environ["DLB_OUT"], "--tile-size", "512",
                             "--model-dir", os.environ["DLB_MODEL"]])
if r.exit_code != 0:
    print(r.output[-3000:]); sys.exit(1)
import torch.distributed as dist
if dist.is_initialized():
    dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cli_test_tile_sharded_over_two_gpus_equals_one_gpu(tmp_path):
    """`deepliif test` under torchrun shards the tiles of every image over the ranks; the stitched PNGs must be
    byte-identical to a single-GPU run (per-sample statistics => tiles are independent)."""
    import numpy as np
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_inference_api_gpu import _write_model_dir
    mdir, _ = _write_model_dir(tmp_path)
    inp = tmp_path / "in"; inp.mkdir()
    rng = np.random.default_rng(21)
    Image.fromarray((rng.random((995, 1250, 3)) * 255).astype(np.uint8)).save(inp / "roi.png")     # 9 tiles
    script = tmp_path / "w.py"; script.write_text(SHARD_WORKER)
    env = dict(os.environ, DLB_ROOT=ROOT, DLB_IN=str(inp), DLB_MODEL=mdir)
    o1, o2 = str(tmp_path / "o1"), str(tmp_path / "o2")
    r = subprocess.run([sys.executable, str(script)], env=dict(env, DLB_OUT=o1), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29579", str(script)], env=dict(env, DLB_OUT=o2), capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    names = sorted(f for f in os.listdir(o1) if f.endswith(".png"))
    assert names and names == sorted(f for f in os.listdir(o2) if f.endswith(".png"))
    for f in names:
        assert np.array_equal(np.asarray(Image.open(os.path.join(o1, f))), np.asarray(Image.open(os.path.join(o2, f)))), f

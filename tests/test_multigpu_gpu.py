"""NCCL paths.  Two-GPU checks (skipped on a one-GPU box): data-parallel training over NCCL equals single-GPU training on
the concatenated batch (instance norm => per-sample statistics, so the mean of per-rank gradients IS the full-batch
gradient), and the tile-sharded bench path runs under torchrun.  One-GPU self-tests (always run): the same code under
`torchrun --nproc-per-node 1`, i.e. a real NCCL process group of size 1 — FlatAdam.broadcast_from_rank0 /
all_reduce_grads, the tile shard + NCCL gather of infer_tiles and the CUDA-graph capture of a step that contains the
all-reduce all execute, and must change nothing."""
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
if world > 1 or os.environ.get("DLB_FORCE_DIST"):
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
if os.environ.get("DLB_GRAPH"):
    stepper = training.GraphedStep(model, warmup=1)
    for _ in range(3):
        stepper({"A": A[sl], "B": [B[sl]], "A_paths": []})
else:
    for _ in range(int(os.environ.get("DLB_STEPS", "2"))):
        model.set_input({"A": A[sl], "B": [B[sl]], "A_paths": []})
        model.optimize_parameters()
torch.cuda.synchronize()
if rank == 0:
    sd = {k: v.detach().cpu() for k, v in model.netG1.module.state_dict().items()}
    torch.save(sd, os.environ["DLB_OUT"])
if dist.is_initialized():
    dist.barrier(); dist.destroy_process_group()
'''


def _torchrun(n, port, script, env, timeout=900):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)], env=env,
                          capture_output=True, text=True, timeout=timeout)


def _same_state(a, b):
    return all(torch.equal(a[k], b[k]) for k in a)


def test_nccl_world1_training_equals_plain_process(tmp_path):
    """1-GPU NCCL self-test: a process group of size 1 sends the gradient bucket through ncclAllReduce and the weights
    through ncclBroadcast; two optimisation steps must give bit-identical weights to a run without any process group."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DLB_ROOT=ROOT)
    one, two = str(tmp_path / "plain.pt"), str(tmp_path / "nccl1.pt")
    r = subprocess.run([sys.executable, str(script)], env=dict(env, DLB_OUT=one), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = _torchrun(1, 29571, script, dict(env, DLB_OUT=two, DLB_FORCE_DIST="1", NCCL_DEBUG="VERSION"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert _same_state(torch.load(one), torch.load(two))


def test_nccl_world1_graph_captured_step_with_all_reduce_equals_eager(tmp_path):
    """training.GraphedStep with an initialised NCCL group: the all-reduce of both buckets is captured inside the CUDA graph
    of the step; one eager + two replayed steps must equal three eager steps bit for bit."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DLB_ROOT=ROOT, DLB_FORCE_DIST="1")
    eager, graph = str(tmp_path / "eager.pt"), str(tmp_path / "graph.pt")
    r = _torchrun(1, 29572, script, dict(env, DLB_OUT=eager, DLB_STEPS="3"))
    assert r.returncode == 0, r.stderr[-2000:]
    r = _torchrun(1, 29573, script, dict(env, DLB_OUT=graph, DLB_GRAPH="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert _same_state(torch.load(eager), torch.load(graph))


def test_nccl_world1_cli_test_gathers_tiles_over_nccl(tmp_path):
    """`deepliif test` under `torchrun --nproc-per-node 1`: the tiles go through sharding.shard / the NCCL gather of
    infer_tiles; the PNGs must equal a plain single-process run byte for byte."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_inference_api_gpu import _write_model_dir
    mdir, _ = _write_model_dir(tmp_path, net_g="resnet_9blocks", net_gs="unet_512")
    inp = tmp_path / "in"; inp.mkdir()
    rng = np.random.default_rng(22)
    Image.fromarray((rng.random((600, 700, 3)) * 255).astype(np.uint8)).save(inp / "roi.png")       # 4 tiles
    script = tmp_path / "w.py"; script.write_text(SHARD_WORKER)
    env = dict(os.environ, DLB_ROOT=ROOT, DLB_IN=str(inp), DLB_MODEL=mdir)
    o1, o2 = str(tmp_path / "o1"), str(tmp_path / "o2")
    r = subprocess.run([sys.executable, str(script)], env=dict(env, DLB_OUT=o1), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    r = _torchrun(1, 29574, script, dict(env, DLB_OUT=o2))
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    names = sorted(f for f in os.listdir(o1) if f.endswith(".png"))
    assert names and names == sorted(f for f in os.listdir(o2) if f.endswith(".png"))
    for f in names:
        assert np.array_equal(np.asarray(Image.open(os.path.join(o1, f))), np.asarray(Image.open(os.path.join(o2, f)))), f


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
                        "--warmup", "3", "--no-cpu-baseline", "--no-extras"], capture_output=True, text=True, timeout=900)
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

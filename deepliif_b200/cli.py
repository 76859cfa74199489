"""`deepliif` console entry point: the `test` / `train` / `trainlaunch` commands of the reference CLI
(/root/reference/cli.py:72-193, 573-758, 833-919) for the DeepLIIF model on the sm_100a path.

  deepliif test --input-dir D --output-dir O --tile-size 512 --model-dir M [--gpu-ids 0] [--seg-intermediate]
                [--seg-only] [--mod-only]         writes <stem>_<name>.png and <stem>.json per input image
  deepliif train --dataroot D --name N ...        one process per GPU (see trainlaunch for several GPUs)
  deepliif trainlaunch --use-torchrun "<torchrun args>" ...   spawns torchrun deepliif_b200/scripts_train.py

There is no CPU fallback: `--gpu-ids -1` is rejected with a clear error (the reference would fall back to CPU).
"""
import glob
import json
import os
import subprocess
import sys

import click
import torch
from PIL import Image

from .models import infer_modalities
from .options import Options, print_options
from .util import allowed_file


@click.group()
def cli():
    """Commonly used DeepLIIF batch operations (B200-native build)."""


@cli.command()
@click.option("--input-dir", default="./Sample_Large_Tissues/", help="reads images from here")
@click.option("--output-dir", help="saves results here.")
@click.option("--tile-size", type=click.IntRange(min=1), required=True, help="tile size")
@click.option("--model-dir", default="./model-server/DeepLIIF_Latest_Model/", help="load models from here.")
@click.option("--filename-pattern", default="*", help="run inference on files of which the name matches the pattern.")
@click.option("--gpu-ids", type=int, multiple=True, help="gpu-ids 0 (one process drives one GPU)")
@click.option("--eager-mode", is_flag=True, help="accepted for compatibility: models are always loaded from .pth files")
@click.option("--epoch", default="latest", help="which epoch to load")
@click.option("--seg-intermediate", is_flag=True, help="also save intermediate segmentation images")
@click.option("--seg-only", is_flag=True, help="save only the final segmentation image; overwrites --seg-intermediate")
@click.option("--mod-only", is_flag=True, help="save only the translated modality images")
@click.option("--color-dapi", is_flag=True)
@click.option("--color-marker", is_flag=True)
@click.option("--BtoA", "btoa", is_flag=True, help="CycleGAN models only (load generator B): accepted, unused by the DeepLIIF model")
def test(input_dir, output_dir, tile_size, model_dir, filename_pattern, gpu_ids, eager_mode, epoch, seg_intermediate,
         seg_only, mod_only, color_dapi, color_marker, btoa=False):
    """Test trained models"""
    output_dir = output_dir or input_dir
    os.makedirs(output_dir, exist_ok=True)
    if mod_only:
        seg_only = seg_intermediate = False
    elif seg_only:
        seg_intermediate = False
    if filename_pattern == "*":
        print("use all alowed files")
        image_files = sorted(fn for fn in os.listdir(input_dir) if allowed_file(fn))
    else:
        print("match files using filename pattern", filename_pattern)
        image_files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(input_dir, filename_pattern)))
    print(len(image_files), "image files")
    assert "train_opt.txt" in os.listdir(model_dir), f"file train_opt.txt is missing from model directory {model_dir}"
    if gpu_ids and gpu_ids[0] == -1:
        raise click.UsageError("deepliif_b200 has no CPU path: --gpu-ids -1 is not available")
    if not torch.cuda.is_available():
        raise click.UsageError("deepliif_b200 needs a CUDA (sm_100a) device")
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or "LOCAL_RANK" in os.environ:   # torchrun (also --nproc-per-node 1): one process per GPU, tiles are
        # sharded across ranks, rank 0 stitches and writes
        import torch.distributed as dist
        torch.cuda.set_device(local)
        gpu_ids = (local,)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(gpu_ids[0] if gpu_ids else 0)
    opt = Options(path_file=os.path.join(model_dir, "train_opt.txt"), mode="test")
    opt.epoch = epoch
    opt.gpu_ids = list(gpu_ids) if gpu_ids else [torch.cuda.current_device()]
    seg_weights = getattr(opt, "seg_weights", None)
    print_options(opt)
    with click.progressbar(image_files, label=f"Processing {len(image_files)} images", item_show_func=lambda fn: fn) as bar:
        for filename in bar:
            img = Image.open(os.path.join(input_dir, filename)).convert("RGB")
            images, scoring = infer_modalities(img, tile_size, model_dir, True, color_dapi, color_marker, opt,
                                               return_seg_intermediate=seg_intermediate, seg_only=seg_only,
                                               mod_only=mod_only, seg_weights=seg_weights)
            if rank != 0:
                continue
            stem = filename[: filename.rfind(".")]
            for name, im in images.items():
                im.save(os.path.join(output_dir, f"{stem}_{name}.png"))
            if scoring is not None:
                with open(os.path.join(output_dir, f"{stem}.json"), "w") as f:
                    json.dump(scoring, f, indent=2)


TRAIN_DEFAULTS = dict(
    dataroot=None, name="experiment_name", checkpoints_dir="./checkpoints", gpu_ids=(), model="DeepLIIF", seg_weights=None,
    loss_weights_g=None, loss_weights_d=None, input_nc=3, output_nc=3, ngf=64, ndf=64, net_d="n_layers", net_g="resnet_9blocks",
    net_gs="unet_512", n_layers_d=4, norm="batch", init_type="normal", init_gain=0.02, no_dropout=False, padding="zero",
    upsample="convtranspose", direction="AtoB", serial_batches=False, num_threads=4, batch_size=1, load_size=512,
    crop_size=512, max_dataset_size=None, preprocess="resize_and_crop", no_flip=False, display_winsize=512, epoch="latest",
    load_iter=0, verbose=False, lambda_l1=100.0, is_train=True, save_latest_freq=500, save_epoch_freq=100,
    save_by_iter=False, continue_train=False, epoch_count=1, phase="train", lr_policy="linear", n_epochs=100,
    n_epochs_decay=100, optimizer="adam", beta1=0.5, lr_g=0.0002, lr_d=0.0002, lr_decay_iters=50, gan_mode="vanilla",
    gan_mode_s="lsgan", pool_size=50, seed=None, modalities_no=4, seg_gen=True, print_freq=100, dataset_mode="aligned",
    input_no=1, scale_size=512, with_val=False, precision="bf16x3",
    # written to train_opt.txt with the reference's defaults so that the reference can read a directory trained here
    # (its BaseModel reads opt.remote_transfer_cmd, base_model.py:49); the visdom / html ones are inert in this package
    label_smoothing=0.0, display_freq=400, display_ncols=4, display_id=1, display_server="http://localhost", display_env="main",
    display_port=8097, update_html_freq=1000, no_html=False, remote=False, remote_transfer_cmd=None, model_dir_teacher="",
    net_ds="n_layers", local_rank=None, debug=False, debug_data_size=10, monitor_image=None)


# `deepliif train`: every option of the reference command (cli.py:72-193), same names and defaults.  Options that only
# drive the visdom / html dashboards are accepted and written to train_opt.txt but have no effect here.
_TRAIN_OPTIONS = [
    # name, kwargs
    ("--dataroot", dict(required=True, type=str, help="path to images (must have the subfolder train)")),
    ("--name", dict(default="experiment_name")),
    ("--gpu-ids", dict(type=int, multiple=True, help="one GPU per process; several GPUs: trainlaunch / torchrun")),
    ("--checkpoints-dir", dict(default="./checkpoints")),
    ("--modalities-no", dict(default=4, type=int)),
    ("--modalities-names", dict(default="", type=str, help="comma-separated names of the input and target modalities")),
    ("--model", dict(default="DeepLIIF", type=str)),
    ("--model-dir-teacher", dict(default="", type=str)),
    ("--seg-weights", dict(default="", type=str, help="comma-separated weights of the seg generators' outputs")),
    ("--loss-weights-g", dict(default="", type=str)),
    ("--loss-weights-d", dict(default="", type=str)),
    ("--input-nc", dict(default=3)), ("--output-nc", dict(default=3)), ("--ngf", dict(default=64)), ("--ndf", dict(default=64)),
    ("--net-d", dict(default="n_layers")), ("--net-g", dict(default="resnet_9blocks")), ("--n-layers-d", dict(default=4)),
    ("--norm", dict(default="batch")), ("--init-type", dict(default="normal")), ("--init-gain", dict(default=0.02)),
    ("--no-dropout", dict(is_flag=True)), ("--upsample", dict(default="convtranspose")),
    ("--label-smoothing", dict(type=float, default=0.0)), ("--direction", dict(default="AtoB")),
    ("--serial-batches", dict(is_flag=True)), ("--num-threads", dict(default=4)), ("--batch-size", dict(default=1)),
    ("--load-size", dict(default=512)), ("--crop-size", dict(default=512)), ("--max-dataset-size", dict(type=int, default=None)),
    ("--preprocess", dict(type=str, default=None, help="resize_and_crop | crop | scale_width | scale_width_and_crop | none")),
    ("--no-flip", dict(is_flag=True)), ("--display-winsize", dict(default=512)), ("--epoch", dict(default="latest")),
    ("--load-iter", dict(default=0)), ("--verbose", dict(is_flag=True)), ("--lambda-L1", dict(default=100.0)),
    ("--is-train", dict(is_flag=True, default=True)), ("--continue-train", dict(is_flag=True)), ("--epoch-count", dict(type=int, default=0)),
    ("--phase", dict(default="train")), ("--n-epochs", dict(type=int, default=100)), ("--n-epochs-decay", dict(type=int, default=100)),
    ("--optimizer", dict(type=str, default="adam")), ("--beta1", dict(default=0.5)), ("--lr-g", dict(default=0.0002)),
    ("--lr-d", dict(default=0.0002)), ("--lr-policy", dict(default="linear")), ("--lr-decay-iters", dict(type=int, default=50)),
    ("--seed", dict(type=int, default=None)), ("--display-freq", dict(default=400)), ("--display-ncols", dict(default=4)),
    ("--display-id", dict(default=1)), ("--display-server", dict(default="http://localhost")), ("--display-env", dict(default="main")),
    ("--display-port", dict(default=8097)), ("--update-html-freq", dict(default=1000)), ("--print-freq", dict(default=100)),
    ("--no-html", dict(is_flag=True)), ("--save-latest-freq", dict(default=500)), ("--save-epoch-freq", dict(default=100)),
    ("--save-by-iter", dict(is_flag=True)), ("--remote", dict(type=bool, default=False)),
    ("--remote-transfer-cmd", dict(type=str, default=None)), ("--dataset-mode", dict(type=str, default="aligned")),
    ("--padding", dict(type=str, default="zero")), ("--seg-gen", dict(type=bool, default=True)),
    ("--net-ds", dict(type=str, default="n_layers")), ("--net-gs", dict(type=str, default="unet_512")),
    ("--gan-mode", dict(type=str, default="vanilla")), ("--gan-mode-s", dict(type=str, default="lsgan")),
    ("--local-rank", dict(type=int, default=None)), ("--with-val", dict(is_flag=True)), ("--debug", dict(is_flag=True)),
    ("--debug-data-size", dict(default=10, type=int)), ("--monitor-image", dict(default=None)),
    # additions of this package
    ("--precision", dict(default="bf16x3", help="bf16x3 (fp32-parity split operands) | bf16 (single pass)")),
    ("--cuda-graph", dict(is_flag=True, help="capture the optimisation step in a CUDA graph and replay it per batch "
                                             "(single GPU; removes the per-launch host cost that dominates at batch 1)")),
]


def _with_options(specs):
    def deco(f):
        for name, kw in reversed(specs):
            f = click.option(name, **kw)(f)
        return f
    return deco


@cli.command()
@_with_options(_TRAIN_OPTIONS)
def train(**kw):
    """General-purpose training script for the DeepLIIF multi-task image-to-image translation model."""
    from . import training
    training.run_training(training.prepare_train_params(kw))


@cli.command(context_settings=dict(ignore_unknown_options=True, allow_extra_args=True))
@click.option("--use-torchrun", type=str, default=None, help='torchrun options, e.g. "--nproc_per_node 8"')
@click.pass_context
def trainlaunch(ctx, use_torchrun):
    """Launch `train` under torchrun: one process per GPU, gradients all-reduced over NCCL (reference cli.py:697-758)."""
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts_train.py")
    if use_torchrun:
        cmd = [sys.executable, "-m", "torch.distributed.run", *use_torchrun.split(), script, *ctx.args]
    else:
        cmd = [sys.executable, script, *ctx.args]
    print("launching:", " ".join(cmd))
    sys.exit(subprocess.run(cmd).returncode)


@cli.command()
@click.option("--model-dir", required=True, help="reads models from here")
@click.option("--output-dir", help="saves serialized models here")
@click.option("--device", default="cpu", type=str)
@click.option("--epoch", default="latest", type=str)
@click.option("--verbose", default=0, type=int)
def serialize(model_dir, output_dir, device, epoch, verbose):
    """Write the load-ready form of a model directory (the counterpart of the reference's TorchScript export,
    cli.py:760-830): per generator one `<name>.pt` holding the fp32 weights (reference state_dict keys) AND the repacked
    tensor-core operand planes ([tap][Cout][Cin] hi/lo 16-bit, what dlb_conv_tc_fwd consumes), plus train_opt.txt.
    `test` / init_nets read such a directory like any other (weights_only load: a model directory is data, not code)."""
    import shutil
    from .models.serialized import write_packed_dir
    if not torch.cuda.is_available():
        raise click.UsageError("deepliif serialize packs the weights with the sm_100a library: a CUDA device is needed")
    output_dir = output_dir or model_dir
    os.makedirs(output_dir, exist_ok=True)
    opt = Options(path_file=os.path.join(model_dir, "train_opt.txt"), mode="test")
    opt.epoch = epoch
    opt.gpu_ids = [torch.cuda.current_device()]
    files = write_packed_dir(model_dir, output_dir, opt, verbose=bool(verbose))
    if os.path.abspath(output_dir) != os.path.abspath(model_dir):
        shutil.copy(os.path.join(model_dir, "train_opt.txt"), os.path.join(output_dir, "train_opt.txt"))
    for f in files:
        print("wrote", f)


if __name__ == "__main__":
    cli()

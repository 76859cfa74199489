"""`deepliif serialize` output: load-ready generator files.

The reference's serialize command traces every network to TorchScript so that inference does not rebuild the Python
modules (cli.py:760-830).  The analogue here is the form the sm_100a kernels consume: per network one ``<name>.pt`` (a plain
``torch.save`` dict, loadable with ``weights_only=True``) with

  format      "deepliif_b200.packed/1"
  arch        {"kind": "resnet" | "unet", ...constructor configuration}
  precision   operand format of the packed planes ("bf16x3")
  state_dict  fp32 weights under the reference's keys (BatchNorm running statistics dropped, as in a traced eval net)
  packed      {conv key: {"hi": [taps, Cout, Cin] 16-bit, "lo": ..., "transposed": bool}}: the K-major B operands of
              dlb_conv_tc_fwd (dlb_pack_weights_tc), for consumers that bind the C ABI directly
"""
import os

import torch

FORMAT = "deepliif_b200.packed/1"


def _packed_planes(net, precision):
    from .. import ops
    from ..engine import Precision
    prec = Precision.parse(precision)
    out = {}
    for name, mod in net.named_modules():
        if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
            tr = isinstance(mod, torch.nn.ConvTranspose2d)
            w = mod.weight.detach().float().contiguous()
            cin, cout = (w.shape[0], w.shape[1]) if tr else (w.shape[1], w.shape[0])
            if cin % 64 or cout % 32:
                continue                       # stem / head / first convs are repacked by the engine (lane-packed forms)
            d = ops.conv_desc(1, 8, 8, [cin], cout, w.shape[2], w.shape[3], mod.stride[0], mod.padding[0], tr,
                              mod.output_padding[0] if tr else 0)
            hi, lo = ops.pack_weights_tc(d, w.cuda(), prec.fmt, prec.split)
            out[name] = {"hi": hi.cpu(), "lo": lo.cpu() if lo is not None else None, "transposed": tr}
    return out


def write_packed_dir(model_dir, output_dir, opt, precision="bf16x3", verbose=False):
    from . import init_nets
    nets = init_nets(model_dir, True, opt)
    files = []
    for name, net in nets.items():
        sd = {k: v.detach().cpu() for k, v in net.state_dict().items()
              if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
        kind = "resnet" if type(net).__name__ == "ResnetGenerator" else "unet"
        blob = {"format": FORMAT, "arch": {"kind": kind, **{k: v for k, v in getattr(net, "cfg", {}).items()}},
                "precision": precision, "state_dict": sd, "packed": _packed_planes(net, precision)}
        path = os.path.join(output_dir, f"{name}.pt")
        torch.save(blob, path)
        files.append(path)
        if verbose:
            print(name, kind, "%d tensors, %d packed convs" % (len(sd), len(blob["packed"])))
    return files


def read_pt(path):
    """state_dict of a serialized network file: this package's packed format, or a TorchScript archive written by the
    reference's serialize (only its weights are used: the traced graph is a cuDNN program)."""
    try:
        blob = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(blob, dict) and blob.get("format") == FORMAT:
            return dict(blob["state_dict"])
    except Exception:
        pass
    return {k: v for k, v in torch.jit.load(path, map_location="cpu").state_dict().items()}

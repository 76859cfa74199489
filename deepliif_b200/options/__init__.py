"""Options: the attribute bag built from CLI parameters or from ``train_opt.txt`` (reference:
deepliif/options/__init__.py:8-217).  The text format ('{:>25}: {:<30}' per key, values re-parsed with eval)
and the train-/test-mode defaults are kept so model directories interchange with the reference."""
import ast
import os
import re
from pathlib import Path

from ..util.util import init_input_and_mod_id, mkdirs


def _parse(v):
    """Value of one train_opt.txt line: a Python literal (int / float / tuple / list / bool / None, as save_options wrote
    it) or, failing that, the string itself.  Literal parsing only: a model directory is data, not code."""
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError, TypeError, MemoryError, RecursionError):
        return v


def read_model_params(file_addr):
    params = {}
    with open(file_addr) as f:
        for line in f:
            if ":" not in line:
                continue
            key, val = line.split(":", 1)
            defaults = [x for x in re.findall(r"\[.+?\]", val) if x.startswith("[default")]
            if len(defaults) > 1:
                raise Exception("train_opt.txt should not contain multiple possible default keys in one line:", defaults)
            if defaults:
                val = val.replace(defaults[0], "")
            params[key.strip()] = _parse(val.strip())
    return params


class Options:
    def __init__(self, d_params=None, path_file=None, mode="train"):
        assert (d_params is None) != (path_file is None), "provide exactly one of d_params / path_file"
        assert mode in ("train", "test"), 'mode should be one of ["train", "test"]'
        if path_file:
            d_params = read_model_params(path_file)
        for k, v in d_params.items():
            setattr(self, k, v if (k == "phase" or not isinstance(v, str)) else _parse(v))
        if not hasattr(self, "optimizer"):
            self.optimizer = "adam"
        if mode == "train":
            self.is_train = True
            if hasattr(self, "net_g") and not hasattr(self, "netG"):
                self.netG = self.net_g
            if hasattr(self, "net_d") and not hasattr(self, "netD"):
                self.netD = self.net_d
            # hard overrides of the reference (options/__init__.py:65-67)
            self.n_layers_D, self.lambda_L1, self.lambda_feat = 4, 100, 100
            return
        # ---- test mode: back-compat defaults (options/__init__.py:69-180) -------------------------------
        self.phase, self.is_train, self.continue_train = "test", False, False
        self.input_nc, self.output_nc, self.ngf = 3, 3, 64
        self.norm = getattr(self, "norm", "batch")
        self.use_dropout = False
        if not hasattr(self, "modalities_no") and hasattr(self, "targets_no"):
            self.modalities_no = self.targets_no - 1
            del self.targets_no
        if not hasattr(self, "input_no"):
            self.input_no = 1
        if self.model in ("DeepLIIF", "DeepLIIFKD"):
            self.mod_id_seg, self.input_id = init_input_and_mod_id(self, os.path.dirname(path_file))
            if getattr(self, "seg_gen", True) is False:
                self.mod_id_seg = None
            self.input_id = int(self.input_id)
            if self.modalities_no == 4 and not hasattr(self, "modalities_names"):
                self.modalities_names = ["IHC", "Hema", "DAPI", "Lap2", "Marker"]
                self.seg_weights = [0.5, 0, 0, 0, 0.5]
            if not getattr(self, "modalities_names", None):
                self.modalities_names = [f"input{i + 1}" for i in range(self.input_no)] + \
                                        [f"mod{i + 1}" for i in range(self.modalities_no)]
        else:
            self.modalities_names = [f"mod{i}" for i in range(self.modalities_no + 1)]
        if not hasattr(self, "background_colors"):
            self.background_colors = ([(201, 211, 208), (10, 10, 10), (0, 0, 0), (10, 10, 10)]
                                      if self.model in ("DeepLIIF", "DeepLIIFKD") else [(10, 10, 10)] * self.modalities_no)
        model_dir = Path(path_file).parent
        self.checkpoints_dir, self.name = str(model_dir.parent), str(model_dir.name)
        if isinstance(getattr(self, "gpu_ids", ()), int):
            self.gpu_ids = (self.gpu_ids,)
        if not hasattr(self, "seg_no"):
            if self.model == "DeepLIIF":
                self.seg_no, self.seg_gen = 1, True
            else:
                raise Exception(f"seg_gen cannot be automatically determined for {self.model}")
        if not hasattr(self, "scale_size"):
            self.scale_size = 512
        if not hasattr(self, "seg_weights"):
            self.seg_weights = [0.25, 0.15, 0.25, 0.1, 0.25]
        n = self.modalities_no
        self.loss_G_weights = getattr(self, "loss_G_weights", [1 / n] * n)
        self.loss_D_weights = getattr(self, "loss_D_weights", [1 / n] * n)
        if not hasattr(self, "upsample"):
            self.upsample = "convtranspose"

    def _get_kwargs(self):
        return {k: v for k, v in vars(self).items()}


def format_options(opt):
    lines = ["----------------- Options ---------------"]
    lines += ["{:>25}: {:<30}".format(str(k), str(v)) for k, v in sorted(vars(opt).items())]
    lines.append("----------------- End -------------------")
    return "\n".join(lines)


def save_options(opt):
    expr_dir = os.path.join(opt.checkpoints_dir, opt.name)
    mkdirs(expr_dir)
    with open(os.path.join(expr_dir, "{}_opt.txt".format(opt.phase)), "wt") as f:
        f.write(format_options(opt) + "\n")


def print_options(opt, save=False):
    print(format_options(opt))
    if save:
        save_options(opt)

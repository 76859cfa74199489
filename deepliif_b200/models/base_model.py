"""BaseModel: the model-object surface the reference's drivers use (deepliif/models/base_model.py:11-341):
setup / train / eval / test / save_networks / load_networks / get_current_losses / get_current_visuals /
update_learning_rate.  Checkpoint naming ({epoch}_net_{name}.pth holding net.module.state_dict()) is kept."""
import os
from abc import ABC, abstractmethod
from collections import OrderedDict

import torch

from . import networks
from ..util import disable_batchnorm_tracking_stats, enable_batchnorm_tracking_stats


class BaseModel(ABC):
    def __init__(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.is_train = opt.is_train
        self.device = torch.device("cuda:{}".format(self.gpu_ids[0])) if self.gpu_ids else torch.device("cpu")
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.visual_names, self.optimizers, self.image_paths = [], [], [], [], []
        self.metric = 0

    @abstractmethod
    def set_input(self, input):
        ...

    @abstractmethod
    def forward(self):
        ...

    @abstractmethod
    def optimize_parameters(self):
        ...

    def _net(self, name):
        return getattr(self, "net" + name)

    def setup(self, opt):
        self.opt = opt
        if self.is_train:
            self.schedulers = [networks.get_scheduler(o, opt) for o in self.optimizers]
        if not self.is_train or getattr(opt, "continue_train", False):
            suffix = "iter_%d" % opt.load_iter if getattr(opt, "load_iter", 0) > 0 else opt.epoch
            self.load_networks(suffix)
        self.print_networks(getattr(opt, "verbose", False))

    def train(self):
        for name in self.model_names:
            enable_batchnorm_tracking_stats(self._net(name).train())

    def eval(self):
        for name in self.model_names:
            disable_batchnorm_tracking_stats(self._net(name).eval())

    def test(self):
        with torch.no_grad():
            self.forward()
            self.compute_visuals()

    def compute_visuals(self):
        pass

    def get_image_paths(self):
        return self.image_paths

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step(self.metric) if self.opt.lr_policy == "plateau" else s.step()
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str) and hasattr(self, n))

    def get_current_losses(self):
        return OrderedDict((n, float(torch.as_tensor(getattr(self, "loss_" + n)).detach())) for n in self.loss_names
                           if hasattr(self, "loss_" + n))

    def _unwrap(self, net):
        return net.module if hasattr(net, "module") and isinstance(net.module, torch.nn.Module) and \
            not isinstance(net, networks._EngineBacked) else net

    def save_networks(self, epoch, save_from_one_process=False):
        """One file per network: <save_dir>/<epoch>_net_<name>.pth = state_dict of the unwrapped module, on CPU."""
        if save_from_one_process and int(os.environ.get("RANK", "0")) != 0:
            return
        os.makedirs(self.save_dir, exist_ok=True)
        for name in self.model_names:
            net = self._unwrap(self._net(name))
            sd = OrderedDict((k, v.detach().cpu()) for k, v in net.state_dict().items())
            torch.save(sd, os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name)))

    def load_networks(self, epoch):
        for name in self.model_names:
            path = os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name))
            net = self._unwrap(self._net(name))
            pt = os.path.join(self.save_dir, "%s.pt" % name)
            if not os.path.exists(path) and os.path.exists(pt):
                # a `deepliif serialize` directory (reference cli.py:770-811): TorchScript archives of the traced nets.
                # Only the weights are taken (the traced graph is a cuDNN program); keys equal the eager state_dict's.
                print("loading the weights of the serialized model %s" % pt)
                have = net.state_dict()
                from .serialized import read_pt
                sd = read_pt(pt)
                for k, v in have.items():      # traced eval nets carry no BatchNorm running statistics: keep the fresh ones
                    if k not in sd and k.endswith(("running_mean", "running_var", "num_batches_tracked")):
                        sd[k] = v
            else:
                print("loading the model from %s" % path)
                sd = torch.load(path, map_location="cpu", weights_only=True)     # a checkpoint is data, not code
            if hasattr(sd, "_metadata"):
                del sd._metadata
            # InstanceNorm checkpoints written by torch < 0.4 may carry running stats: drop them
            for k in list(sd.keys()):
                if k.endswith(("running_mean", "running_var", "num_batches_tracked")) and k not in net.state_dict():
                    sd.pop(k)
            net.load_state_dict(sd)

    def print_networks(self, verbose):
        print("---------- Networks initialized -------------")
        for name in self.model_names:
            n = sum(p.numel() for p in self._net(name).parameters())
            if verbose:
                print(self._net(name))
            print("[Network %s] Total number of parameters : %.3f M" % (name, n / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):
        for net in (nets if isinstance(nets, list) else [nets]):
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad

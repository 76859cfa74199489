"""Drop-in network factory: ``define_G`` / ``define_D`` with the reference's signatures and state_dict layout.

Mirrors /root/reference/deepliif/models/networks.py:142-238 (factories), :357-450 (ResnetGenerator),
:516-615 (UnetGenerator), :618-664 (NLayerDiscriminator), :84-139 (init_weights / init_net).

The returned modules are *parameter containers*: ``self.model`` is an ``nn.Sequential`` of stock torch
modules laid out exactly like the reference's, so ``state_dict()`` keys/shapes (``model.1.weight``,
``model.10.conv_block.5.weight``, ``model.model.1.model.3...``) and ``{epoch}_net_{name}.pth`` files
interchange with the reference.  ``forward`` never runs those torch modules: it hands the parameters to the
sm_100a engine (deepliif_b200/engine.py -> libdeepliif_b200.so).  There is no CPU or cuDNN fallback; a
non-CUDA input raises.
"""
import functools
import os

import torch
import torch.nn as nn
from torch.nn import init

from .. import engine as _engine
from .._lib import DeepliifB200Error


class Identity(nn.Module):
    def forward(self, x):
        return x


def get_norm_layer(norm_type="instance"):
    """batch -> affine BatchNorm2d (tracks running stats); instance -> affine-free InstanceNorm2d; none."""
    if norm_type == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "none":
        return lambda _c: Identity()
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def _norm_name(norm_layer):
    f = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    if f is nn.BatchNorm2d:
        return "batch"
    if f is nn.InstanceNorm2d:
        return "instance"
    return "none"


def init_weights(net, init_type="normal", init_gain=0.02):
    """Conv/Linear weights ~ init_type, biases 0; BatchNorm gamma ~ N(1, gain), beta 0 (networks.py:84-112)."""
    fill = {"normal": lambda w: init.normal_(w, 0.0, init_gain),
            "xavier": lambda w: init.xavier_normal_(w, gain=init_gain),
            "kaiming": lambda w: init.kaiming_normal_(w, a=0, mode="fan_in"),
            "orthogonal": lambda w: init.orthogonal_(w, gain=init_gain)}
    if init_type not in fill:
        raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
    for m in net.modules():
        name = type(m).__name__
        if getattr(m, "weight", None) is not None and ("Conv" in name or "Linear" in name):
            fill[init_type](m.weight.data)
            if getattr(m, "bias", None) is not None:
                init.constant_(m.bias.data, 0.0)
        elif "BatchNorm2d" in name:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    print("initialize network with %s" % init_type)


class _Holder(nn.Module):
    """Stands in for the reference's DataParallel / DDP wrapper so ``net.module`` keeps working
    (base_model.py:208-210).  One process drives one GPU; gradient exchange is done by the trainer."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        if len(gpu_ids) > 1 and os.getenv("LOCAL_RANK") is None and os.getenv("RANK") is None:
            raise NotImplementedError("deepliif_b200 runs one process per GPU: launch with trainlaunch/torchrun "
                                      "instead of passing several --gpu-ids to one process")
        net.to(gpu_ids[0])
        net = _Holder(net)
    init_weights(net, init_type, init_gain=init_gain)
    return net


class _EngineBacked(nn.Module):
    """Common forward: (re)build the engine from the current parameters, run it."""
    precision = "bf16x3"       # parity default; "bf16" = single pass
    backend = "tc"
    GLOBAL_VERSION = 0         # bumped by optimizers that update parameter storage behind autograd's back

    def __init__(self):
        super().__init__()
        self._engine = None
        self._engine_key = None

    def _param_key(self):
        dev = next(self.parameters()).device
        return (str(dev), self.precision, self.backend, self.training, _EngineBacked.GLOBAL_VERSION,
                tuple((id(p), p._version) for p in self.parameters()))

    def _build_engine(self, device):
        raise NotImplementedError

    def engine(self):
        key = self._param_key()
        if self._engine is None or key != self._engine_key:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise DeepliifB200Error("deepliif_b200 networks run on CUDA (sm_100a) only; move the module to a GPU")
            self._engine = self._build_engine(dev)
            self._engine_key = key
        return self._engine

    def forward(self, input):
        if not input.is_cuda:
            raise DeepliifB200Error("deepliif_b200 has no CPU path: input must be a CUDA tensor")
        return self.engine().forward(input.float())


class ResnetBlock(nn.Module):
    """x + conv_block(x); conv_block indices follow networks.py:479-506 (pad modules, optional dropout)."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if padding_type not in ("reflect", "replicate", "zero"):
            raise NotImplementedError("padding [%s] is not implemented" % padding_type)
        if padding_type == "replicate":
            raise NotImplementedError("padding [replicate] is outside the B200 hot-path scope (SURVEY.md §2 row 1)")
        seq = []
        for half in range(2):
            if padding_type == "reflect":
                seq.append(nn.ReflectionPad2d(1))
            seq += [nn.Conv2d(dim, dim, kernel_size=3, padding=1 if padding_type == "zero" else 0, bias=use_bias),
                    norm_layer(dim)]
            if half == 0:
                seq.append(nn.ReLU(True))
                if use_dropout:
                    seq.append(nn.Dropout(0.5))
        self.conv_block = nn.Sequential(*seq)


class ResnetGenerator(_EngineBacked):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False, n_blocks=6,
                 padding_type="zero", upsample="convtranspose", use_spectral_norm=False):
        assert n_blocks >= 0
        super().__init__()
        if upsample != "convtranspose":
            raise NotImplementedError(f"upsample layer type {upsample} is outside the B200 hot-path scope")
        if use_spectral_norm:
            raise NotImplementedError("spectral norm is outside the B200 hot-path scope")
        self.cfg = dict(n_blocks=n_blocks, norm=_norm_name(norm_layer), use_dropout=use_dropout,
                        padding_type=padding_type)
        bias = self.cfg["norm"] == "instance"
        Pad = nn.ReflectionPad2d if padding_type == "reflect" else nn.ZeroPad2d
        seq = [Pad(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=bias), norm_layer(ngf), nn.ReLU(True)]
        c = ngf
        for _ in range(2):
            seq += [nn.Conv2d(c, 2 * c, kernel_size=3, stride=2, padding=1, bias=bias), norm_layer(2 * c), nn.ReLU(True)]
            c *= 2
        seq += [ResnetBlock(c, padding_type, norm_layer, use_dropout, bias) for _ in range(n_blocks)]
        for _ in range(2):
            seq += [nn.ConvTranspose2d(c, c // 2, kernel_size=3, stride=2, padding=1, output_padding=1, bias=bias),
                    norm_layer(c // 2), nn.ReLU(True)]
            c //= 2
        seq += [Pad(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*seq)

    def _build_engine(self, device):
        return _engine.ResnetEngine(self.state_dict(), device=device, precision=self.precision, backend=self.backend,
                                    trunk_n_tile=getattr(self, "trunk_n_tile", 0), fused=getattr(self, "fused", None),
                                    fuse_residual=getattr(self, "fuse_residual", None),
                                    norm_mode="batch" if (self.training and self.cfg["norm"] == "batch") else "sample",
                                    **self.cfg)


# ------------------------------------------------------------------------------------------------------------
# UnetGenerator / NLayerDiscriminator containers (state_dict layout of networks.py:516-664)
# ------------------------------------------------------------------------------------------------------------
class UnetSkipConnectionBlock(nn.Module):
    """Parameter container for one U-Net level; module order inside ``self.model`` fixes the state_dict keys:
    outermost [conv, sub, relu, convT, tanh]; innermost [lrelu, conv, relu, convT, norm];
    middle [lrelu, conv, norm, sub, relu, convT, norm, (dropout)]."""

    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        self.outermost = outermost
        bias = _norm_name(norm_layer) == "instance"
        cin = outer_nc if input_nc is None else input_nc
        down = nn.Conv2d(cin, inner_nc, kernel_size=4, stride=2, padding=1, bias=bias)
        if outermost:
            seq = [down, submodule, nn.ReLU(True),
                   nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1), nn.Tanh()]
        elif innermost:
            seq = [nn.LeakyReLU(0.2, True), down, nn.ReLU(True),
                   nn.ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=bias),
                   norm_layer(outer_nc)]
        else:
            seq = [nn.LeakyReLU(0.2, True), down, norm_layer(inner_nc), submodule, nn.ReLU(True),
                   nn.ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=bias),
                   norm_layer(outer_nc)]
            if use_dropout:
                seq.append(nn.Dropout(0.5))
        self.model = nn.Sequential(*seq)


class UnetGenerator(_EngineBacked):
    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm2d, use_dropout=False):
        super().__init__()
        self.cfg = dict(num_downs=num_downs, norm=_norm_name(norm_layer))
        self.use_dropout = bool(use_dropout)        # nn.Dropout(0.5) on the inner ngf*8 blocks in training (networks.py:536)
        blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, norm_layer=norm_layer, innermost=True)
        for _ in range(num_downs - 5):
            blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, submodule=blk, norm_layer=norm_layer, use_dropout=use_dropout)
        for mult in (4, 2, 1):
            blk = UnetSkipConnectionBlock(ngf * mult, ngf * mult * 2, submodule=blk, norm_layer=norm_layer)
        self.model = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=blk, outermost=True,
                                             norm_layer=norm_layer)

    def _build_engine(self, device):
        return _engine.UnetEngine(self.state_dict(), device=device, precision=self.precision,
                                  norm_mode="batch" if (self.training and self.cfg["norm"] == "batch") else "sample",
                                  **self.cfg)


class NLayerDiscriminator(_EngineBacked):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_spectral_norm=False):
        super().__init__()
        if use_spectral_norm:
            raise NotImplementedError("spectral norm is outside the B200 hot-path scope")
        self.cfg = dict(n_layers=n_layers, norm=_norm_name(norm_layer))
        bias = self.cfg["norm"] == "instance"
        seq = [nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            seq += [nn.Conv2d(ndf * prev, ndf * mult, kernel_size=4, stride=2 if n < n_layers else 1, padding=1, bias=bias),
                    norm_layer(ndf * mult), nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, kernel_size=4, stride=1, padding=1)]
        self.model = nn.Sequential(*seq)

    def _build_engine(self, device):
        # the discriminator only exists in training: BatchNorm2d uses pooled batch statistics
        return _engine.NLayerDEngine(self.state_dict(), device=device, precision=self.precision,
                                     norm_mode="batch" if self.cfg["norm"] == "batch" else "sample", **self.cfg)


_UNET_DOWNS = {"unet_32": 5, "unet_64": 6, "unet_128": 7, "unet_256": 8, "unet_512": 9}


def _define_G_impl(input_nc, output_nc, ngf, netG, norm, use_dropout, padding_type, upsample):
    norm_layer = get_norm_layer(norm_type=norm)
    if netG.startswith("resnet_"):
        n_blocks = int(netG.split("_")[1].replace("blocks", ""))
        return ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                               n_blocks=n_blocks, padding_type=padding_type, upsample=upsample)
    if netG in _UNET_DOWNS:
        return UnetGenerator(input_nc, output_nc, _UNET_DOWNS[netG], ngf, norm_layer=norm_layer, use_dropout=use_dropout)
    if netG == "unet_512_attention":
        raise NotImplementedError("Generator model name [unet_512_attention] is outside the B200 hot-path scope")
    raise NotImplementedError("Generator model name [%s] is not recognized" % netG)


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=[], padding_type="reflect", upsample="convtranspose"):
    """Create a generator (reference signature, networks.py:142-144): resnet_{n}blocks | unet_{32..512}."""
    net = _define_G_impl(input_nc, output_nc, ngf, netG, norm, use_dropout, padding_type, upsample)
    return init_net(net, init_type, init_gain, gpu_ids)


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02, gpu_ids=[]):
    """Create a discriminator (networks.py:196): basic (70x70 PatchGAN) | n_layers."""
    norm_layer = get_norm_layer(norm_type=norm)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm_layer=norm_layer)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer=norm_layer)
    elif netD == "pixel":
        raise NotImplementedError("Discriminator model name [pixel] is outside the B200 hot-path scope")
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


# ------------------------------------------------------------------------------------------------------------
# losses / schedulers (tiny tensors: plain torch, as the reference; networks.py:46-81, 244-317)
# ------------------------------------------------------------------------------------------------------------
OPTIMIZER_MAPPING = {n.lower(): n for n in dir(torch.optim) if n[0].isupper()}


def get_optimizer(optimizer_name):
    name = optimizer_name if hasattr(torch.optim, optimizer_name) else OPTIMIZER_MAPPING.get(optimizer_name)
    if name is None:
        raise NotImplementedError("optimizer [%s] is not found" % optimizer_name)
    return getattr(torch.optim, name)


def get_scheduler(optimizer, opt):
    from torch.optim import lr_scheduler
    if opt.lr_policy == "linear":
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: 1.0 - max(0, epoch + opt.epoch_count - opt.n_epochs)
                                     / float(opt.n_epochs_decay + 1))
    if opt.lr_policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == "plateau":
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5)
    if opt.lr_policy == "cosine":
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.n_epochs, eta_min=0)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


class GANLoss(nn.Module):
    """vanilla -> BCEWithLogits, lsgan -> MSE against a constant label map (networks.py:244-317)."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, label_smoothing=0.0):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        self.label_smoothing = label_smoothing
        if gan_mode == "lsgan":
            self.loss = nn.MSELoss()
        elif gan_mode == "vanilla":
            self.loss = nn.BCEWithLogitsLoss()
        elif gan_mode == "wgangp":
            self.loss = None
        else:
            raise NotImplementedError("gan mode %s not implemented" % gan_mode)

    def get_target_tensor(self, prediction, target_is_real):
        # networks.py:278-292: real labels are scaled by (1 - s), fake labels by s (with the default labels 1 / 0 the
        # fake target therefore stays 0 — kept as the reference has it)
        if target_is_real:
            return self.real_label.expand_as(prediction) * (1 - self.label_smoothing)
        return self.fake_label.expand_as(prediction) * self.label_smoothing

    def __call__(self, prediction, target_is_real):
        if self.gan_mode in ("lsgan", "vanilla"):
            return self.loss(prediction, self.get_target_tensor(prediction, target_is_real))
        return -prediction.mean() if target_is_real else prediction.mean()

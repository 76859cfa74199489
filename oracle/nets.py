"""Oracle: plain-PyTorch fp32 CPU restatement of the reference networks (TEST INFRASTRUCTURE).

Functional form: every function takes a reference-format ``state_dict`` (same keys / shapes as
``torch.save(net.state_dict())`` of the reference modules) and evaluates the
network with ``torch.nn.functional`` calls only.  No class of the reference is restated; the
layer order is derived from

  ResnetGenerator.__init__      /root/reference/deepliif/models/networks.py:366-446
  ResnetBlock.build_conv_block  networks.py:467-508   (forward: x + conv_block(x), :510-513)
  UnetSkipConnectionBlock       networks.py:548-615   (cat([x, model(x)], 1), :611-615)
  NLayerDiscriminator           networks.py:621-664
  get_norm_layer                networks.py:25-44     (batch: affine BN; instance: affine-free IN)
  disable_batchnorm_tracking_stats  deepliif/util/__init__.py:743-755 (eval => batch statistics)

Normalisation semantics ("norm_mode"):
  'sample'  statistics per (n, c) over H*W, optional affine.  This is InstanceNorm2d, and it is
            also what BatchNorm2d computes in the reference's *inference* path, which always runs
            batch = 1 with running stats nulled (data/__init__.py:133-138, util/__init__.py:743).
  'batch'   statistics per c over N*H*W (BatchNorm2d in training mode with N > 1).
Biased variance, eps = 1e-5 (PyTorch defaults used at networks.py:35-37).

Parity pinning: see tests/golden/ + oracle/gen_golden.py (outputs of the real reference modules
run in the build container on the same seeded state_dicts).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

EPS = 1e-5


# --------------------------------------------------------------------------------------------
# normalisation
# --------------------------------------------------------------------------------------------
def _norm(x, sd, key, norm, norm_mode):
    """norm in {'batch','instance','none'}; key = module prefix holding weight/bias if affine."""
    if norm == "none":
        return x
    gamma = sd.get(key + ".weight") if norm == "batch" else None
    beta = sd.get(key + ".bias") if norm == "batch" else None
    if norm_mode == "sample" or norm == "instance":
        return F.instance_norm(x, None, None, gamma, beta, True, 0.0, EPS)
    # pooled batch statistics (training-mode BatchNorm2d)
    return F.batch_norm(x, None, None, gamma, beta, True, 0.0, EPS)


def _pad(x, p, padding_type):
    if p == 0:
        return x
    if padding_type == "reflect":
        return F.pad(x, (p, p, p, p), mode="reflect")
    if padding_type == "replicate":
        return F.pad(x, (p, p, p, p), mode="replicate")
    return F.pad(x, (p, p, p, p))  # zero


# --------------------------------------------------------------------------------------------
# ResnetGenerator (networks.py:357-450)
# --------------------------------------------------------------------------------------------
def resnet_block_conv_indices(padding_type: str, use_dropout: bool):
    """Indices of (conv1, norm1, conv2, norm2) inside ``conv_block`` (networks.py:479-506).

    reflect/replicate insert a Pad module before each conv; dropout inserts one module after the
    first ReLU."""
    pad = 0 if padding_type == "zero" else 1
    c1 = pad
    n1 = c1 + 1
    c2 = n1 + 2 + (1 if use_dropout else 0) + pad
    n2 = c2 + 1
    return c1, n1, c2, n2


def resnet_forward(x, sd, *, n_blocks=9, norm="batch", use_dropout=False, padding_type="zero",
                   norm_mode="sample", taps=None):
    """Eval-mode forward (dropout = identity).  ``taps``: optional dict filled with named
    intermediate activations (for layer-wise bisecting of the CUDA path)."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    # stem: Pad3, Conv7x7 (model.1), Norm (model.2), ReLU             networks.py:386-397
    h = F.conv2d(_pad(x, 3, padding_type), sd["model.1.weight"], sd.get("model.1.bias"))
    tap("stem_conv", h)
    h = F.relu(_norm(h, sd, "model.2", norm, norm_mode))
    tap("stem", h)
    # 2x downsampling: Conv3x3 s2 p1 (model.4, model.7)               networks.py:399-404
    idx = 4
    for i in range(2):
        h = F.conv2d(h, sd[f"model.{idx}.weight"], sd.get(f"model.{idx}.bias"), stride=2, padding=1)
        h = F.relu(_norm(h, sd, f"model.{idx + 1}", norm, norm_mode))
        tap(f"down{i}", h)
        idx += 3
    # ResNet blocks (model.10 ..)                                     networks.py:406-411, 467-513
    c1, n1, c2, n2 = resnet_block_conv_indices(padding_type, use_dropout)
    bp = 1 if padding_type != "zero" else 0   # explicit pad module => conv padding 0
    for b in range(n_blocks):
        pre = f"model.{idx}.conv_block"
        y = F.conv2d(_pad(h, bp, padding_type), sd[f"{pre}.{c1}.weight"], sd.get(f"{pre}.{c1}.bias"),
                     padding=1 - bp)
        y = F.relu(_norm(y, sd, f"{pre}.{n1}", norm, norm_mode))
        y = F.conv2d(_pad(y, bp, padding_type), sd[f"{pre}.{c2}.weight"], sd.get(f"{pre}.{c2}.bias"),
                     padding=1 - bp)
        y = _norm(y, sd, f"{pre}.{n2}", norm, norm_mode)
        h = h + y
        tap(f"block{b}", h)
        idx += 1
    # 2x upsampling: ConvT3x3 s2 p1 op1                               networks.py:413-436
    for i in range(2):
        h = F.conv_transpose2d(h, sd[f"model.{idx}.weight"], sd.get(f"model.{idx}.bias"),
                               stride=2, padding=1, output_padding=1)
        h = F.relu(_norm(h, sd, f"model.{idx + 1}", norm, norm_mode))
        tap(f"up{i}", h)
        idx += 3
    # head: Pad3, Conv7x7 (+bias always), Tanh                        networks.py:438-444
    idx += 1
    h = F.conv2d(_pad(h, 3, padding_type), sd[f"model.{idx}.weight"], sd[f"model.{idx}.bias"])
    return torch.tanh(h)


def resnet_param_shapes(input_nc=3, output_nc=3, ngf=64, n_blocks=9, norm="batch",
                        use_dropout=False, padding_type="zero"):
    """Ordered {key: shape} of the reference ResnetGenerator state_dict."""
    bias = norm == "instance"      # use_bias = norm is InstanceNorm   networks.py:381-384
    out = OrderedDict()

    def conv(k, co, ci, r, b):
        out[k + ".weight"] = (co, ci, r, r)
        if b:
            out[k + ".bias"] = (co,)

    def nrm(k, c):
        if norm == "batch":
            out[k + ".weight"] = (c,)
            out[k + ".bias"] = (c,)
            out[k + ".running_mean"] = (c,)
            out[k + ".running_var"] = (c,)
            out[k + ".num_batches_tracked"] = ()

    conv("model.1", ngf, input_nc, 7, bias); nrm("model.2", ngf)
    idx, c = 4, ngf
    for _ in range(2):
        conv(f"model.{idx}", 2 * c, c, 3, bias); nrm(f"model.{idx + 1}", 2 * c)
        idx += 3; c *= 2
    c1, n1, c2, n2 = resnet_block_conv_indices(padding_type, use_dropout)
    for _ in range(n_blocks):
        pre = f"model.{idx}.conv_block"
        conv(f"{pre}.{c1}", c, c, 3, bias); nrm(f"{pre}.{n1}", c)
        conv(f"{pre}.{c2}", c, c, 3, bias); nrm(f"{pre}.{n2}", c)
        idx += 1
    for _ in range(2):
        out[f"model.{idx}.weight"] = (c, c // 2, 3, 3)     # ConvTranspose2d: (Cin, Cout, R, S)
        if bias:
            out[f"model.{idx}.bias"] = (c // 2,)
        nrm(f"model.{idx + 1}", c // 2)
        idx += 3; c //= 2
    idx += 1
    conv(f"model.{idx}", output_nc, ngf, 7, True)
    return out


# --------------------------------------------------------------------------------------------
# UnetGenerator (networks.py:516-615)
# --------------------------------------------------------------------------------------------
def unet_levels(num_downs, ngf=64, input_nc=3, output_nc=3):
    """Per level (outermost first): (outer_nc, inner_nc, in_ch).  networks.py:533-542."""
    chans = [(output_nc, ngf, input_nc), (ngf, ngf * 2, ngf), (ngf * 2, ngf * 4, ngf * 2),
             (ngf * 4, ngf * 8, ngf * 4)]
    for _ in range(num_downs - 5):
        chans.append((ngf * 8, ngf * 8, ngf * 8))
    chans.append((ngf * 8, ngf * 8, ngf * 8))  # innermost
    return chans


def _unet_prefixes(num_downs):
    """state_dict prefixes: level 0 = 'model.model', then '.1.model' (outermost child at index 1)
    or '.3.model' (child index 3 inside middle blocks: [lrelu, conv, norm, sub, ...])."""
    pre = ["model.model"]
    for lvl in range(1, num_downs):
        child = 1 if lvl == 1 else 3
        pre.append(f"{pre[-1]}.{child}.model")
    return pre


def unet_forward(x, sd, *, num_downs=9, norm="batch", norm_mode="sample", taps=None):
    """Eval-mode UNet forward.  Level layouts (networks.py:583-609):
       outermost: [conv(0), sub(1), relu(2), convT(3), tanh(4)]
       middle   : [lrelu(0), conv(1), norm(2), sub(3), relu(4), convT(5), norm(6), (dropout)]
       innermost: [lrelu(0), conv(1), relu(2), convT(3), norm(4)]"""
    pre = _unet_prefixes(num_downs)

    def run(lvl, h):
        p = pre[lvl]
        if lvl == 0:
            d = F.conv2d(h, sd[f"{p}.0.weight"], sd.get(f"{p}.0.bias"), stride=2, padding=1)
            s = run(1, d)
            u = F.conv_transpose2d(F.relu(s), sd[f"{p}.3.weight"], sd[f"{p}.3.bias"], stride=2, padding=1)
            return torch.tanh(u)
        a = F.leaky_relu(h, 0.2)
        if lvl == num_downs - 1:
            d = F.conv2d(a, sd[f"{p}.1.weight"], sd.get(f"{p}.1.bias"), stride=2, padding=1)
            u = F.conv_transpose2d(F.relu(d), sd[f"{p}.3.weight"], sd.get(f"{p}.3.bias"), stride=2, padding=1)
            u = _norm(u, sd, f"{p}.4", norm, norm_mode)
        else:
            d = F.conv2d(a, sd[f"{p}.1.weight"], sd.get(f"{p}.1.bias"), stride=2, padding=1)
            d = _norm(d, sd, f"{p}.2", norm, norm_mode)
            s = run(lvl + 1, d)
            u = F.conv_transpose2d(F.relu(s), sd[f"{p}.5.weight"], sd.get(f"{p}.5.bias"), stride=2, padding=1)
            u = _norm(u, sd, f"{p}.6", norm, norm_mode)
        # In-place LeakyReLU mutates the skip tensor in the reference (networks.py:573); the
        # only consumer of the skip is the parent's ReLU and relu(lrelu(x)) == relu(x), so the
        # concatenation of the un-activated h is numerically identical downstream except for
        # what the *caller* sees in the first half: the reference returns cat([lrelu(h), u]).
        out = torch.cat([a, u], 1)
        if taps is not None:
            taps[f"level{lvl}"] = out
        return out

    return run(0, x)


def unet_param_shapes(num_downs=9, ngf=64, input_nc=3, output_nc=3, norm="batch"):
    bias = norm == "instance"
    out = OrderedDict()
    pre = _unet_prefixes(num_downs)
    lv = unet_levels(num_downs, ngf, input_nc, output_nc)

    def nrm(k, c):
        if norm == "batch":
            out[k + ".weight"] = (c,); out[k + ".bias"] = (c,)
            out[k + ".running_mean"] = (c,); out[k + ".running_var"] = (c,)
            out[k + ".num_batches_tracked"] = ()

    def down(lvl):
        outer, inner, cin = lv[lvl]
        p = pre[lvl]
        if lvl == 0:
            out[f"{p}.0.weight"] = (inner, cin, 4, 4)
            if bias: out[f"{p}.0.bias"] = (inner,)
        else:
            out[f"{p}.1.weight"] = (inner, cin, 4, 4)
            if bias: out[f"{p}.1.bias"] = (inner,)
            if lvl != num_downs - 1:
                nrm(f"{p}.2", inner)

    def up(lvl):
        outer, inner, cin = lv[lvl]
        p = pre[lvl]
        if lvl == 0:
            out[f"{p}.3.weight"] = (inner * 2, outer, 4, 4); out[f"{p}.3.bias"] = (outer,)
        elif lvl == num_downs - 1:
            out[f"{p}.3.weight"] = (inner, outer, 4, 4)
            if bias: out[f"{p}.3.bias"] = (outer,)
            nrm(f"{p}.4", outer)
        else:
            out[f"{p}.5.weight"] = (inner * 2, outer, 4, 4)
            if bias: out[f"{p}.5.bias"] = (outer,)
            nrm(f"{p}.6", outer)

    # state_dict order = module registration order: down(l), sub..., up(l)
    def rec(lvl):
        down(lvl)
        if lvl < num_downs - 1:
            rec(lvl + 1)
        up(lvl)
    rec(0)
    return out


# --------------------------------------------------------------------------------------------
# NLayerDiscriminator (networks.py:618-664)
# --------------------------------------------------------------------------------------------
def nlayer_d_layers(n_layers=3, ndf=64, input_nc=6, norm="batch"):
    """[(key_conv, key_norm|None, cin, cout, stride, has_bias, lrelu)]"""
    bias = norm == "instance"
    L = [("model.0", None, input_nc, ndf, 2, True, True)]
    idx, mult = 2, 1
    for n in range(1, n_layers):
        prev, mult = mult, min(2 ** n, 8)
        L.append((f"model.{idx}", f"model.{idx + 1}", ndf * prev, ndf * mult, 2, bias, True))
        idx += 3
    prev, mult = mult, min(2 ** n_layers, 8)
    L.append((f"model.{idx}", f"model.{idx + 1}", ndf * prev, ndf * mult, 1, bias, True))
    idx += 3
    L.append((f"model.{idx}", None, ndf * mult, 1, 1, True, False))
    return L


def nlayer_d_forward(x, sd, *, n_layers=3, norm="batch", norm_mode="batch", taps=None):
    ndf = sd["model.0.weight"].shape[0]
    h = x
    for kc, kn, ci, co, st, hb, act in nlayer_d_layers(n_layers, ndf, x.shape[1], norm):
        h = F.conv2d(h, sd[kc + ".weight"], sd.get(kc + ".bias"), stride=st, padding=1)
        if kn is not None:
            h = _norm(h, sd, kn, norm, norm_mode)
        if act:
            h = F.leaky_relu(h, 0.2)
        if taps is not None:
            taps[kc] = h
    return h


def nlayer_d_param_shapes(n_layers=3, ndf=64, input_nc=6, norm="batch"):
    out = OrderedDict()
    for kc, kn, ci, co, st, hb, act in nlayer_d_layers(n_layers, ndf, input_nc, norm):
        out[kc + ".weight"] = (co, ci, 4, 4)
        if hb:
            out[kc + ".bias"] = (co,)
        if kn is not None and norm == "batch":
            out[kn + ".weight"] = (co,); out[kn + ".bias"] = (co,)
            out[kn + ".running_mean"] = (co,); out[kn + ".running_var"] = (co,)
            out[kn + ".num_batches_tracked"] = ()
    return out


# --------------------------------------------------------------------------------------------
# deterministic weights (own procedure — NOT the reference's RNG consumption order)
# --------------------------------------------------------------------------------------------
def make_state_dict(shapes, seed, init="reference"):
    """Seeded state_dict for a {key: shape} table.

    init='reference' follows the *distribution* of init_weights (networks.py:84-112): conv
    weights N(0, 0.02), conv biases 0, BatchNorm gamma N(1, 0.02), beta 0.
    init='stress' additionally draws conv biases and beta from N(0, 0.1) and gamma from
    N(1, 0.1) so bias / affine handling cannot hide behind zeros.
    Each tensor has its own generator seeded from (seed, position) so tensors are independent
    of each other and of torch's global RNG; CPU generation is deterministic for a torch build.
    """
    sd = OrderedDict()
    for i, (k, shp) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        leaf = k.rsplit(".", 1)[1]
        if leaf == "num_batches_tracked":
            sd[k] = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":
            sd[k] = torch.zeros(shp)
        elif leaf == "running_var":
            sd[k] = torch.ones(shp)
        elif leaf == "weight" and len(shp) == 4:
            sd[k] = torch.randn(shp, generator=g) * 0.02
        elif leaf == "weight":          # BatchNorm gamma
            sd[k] = 1.0 + torch.randn(shp, generator=g) * (0.1 if init == "stress" else 0.02)
        else:                           # bias: conv bias or BatchNorm beta
            sd[k] = torch.randn(shp, generator=g) * 0.1 if init == "stress" else torch.zeros(shp)
    return sd


# --------------------------------------------------------------------------------------------
# DeepLIIFModel.forward / run_dask cascade  (DeepLIIF_model.py:175-203, models/__init__.py:293-340)
# --------------------------------------------------------------------------------------------
def deepliif_forward(x, gens, segs, seg_weights, run_g, run_s):
    """gens: list of 4 state_dicts (G1..G4); segs: list of 5 (GS0..GS4).
    fake_B_i = G_i(x); fake_B_S_0 = GS0(x); fake_B_S_i = GS_i(fake_B_i);
    seg = sum_i w_i * fake_B_S_i  (torch.stack([mul]).sum(0), models/__init__.py:338)."""
    mods = [run_g(x, sd) for sd in gens]
    seg_parts = [run_s(x, segs[0])] + [run_s(m, sd) for m, sd in zip(mods, segs[1:])]
    seg = torch.stack([torch.mul(s, w) for s, w in zip(seg_parts, seg_weights)]).sum(dim=0)
    return mods, seg_parts, seg


def conv_flops(shapes_and_hw):
    """Algorithmic FLOPs helper: sum of 2*Cout*Hout*Wout*Cin*R*S over (co,ci,r,s,ho,wo)."""
    return sum(2 * co * ci * r * s * ho * wo for co, ci, r, s, ho, wo in shapes_and_hw)


def resnet_gflop(hw=512, n_blocks=9, ngf=64, nc=3):
    """Algorithmic GFLOP of one ResnetGenerator forward at hw x hw (396.41 @512)."""
    items = [(ngf, nc, 7, 7, hw, hw), (2 * ngf, ngf, 3, 3, hw // 2, hw // 2),
             (4 * ngf, 2 * ngf, 3, 3, hw // 4, hw // 4)]
    items += [(4 * ngf, 4 * ngf, 3, 3, hw // 4, hw // 4)] * (2 * n_blocks)
    # ConvT: 2*Cin*Hin*Win*Cout*R*S
    items += [(2 * ngf, 4 * ngf, 3, 3, hw // 4, hw // 4), (ngf, 2 * ngf, 3, 3, hw // 2, hw // 2)]
    items += [(nc, ngf, 7, 7, hw, hw)]
    return conv_flops(items) / 1e9

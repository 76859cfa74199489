"""Network executors: sequence the C-ABI kernels for the reference's generators / discriminators.

Each engine is built from a reference-format ``state_dict`` (same keys/shapes as the reference modules,
SURVEY.md §8b) and owns the repacked weights; ``forward`` enqueues only library kernels (see ops.py) on the
current CUDA stream.  Layer order follows

  ResnetGenerator         /root/reference/deepliif/models/networks.py:386-446 (+ ResnetBlock :479-513)
  UnetGenerator           networks.py:533-545 (+ UnetSkipConnectionBlock :573-615)
  NLayerDiscriminator     networks.py:636-660

Normalisation semantics: ``norm_mode='sample'`` = statistics per (n, c) — InstanceNorm2d, and also what the
reference's BatchNorm2d computes on its inference path (batch 1, running stats nulled,
deepliif/util/__init__.py:743-755), so batched tiles reproduce the reference's per-tile results;
``norm_mode='batch'`` = statistics pooled over the batch (training-mode BatchNorm2d).
"""
from dataclasses import dataclass

import torch

from . import ops
from .ops import (ACT_LRELU02, ACT_NONE, ACT_RELU, ACT_TANH, FMT_BF16, FMT_FP16, PAD_REFLECT, PAD_ZERO)


@dataclass
class Precision:
    """Operand format of the tensor-core convs.  split=True: hi/lo 16-bit planes, 3 MMAs per K step
    (fp32-grade products; parity default).  split=False: single pass (bf16/fp16 operand rounding)."""
    fmt: int = FMT_BF16
    split: bool = True

    @staticmethod
    def parse(name):
        table = {"bf16x3": Precision(FMT_BF16, True), "fp16x3": Precision(FMT_FP16, True),
                 "bf16": Precision(FMT_BF16, False), "fp16": Precision(FMT_FP16, False)}
        if name not in table:
            raise ValueError(f"precision [{name}] is not recognized (bf16x3 | fp16x3 | bf16 | fp16)")
        return table[name]


@dataclass
class Act:
    """An activation in HBM: fp32 NHWC and/or split 16-bit NHWC planes (optionally with a border)."""
    f32: torch.Tensor = None
    hi: torch.Tensor = None
    lo: torch.Tensor = None
    pad: int = 0


def _tc_ok(cins, cout):
    return all(c % 64 == 0 for c in cins) and cout % 32 == 0


class ConvLayer:
    """One nn.Conv2d / nn.ConvTranspose2d with weights repacked for its kernel."""

    def __init__(self, weight, bias, *, transposed=False, stride=1, pad=0, output_padding=0, cins=None,
                 prec=Precision(), backend="tc", n_tile=0):
        self.transposed, self.stride, self.pad, self.output_padding = transposed, stride, pad, output_padding
        w = weight.detach().to(torch.float32).contiguous()
        if transposed:
            cin_total, self.cout, self.R, self.S = w.shape
        else:
            self.cout, cin_total, self.R, self.S = w.shape
        self.cins = list(cins) if cins is not None else [cin_total]
        assert sum(self.cins) == cin_total
        self.bias = bias.detach().to(torch.float32).contiguous() if bias is not None else None
        self.prec = prec
        self.n_tile = n_tile
        self.use_tc = backend == "tc" and _tc_ok(self.cins, self.cout)
        d = ops.conv_desc(1, 8, 8, self.cins, self.cout, self.R, self.S, stride, pad, transposed, output_padding)
        if self.use_tc:
            self.w_hi, self.w_lo = ops.pack_weights_tc(d, w, prec.fmt, prec.split)
        else:
            assert len(self.cins) == 1, "direct kernel takes one source"
            self.w_packed = ops.pack_weights_direct(d, w)

    def desc(self, N, H, W, pad=None, pad_mode=PAD_ZERO):
        return ops.conv_desc(N, H, W, self.cins, self.cout, self.R, self.S, self.stride,
                             self.pad if pad is None else pad, self.transposed, self.output_padding, pad_mode)

    def run_tc(self, acts, N, H, W, pad=None):
        """acts: list of Act (split planes).  H, W: extents of the (possibly border-padded) operand."""
        d = self.desc(N, H, W, pad)
        return ops.conv_tc(d, [a.hi for a in acts], [a.lo for a in acts], self.w_hi, self.w_lo, self.bias,
                           self.prec.fmt, self.prec.split, self.n_tile)

    def run_direct(self, x, N, H, W, *, pad_mode=PAD_ZERO, in_nchw=False, in_scale=None, in_shift=None,
                   in_act=ACT_NONE, out_act=ACT_NONE, out_nchw=False):
        d = self.desc(N, H, W, None, pad_mode)
        return ops.conv_direct(d, x, self.w_packed, self.bias, in_nchw, in_scale, in_shift, in_act, out_act, out_nchw)


class _NormParams:
    def __init__(self, sd, key, norm, device):
        self.gamma = self.beta = None
        if norm == "batch":
            self.gamma = sd[key + ".weight"].detach().to(device=device, dtype=torch.float32).contiguous()
            self.beta = sd[key + ".bias"].detach().to(device=device, dtype=torch.float32).contiguous()


class _EngineBase:
    def __init__(self, norm, norm_mode, prec, backend, device):
        if norm not in ("batch", "instance", "none"):
            raise NotImplementedError("normalization layer [%s] is not found" % norm)
        self.norm, self.norm_mode, self.prec, self.backend, self.device = norm, norm_mode, prec, backend, device

    def _stats(self, y, np_):
        """raw conv output -> (scale, shift) or (None, None) for norm='none'."""
        if self.norm == "none" or np_ is None:
            return None, None
        pooled = self.norm == "batch" and self.norm_mode == "batch"
        return ops.norm_stats(y, np_.gamma, np_.beta, pooled)

    def _apply(self, y, scale, shift, act, *, residual=None, want_f32=False, want_split=True, pad=0,
               pad_mode=PAD_ZERO):
        f32, hi, lo = ops.norm_apply(y, scale, shift, act, residual, want_f32, want_split, self.prec.fmt, pad,
                                     pad_mode, need_lo=self.prec.split)
        return Act(f32, hi, lo, pad)

    def _conv(self, layer, act, N, H, W):
        """Run a conv on an Act with whichever kernel the layer was packed for.  Zero padding only."""
        if layer.use_tc:
            return layer.run_tc([act], N, H, W)
        return layer.run_direct(act.f32, N, H, W)


class ResnetEngine(_EngineBase):
    """ResnetGenerator forward (eval semantics: dropout = identity)."""

    def __init__(self, sd, *, n_blocks=9, norm="batch", use_dropout=False, padding_type="zero", norm_mode="sample",
                 precision="bf16x3", backend="tc", device="cuda", n_tile=0):
        prec = Precision.parse(precision) if isinstance(precision, str) else precision
        super().__init__(norm, norm_mode, prec, backend, device)
        if padding_type not in ("zero", "reflect"):
            raise NotImplementedError("padding [%s] is not implemented" % padding_type)
        self.n_blocks, self.padding_type = n_blocks, padding_type
        self.pad_mode = PAD_REFLECT if padding_type == "reflect" else PAD_ZERO
        g = lambda k: sd[k].to(device) if k in sd else None
        mk = lambda k, **kw: ConvLayer(g(k + ".weight"), g(k + ".bias"), prec=prec, backend=backend, n_tile=n_tile, **kw)
        nrm = lambda k: _NormParams(sd, k, norm, device)
        # stem / head always run on the fp32 direct kernel (Cin = 3 / Cout = 3)
        self.stem = ConvLayer(g("model.1.weight"), g("model.1.bias"), pad=3, backend="direct")
        self.stem_norm = nrm("model.2")
        idx = 4
        self.down, self.down_norm = [], []
        for _ in range(2):
            self.down.append(mk(f"model.{idx}", stride=2, pad=1)); self.down_norm.append(nrm(f"model.{idx + 1}"))
            idx += 3
        padm = 0 if padding_type == "zero" else 1
        c1 = padm; n1 = c1 + 1; c2 = n1 + 2 + (1 if use_dropout else 0) + padm; n2 = c2 + 1
        self.blocks = []
        for _ in range(n_blocks):
            pre = f"model.{idx}.conv_block"
            self.blocks.append((mk(f"{pre}.{c1}", pad=1), nrm(f"{pre}.{n1}"), mk(f"{pre}.{c2}", pad=1), nrm(f"{pre}.{n2}")))
            idx += 1
        self.up, self.up_norm = [], []
        for _ in range(2):
            self.up.append(mk(f"model.{idx}", transposed=True, stride=2, pad=1, output_padding=1))
            self.up_norm.append(nrm(f"model.{idx + 1}"))
            idx += 3
        idx += 1
        self.head = ConvLayer(g(f"model.{idx}.weight"), g(f"model.{idx}.bias"), pad=3, backend="direct")

    @torch.no_grad()
    def forward(self, x, taps=None):
        """x: fp32 NCHW [N,3,H,W] CUDA -> fp32 NCHW [N,3,H,W]."""
        x = x.contiguous()
        N, _, H, W = x.shape
        tc = self.backend == "tc"
        refl = self.pad_mode == PAD_REFLECT
        want = dict(want_f32=not tc, want_split=tc)

        def tap(name, a):
            if taps is not None:
                taps[name] = a

        # stem: Pad3 + Conv7x7 (NCHW input read directly) -> norm -> ReLU
        y = self.stem.run_direct(x, N, H, W, pad_mode=self.pad_mode, in_nchw=True)
        tap("stem_conv", y)
        sc, sh = self._stats(y, self.stem_norm)
        a = self._apply(y, sc, sh, ACT_RELU, **want)
        h, w = H, W
        # two stride-2 down convs
        for i in range(2):
            y = self._conv(self.down[i], a, N, h, w)
            h, w = h // 2, w // 2
            sc, sh = self._stats(y, self.down_norm[i])
            last = i == 1
            # the trunk keeps an fp32 residual stream next to the operand planes
            a = self._apply(y, sc, sh, ACT_RELU, want_f32=(not tc) or last, want_split=tc,
                            pad=1 if (refl and last and tc and self.n_blocks > 0) else 0, pad_mode=self.pad_mode)
            tap(f"down{i}", a)
        # ResNet blocks: x + Norm(Conv(ReLU(Norm(Conv(x)))))
        for b, (cv1, nm1, cv2, nm2) in enumerate(self.blocks):
            last = b == self.n_blocks - 1
            if tc:
                p = 1 if refl else 0
                y = cv1.run_tc([a], N, h + 2 * p, w + 2 * p, pad=0 if refl else 1)
                sc, sh = self._stats(y, nm1)
                t = self._apply(y, sc, sh, ACT_RELU, pad=p, pad_mode=self.pad_mode)
                y = cv2.run_tc([t], N, h + 2 * p, w + 2 * p, pad=0 if refl else 1)
                sc, sh = self._stats(y, nm2)
                a = self._apply(y, sc, sh, ACT_NONE, residual=a.f32, want_f32=True,
                                pad=0 if last else p, pad_mode=self.pad_mode)
            else:
                y = cv1.run_direct(a.f32, N, h, w, pad_mode=self.pad_mode)
                sc, sh = self._stats(y, nm1)
                y = cv2.run_direct(y, N, h, w, pad_mode=self.pad_mode, in_scale=sc, in_shift=sh, in_act=ACT_RELU)
                sc, sh = self._stats(y, nm2)
                a = self._apply(y, sc, sh, ACT_NONE, residual=a.f32, want_f32=True, want_split=False)
            tap(f"block{b}", a)
        # two ConvTranspose upsamplings
        for i in range(2):
            y = self._conv(self.up[i], a, N, h, w)
            h, w = h * 2, w * 2
            sc, sh = self._stats(y, self.up_norm[i])
            if i == 0:
                a = self._apply(y, sc, sh, ACT_RELU, **want)
                tap("up0", a)
        # head: (norm + ReLU fused into the load) Pad3 + Conv7x7 + bias + Tanh, NCHW out
        out = self.head.run_direct(y, N, h, w, pad_mode=self.pad_mode, in_scale=sc, in_shift=sh, in_act=ACT_RELU,
                                   out_act=ACT_TANH, out_nchw=True)
        return out

    __call__ = forward

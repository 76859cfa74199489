"""Training-mode executors: forward that records a tape + backward that walks it (autograd of the reference's
generators / discriminators, i.e. what ``loss.backward()`` does in DeepLIIF_model.py:332 and :429, on the sm_100a
kernels).  Per recorded layer  y = conv(x);  a = act(norm(y)) [+ residual]:

    backward:  dn-path   ops.norm_bwd      (norm + activation backward, param grads dgamma/dbeta)
               dW        ConvLayer.wgrad   (tcgen05, K over pixels)   [+ bias grad = ops.channel_sum]
               dx        ConvLayer.dgrad   (conv_tc with the weight in the opposite role)

Scope: ResnetGenerator (zero / reflect padding, input gradient), UnetGenerator, NLayerDiscriminator; nn.Dropout(0.5) with the counter-based mask
of csrc/rng.cuh (regenerated in backward from (seed, element index)).  Gradients are returned keyed by the reference's state_dict
names so the nn.Module containers can store them in ``param.grad``."""
import torch

from . import engine as _engine
from . import ops
from .engine import (ACT_LRELU02, ACT_NONE, ACT_RELU, ACT_TANH, PAD_REFLECT, PAD_ZERO, Act, ConvLayer, NLayerDEngine,
                     Precision, ResnetEngine, _EngineBase, _NormParams)


class _Rec:
    """One conv -> norm -> activation step of the tape."""
    __slots__ = ("layer", "x", "dims", "pad", "y", "sc", "sh", "mean", "rstd", "act", "np", "wkey", "nkey", "extra")

    def __init__(self, **kw):
        self.extra = None
        for k, v in kw.items():
            setattr(self, k, v)


def _pad_lanes64(w, transposed=False):
    """Zero-pad the output-channel dim of a conv weight to 64 lanes (GEMM M/N tiles are multiples of 64)."""
    if transposed:
        ci, co, R, S = w.shape
        out = torch.zeros((ci, 64, R, S), dtype=torch.float32, device=w.device); out[:, :co] = w
    else:
        co, ci, R, S = w.shape
        out = torch.zeros((64, ci, R, S), dtype=torch.float32, device=w.device); out[:co] = w
    return out


class _TrainOps:
    """Shared tape helpers (mixed into the engines below)."""
    # training forwards of a tracking BatchNorm2d move its running statistics (momentum 0.1, networks.py:34-37):
    # _EngineBase._stats passes the module's buffers to the finalize kernel when this is set
    update_running_stats = True

    @staticmethod
    def _new_seed():
        """Dropout seed for one layer invocation, drawn from torch's CPU generator (reproducible under --seed)."""
        return int(torch.randint(0, 2 ** 62, (1,)).item())

    def _fwd(self, tape, layer, np_, srcs, N, H, W, pad, act, wkey, nkey, *, residual=None, want_f32=False,
             want_split=True, out_pad=0, out_pad_mode=PAD_ZERO, drop=None):
        y, ws = layer.run_tc(srcs, N, H, W, pad)
        sc, sh, mean, rstd = self._stats(y, np_, ws, want_stats=True)
        a = self._apply(y, sc, sh, act, residual=residual, want_f32=want_f32, want_split=want_split, pad=out_pad,
                        pad_mode=out_pad_mode, drop=drop)
        rec = _Rec(layer=layer, x=srcs[0], dims=(N, H, W), pad=pad, y=y, sc=sc, sh=sh, mean=mean, rstd=rstd, act=act,
                   np=np_, wkey=wkey, nkey=nkey)
        rec.extra = drop
        tape.append(rec)
        return a, y

    def _bwd(self, rec, grads, dout, dout2=None, need_dx=True):
        """dout: fp32 NHWC gradient wrt the activation output of `rec`.  Returns dx (fp32 NHWC) or None."""
        layer, (N, H, W) = rec.layer, rec.dims
        dg = db = None
        if rec.np is not None and rec.np.gamma is not None:
            dg = torch.empty_like(rec.np.gamma); db = torch.empty_like(rec.np.gamma)
        # A bias in front of a norm layer has an analytically zero gradient (the norm removes per-channel constants;
        # autograd only produces rounding noise there): it is returned as exact zeros without touching dy again.
        normed = rec.sc is not None
        need_f32 = layer.bias is not None and not normed
        dp, dseed = rec.extra if rec.extra is not None else (0.0, 0)
        f32, hi, lo = ops.norm_bwd(dout, rec.y, rec.sc, rec.sh, rec.mean, rec.rstd, rec.act, dout2=dout2,
                                   pooled=self._pooled(), dgamma=dg, dbeta=db, want_f32=need_f32, want_split=True,
                                   fmt=self.prec.fmt, need_lo=self.prec.split, drop_p=dp, drop_seed=dseed,
                                   drop_epoch=_engine.DROP_EPOCH[0] if dp > 0 else None)
        if dg is not None:
            grads[rec.nkey + ".weight"], grads[rec.nkey + ".bias"] = dg, db
        grads[rec.wkey + ".weight"] = layer.wgrad(rec.x, hi, lo, N, H, W, rec.pad)
        if layer.bias is not None:
            grads[rec.wkey + ".bias"] = torch.zeros_like(layer.bias) if normed else ops.channel_sum(f32)
        return layer.dgrad(hi, lo, N, H, W, rec.pad) if need_dx else None


class ResnetTrainEngine(ResnetEngine, _TrainOps):
    def __init__(self, sd, **kw):
        self.use_dropout = bool(kw.get("use_dropout", False))
        kw.setdefault("backend", "tc")
        super().__init__(sd, **kw)
        if not (self.stem_tc and self.head_tc):
            raise NotImplementedError("training path needs the tensor-core stem/head (input_nc, output_nc <= 4)")
        dev = self.device
        # 64-lane copy of the head's virtual-channel weight for its wgrad / dgrad GEMMs
        wv32 = self.head.w_f32
        wv64 = torch.zeros((64,) + tuple(wv32.shape[1:]), dtype=torch.float32, device=wv32.device)
        wv64[:32] = wv32
        self.head64 = ConvLayer(wv64, None, pad=0, prec=self.prec, backend="tc")
        self.keys = self._layer_keys(sd)

    def _layer_keys(self, sd):
        idx = 4
        down = []
        for _ in range(2):
            down.append((f"model.{idx}", f"model.{idx + 1}")); idx += 3
        blocks = []
        # ReflectionPad2d modules (padding_type='reflect') and the Dropout module shift the indices (networks.py:479-506)
        padm = 1 if self.pad_mode == PAD_REFLECT else 0
        c1 = padm; n1 = c1 + 1
        c2 = n1 + 2 + (1 if self.use_dropout else 0) + padm
        n2 = c2 + 1
        for _ in range(self.n_blocks):
            pre = f"model.{idx}.conv_block"
            blocks.append(((f"{pre}.{c1}", f"{pre}.{n1}"), (f"{pre}.{c2}", f"{pre}.{n2}"))); idx += 1
        up = []
        for _ in range(2):
            up.append((f"model.{idx}", f"model.{idx + 1}")); idx += 3
        idx += 1
        return dict(stem=("model.1", "model.2"), down=down, blocks=blocks, up=up, head=f"model.{idx}")

    def forward_train(self, x):
        x = x.contiguous()
        N, _, H, W = x.shape
        tape = []
        K = self.keys
        pm = self.pad_mode
        p = 1 if pm == PAD_REFLECT else 0        # reflect: block operands are materialised with their border, convs run unpadded
        xh, xl = ops.stem_window_pack(x, 3, self.stem_S, pm, self.prec.fmt, self.prec.split)
        a, _ = self._fwd(tape, self.stem, self.stem_norm, [Act(None, xh, xl)], N, H + 6, W, 0, ACT_RELU, *K["stem"])
        h, w = H, W
        for i in range(2):
            last = i == 1
            a, _ = self._fwd(tape, self.down[i], self.down_norm[i], [a], N, h, w, None, ACT_RELU, *K["down"][i],
                             want_f32=last, out_pad=p if (last and self.n_blocks > 0) else 0, out_pad_mode=pm)
            h, w = h // 2, w // 2
        for b, (cv1, nm1, cv2, nm2) in enumerate(self.blocks):
            (k1, kn1), (k2, kn2) = K["blocks"][b]
            last = b == self.n_blocks - 1
            t, _ = self._fwd(tape, cv1, nm1, [a], N, h + 2 * p, w + 2 * p, 0 if p else None, ACT_RELU, k1, kn1,
                             out_pad=p, out_pad_mode=pm, drop=(0.5, self._new_seed()) if self.use_dropout else None)
            a, _ = self._fwd(tape, cv2, nm2, [t], N, h + 2 * p, w + 2 * p, 0 if p else None, ACT_NONE, k2, kn2,
                             residual=a.f32, want_f32=True, out_pad=0 if last else p, out_pad_mode=pm)
        for i in range(2):
            last = i == 1
            a, _ = self._fwd(tape, self.up[i], self.up_norm[i], [a], N, h, w, None, ACT_RELU, *K["up"][i],
                             out_pad=3 if last else 0, out_pad_mode=pm)
            h, w = h * 2, w * 2
        z, _ = self.head.run_tc([a], N, h + 6, w + 6, fuse_stats=False)
        out = ops.head_finish(z, self.head_bias, w, self.head_S, self.head_co, ACT_TANH)
        return out, dict(tape=tape, a_pad=a, out=out, N=N, H=H, W=W)

    def backward(self, ctx, dY, need_dx=False):
        """dY: fp32 NCHW gradient wrt the tanh output.  Returns ({state_dict key: gradient}, dx fp32 NCHW | None)."""
        tape, N, H, W = ctx["tape"], ctx["N"], ctx["H"], ctx["W"]
        refl = self.pad_mode == PAD_REFLECT
        grads = {}
        Y = ctx["out"]
        dzz = (dY * (1.0 - Y * Y)).contiguous()                     # tanh'  (3-channel image: glue)
        hk = self.keys["head"]
        grads[hk + ".bias"] = dzz.sum(dim=(0, 2, 3))
        S, co = self.head_S, self.head_co
        dzh, dzl = ops.head_bwd_pack(dzz, S, self.prec.fmt, self.prec.split)
        a_pad = ctx["a_pad"]
        G = self.head64.wgrad(a_pad, dzh, dzl, N, H + 6, W + 6, 0)                 # [64 j][ngf][R][1]
        ngf, R = G.shape[1], G.shape[2]
        grads[hk + ".weight"] = G[:, :, :, 0].reshape(16, 4, ngf, R)[:S, :co].permute(1, 2, 3, 0).contiguous()
        # data gradient wrt the head input: ConvTranspose(dz, wv64); zero padding = crop by 3 on every side (the
        # transposed conv's own padding), reflection padding = full padded extent folded back onto the interior
        dd = ops.conv_desc(N, H, W + 6, [64], ngf, R, 1, 1, 0 if refl else 3, True, 0)
        if getattr(self, "_head_wT", None) is None:
            self._head_wT = ops.pack_weights_tc(dd, self.head64.w_f32, self.prec.fmt, self.prec.split)
        dout = ops.conv_tc(dd, [dzh], [dzl], self._head_wT[0], self._head_wT[1], None, self.prec.fmt, self.prec.split)
        if refl:
            dout = ops.reflect_fold(dout, 3)
        # ---- up convs ------------------------------------------------------------------------------------------------
        n_rec = len(tape)
        i = n_rec - 1
        for _ in range(2):
            dout = self._bwd(tape[i], grads, dout); i -= 1
        # ---- ResNet blocks: x_{k+1} = x_k + n2(c2(relu(n1(c1(x_k))))) ---------------------------------------------------
        for _ in range(self.n_blocks):
            dt = self._bwd(tape[i], grads, dout); i -= 1            # through norm2 + conv2 -> d(relu(n1(.)))
            if refl:
                dt = ops.reflect_fold(dt, 1)
            dxb = self._bwd(tape[i], grads, dt); i -= 1             # through relu + norm1 + conv1 -> branch part of dx_k
            if refl:
                dout = ops.reflect_fold(dxb, 1, add=dout)           # padding backward + skip part
            else:
                dout, _, _ = ops.norm_apply(dxb, None, None, ACT_NONE, dout, want_f32=True, want_split=False)   # + skip part
        # ---- down convs + stem ----------------------------------------------------------------------------------------
        dout = self._bwd(tape[i], grads, dout); i -= 1
        dout = self._bwd(tape[i], grads, dout); i -= 1
        rec = tape[i]
        dxw = self._bwd(rec, grads, dout, need_dx=need_dx)          # wrt the horizontal-window operand [N,H+6,W,64]
        # stem weight: G[co][s*8 + c][r][0] -> w[co][c][r][s]
        G = grads[rec.wkey + ".weight"]
        cin = self.stem_in_nc
        grads[rec.wkey + ".weight"] = G[:, :, :, 0].reshape(G.shape[0], 8, 8, G.shape[2])[:, :self.stem_S, :cin] \
            .permute(0, 2, 3, 1).contiguous()
        dx = ops.stem_window_bwd(dxw, cin, 3, self.stem_S, self.pad_mode) if need_dx else None
        return grads, dx


class NLayerDTrainEngine(NLayerDEngine, _TrainOps):
    """PatchGAN forward/backward with gradient wrt the input (backward_G flows through the frozen D into G)."""

    def __init__(self, sd, **kw):
        super().__init__(sd, **kw)
        dev = self.device
        self.keys = [f"model.{2 + 3 * i}" for i in range(len(self.mid))]
        self.last_key = f"model.{2 + 3 * len(self.mid)}"
        wl64 = torch.zeros((64,) + tuple(self.last.w_f32.shape[1:]), dtype=torch.float32, device=dev)
        wl64[:32] = self.last.w_f32
        self.last64 = ConvLayer(wl64, None, stride=1, pad=1, prec=self.prec, backend="tc")

    def forward_train(self, x):
        x = x.contiguous()
        N, C, H, W = x.shape
        tape = []
        xh, xl = ops.stem_window_pack(x, 0, 1, PAD_ZERO, self.prec.fmt, self.prec.split)
        # first conv (+bias), LeakyReLU: no norm
        a, _ = self._fwd(tape, self.first, None, [Act(None, xh, xl)], N, H, W, None, ACT_LRELU02, "model.0", None)
        h, w = H // 2, W // 2
        for (cv, nm), key in zip(self.mid, self.keys):
            a, y = self._fwd(tape, cv, nm, [a], N, h, w, None, ACT_LRELU02, key, f"model.{int(key.split('.')[1]) + 1}")
            h, w = y.shape[1], y.shape[2]
        z, _ = self.last.run_tc([a], N, h, w, fuse_stats=False)
        out = ops.head_finish(z, self.last_bias, z.shape[2], 1, self.out_nc, ACT_NONE)
        return out, dict(tape=tape, a_last=a, hw=(h, w), N=N, C=C, H=H, W=W)

    def backward(self, ctx, dY, need_dx=True, param_grads=True):
        """dY: fp32 NCHW [N,1,h',w'] gradient wrt the logits.  Returns (grads, dx NCHW | None)."""
        tape, N, (h, w) = ctx["tape"], ctx["N"], ctx["hw"]
        grads = {}
        dY = dY.contiguous()
        dzh, dzl = ops.head_bwd_pack(dY, 1, self.prec.fmt, self.prec.split)        # lane 0 = dY
        if param_grads:
            grads[self.last_key + ".bias"] = dY.sum(dim=(0, 2, 3))
            G = self.last64.wgrad(ctx["a_last"], dzh, dzl, N, h, w, 1)             # [64][Cin][4][4]
            grads[self.last_key + ".weight"] = G[: self.out_nc].contiguous()
        dout = self.last64.dgrad(dzh, dzl, N, h, w, 1)
        for i in range(len(tape) - 1, -1, -1):
            first = i == 0
            if param_grads:
                dout = self._bwd(tape[i], grads, dout, need_dx=(not first) or need_dx)
            else:
                dout = self._bwd_data_only(tape[i], dout)
        if param_grads:
            G = grads["model.0.weight"]                                            # [ndf][64 padded lanes][4][4]
            grads["model.0.weight"] = G[:, : ctx["C"]].contiguous()
        dx = None
        if need_dx:
            dx = dout[..., : ctx["C"]].permute(0, 3, 1, 2).contiguous()            # padded lanes dropped, NHWC -> NCHW
        return grads, dx

    def _bwd_data_only(self, rec, dout):
        """Frozen discriminator (set_requires_grad(False) in optimize_parameters): only dx is needed."""
        N, H, W = rec.dims
        _, hi, lo = ops.norm_bwd(dout, rec.y, rec.sc, rec.sh, rec.mean, rec.rstd, rec.act, pooled=self._pooled(),
                                 want_f32=False, want_split=True, fmt=self.prec.fmt, need_lo=self.prec.split)
        return rec.layer.dgrad(hi, lo, N, H, W, rec.pad)


class UnetTrainEngine(_EngineBase, _TrainOps):
    """UnetGenerator forward with tape + backward (incl. the gradient wrt the input: the seg generators of the DeepLIIF
    cascade sit behind the modality generators, DeepLIIF_model.py:175-203, so dL/d(fake_B_i) flows through them).

    n_k = norm(d_k) is consumed twice — LeakyReLU into the next down conv and ReLU (skip) into this level's up conv —
    so its backward is one ``norm_bwd`` call with two (gradient, activation) pairs.  The skip concat is never
    materialised: the up-conv weight gradient is computed per source into row slices of dW, and the data gradient per
    source from the matching weight rows."""

    def __init__(self, sd, *, num_downs=9, norm="batch", norm_mode="batch", precision="bf16x3", device="cuda", use_dropout=False):
        # Dropout(0.5) closes the num_downs-5 inner ngf*8 blocks (networks.py:536, 604-605): levels 4 .. num_downs-2
        self.drop_levels = set(range(4, num_downs - 1)) if use_dropout else set()
        _EngineBase.__init__(self, norm, norm_mode, Precision.parse(precision) if isinstance(precision, str) else precision,
                             "tc", device)
        self.nd = num_downs
        g = lambda k: sd[k].to(device) if k in sd else None
        pre = ["model.model"]
        for lvl in range(1, num_downs):
            pre.append(f"{pre[-1]}.{1 if lvl == 1 else 3}.model")
        self.dkey, self.ukey, self.dnkey, self.unkey = [], [], [], []
        self.down, self.down_norm, self.up_w, self.up_b, self.up_norm, self.up_cins = [], [], [], [], [], []
        prec = self.prec
        for lvl in range(num_downs):
            p, inner = pre[lvl], lvl == num_downs - 1
            dk = f"{p}.0" if lvl == 0 else f"{p}.1"
            uk = f"{p}.3" if (lvl == 0 or inner) else f"{p}.5"
            self.dkey.append(dk); self.ukey.append(uk)
            wd = g(dk + ".weight").to(torch.float32)
            if lvl == 0:                                             # input_nc -> 64 zero lanes
                self.in_nc = wd.shape[1]
                wp = torch.zeros((wd.shape[0], 64, 4, 4), dtype=torch.float32, device=device); wp[:, : self.in_nc] = wd
                wd = wp
            self.down.append(ConvLayer(wd, g(dk + ".bias"), stride=2, pad=1, prec=prec, backend="tc"))
            has_dn = 0 < lvl < num_downs - 1
            self.dnkey.append(f"{p}.2" if has_dn else None)
            self.down_norm.append(_NormParams(sd, f"{p}.2", norm, device) if has_dn else None)
            wu = g(uk + ".weight").to(torch.float32)                 # ConvTranspose2d: (Cin_total, Cout, 4, 4)
            if lvl == 0:
                self.out_nc = wu.shape[1]
                wp = torch.zeros((wu.shape[0], 64, 4, 4), dtype=torch.float32, device=device); wp[:, : self.out_nc] = wu
                wu = wp
            self.up_w.append(wu); self.up_b.append(g(uk + ".bias"))
            ct = wu.shape[0]
            self.up_cins.append([ct] if inner else [ct // 2, ct // 2])
            nk = None if lvl == 0 else (f"{p}.4" if inner else f"{p}.6")
            self.unkey.append(nk)
            self.up_norm.append(_NormParams(sd, nk, norm, device) if nk else None)
        # per-source up-conv layers: forward uses one dual-source layer, backward one single-source layer per source
        self.up = [ConvLayer(self.up_w[l], None if l == 0 else self.up_b[l], transposed=True, stride=2, pad=1, cins=self.up_cins[l],
                             prec=prec, backend="tc", n_tile=0) for l in range(num_downs)]
        self.up_src = []
        for l in range(num_downs):
            off, layers = 0, []
            for c in self.up_cins[l]:
                layers.append(ConvLayer(self.up_w[l][off:off + c].contiguous(), None, transposed=True, stride=2, pad=1, prec=prec,
                                        backend="tc"))
                off += c
            self.up_src.append(layers)
        self.out_bias = self.up_b[0].detach().to(torch.float32).contiguous()

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward_train(self, x):
        x = x.contiguous()
        N, _, H, W = x.shape
        L = self.nd
        xh, xl = ops.stem_window_pack(x, 0, 1, PAD_ZERO, self.prec.fmt, self.prec.split)
        d_in, d_raw, d_stats, dims = [Act(None, xh, xl)], [], [], []
        h, w = H, W
        for lvl in range(L):
            if lvl > 0:
                sc, sh = d_stats[lvl - 1][0], d_stats[lvl - 1][1]
                d_in.append(self._apply(d_raw[lvl - 1], sc, sh, ACT_LRELU02))
            y, ws = self.down[lvl].run_tc([d_in[lvl]], N, h, w)
            dims.append((h, w))
            h, w = h // 2, w // 2
            d_raw.append(y)
            d_stats.append(self._stats(y, self.down_norm[lvl], ws, want_stats=True))
        u_src, u_raw, u_stats, u_drop = [None] * L, [None] * L, [None] * L, [None] * L
        below = None
        for lvl in range(L - 1, -1, -1):
            hh, ww = dims[lvl][0] // 2, dims[lvl][1] // 2            # spatial extent of d_lvl = input of the up conv
            sc, sh = d_stats[lvl][0], d_stats[lvl][1]
            skip = self._apply(d_raw[lvl], sc, sh, ACT_RELU)
            srcs = [skip] if lvl == L - 1 else [skip, below]
            u_src[lvl] = srcs
            y, ws = self.up[lvl].run_tc(srcs, N, hh, ww, fuse_stats=(lvl != 0))
            u_raw[lvl] = y
            if lvl == 0:
                z32 = y[..., :32].contiguous() if y.shape[3] != 32 else y
                out = ops.head_finish(z32, self.out_bias, 2 * ww, 1, self.out_nc, ACT_TANH)
            else:
                u_stats[lvl] = self._stats(y, self.up_norm[lvl], ws, want_stats=True)
                u_drop[lvl] = (0.5, self._new_seed()) if lvl in self.drop_levels else None
                # relu(dropout(norm(y))) == dropout(relu(norm(y))): the mask multiplier is non-negative
                below = self._apply(y, u_stats[lvl][0], u_stats[lvl][1], ACT_RELU, drop=u_drop[lvl])
        return out, dict(N=N, H=H, W=W, dims=dims, d_in=d_in, d_raw=d_raw, d_stats=d_stats, u_src=u_src, u_raw=u_raw,
                         u_stats=u_stats, u_drop=u_drop, out=out)

    # ---- backward --------------------------------------------------------------------------------------------------
    def _up_backward(self, lvl, ctx, grads, dy_hi, dy_lo, N, hh, ww):
        """wgrad + per-source dgrad of the up ConvTranspose at `lvl`; returns the list of source gradients (fp32 NHWC)."""
        srcs = ctx["u_src"][lvl]
        dws, dxs = [], []
        for layer, src in zip(self.up_src[lvl], srcs):
            dws.append(layer.wgrad(src, dy_hi, dy_lo, N, hh, ww))
            dxs.append(layer.dgrad(dy_hi, dy_lo, N, hh, ww))
        dW = torch.cat(dws, dim=0)
        if lvl == 0:
            dW = dW[:, : self.out_nc].contiguous()
        grads[self.ukey[lvl] + ".weight"] = dW
        return dxs

    def backward(self, ctx, dY, need_dx=True):
        N, L, dims = ctx["N"], self.nd, ctx["dims"]
        grads = {}
        Y = ctx["out"]
        dzz = (dY * (1.0 - Y * Y)).contiguous()
        grads[self.ukey[0] + ".bias"] = dzz.sum(dim=(0, 2, 3))
        dzh, dzl = ops.head_bwd_pack(dzz, 1, self.prec.fmt, self.prec.split)          # [N, H, W, 64 lanes]
        g_relu, g_below = [None] * L, [None] * (L + 1)
        # ---- up path, outermost -> innermost ------------------------------------------------------------------------
        hh, ww = dims[0][0] // 2, dims[0][1] // 2
        dxs = self._up_backward(0, ctx, grads, dzh, dzl, N, hh, ww)
        g_relu[0] = dxs[0]
        if L > 1:
            g_below[1] = dxs[1]
        for lvl in range(1, L):
            hh, ww = dims[lvl][0] // 2, dims[lvl][1] // 2
            sc, sh, mean, rstd = ctx["u_stats"][lvl]
            np_ = self.up_norm[lvl]
            dg = db = None
            if np_ is not None and np_.gamma is not None:
                dg, db = torch.empty_like(np_.gamma), torch.empty_like(np_.gamma)
            has_bias = self.up_b[lvl] is not None
            normed = sc is not None
            dp, dseed = ctx["u_drop"][lvl] if ctx["u_drop"][lvl] is not None else (0.0, 0)
            f32, hi, lo = ops.norm_bwd(g_below[lvl], ctx["u_raw"][lvl], sc, sh, mean, rstd, ACT_RELU, pooled=self._pooled(),
                                       dgamma=dg, dbeta=db, want_f32=has_bias and not normed, want_split=True,
                                       fmt=self.prec.fmt, need_lo=self.prec.split, drop_p=dp, drop_seed=dseed,
                                   drop_epoch=_engine.DROP_EPOCH[0] if dp > 0 else None)
            if dg is not None:
                grads[self.unkey[lvl] + ".weight"], grads[self.unkey[lvl] + ".bias"] = dg, db
            if has_bias:
                grads[self.ukey[lvl] + ".bias"] = torch.zeros_like(self.up_b[lvl]) if normed else ops.channel_sum(f32)
            dxs = self._up_backward(lvl, ctx, grads, hi, lo, N, hh, ww)
            g_relu[lvl] = dxs[0]
            if lvl < L - 1:
                g_below[lvl + 1] = dxs[1]
        # ---- down path, innermost -> outermost ------------------------------------------------------------------------
        g_lrelu = None
        for lvl in range(L - 1, -1, -1):
            h, w = dims[lvl]
            sc, sh, mean, rstd = ctx["d_stats"][lvl]
            np_ = self.down_norm[lvl]
            dg = db = None
            if np_ is not None and np_.gamma is not None:
                dg, db = torch.empty_like(np_.gamma), torch.empty_like(np_.gamma)
            layer = self.down[lvl]
            normed = sc is not None
            has_bias = layer.bias is not None
            want32 = has_bias and not normed
            if g_lrelu is None:       # innermost: d is consumed only through the skip ReLU
                f32, hi, lo = ops.norm_bwd(g_relu[lvl], ctx["d_raw"][lvl], sc, sh, mean, rstd, ACT_RELU, pooled=self._pooled(),
                                           dgamma=dg, dbeta=db, want_f32=want32, want_split=True, fmt=self.prec.fmt,
                                           need_lo=self.prec.split)
            else:
                f32, hi, lo = ops.norm_bwd(g_lrelu, ctx["d_raw"][lvl], sc, sh, mean, rstd, ACT_LRELU02, dout2=g_relu[lvl],
                                           act2=ACT_RELU, pooled=self._pooled(), dgamma=dg, dbeta=db, want_f32=want32,
                                           want_split=True, fmt=self.prec.fmt, need_lo=self.prec.split)
            if dg is not None:
                grads[self.dnkey[lvl] + ".weight"], grads[self.dnkey[lvl] + ".bias"] = dg, db
            if has_bias:
                grads[self.dkey[lvl] + ".bias"] = torch.zeros_like(layer.bias) if normed else ops.channel_sum(f32)
            dW = layer.wgrad(ctx["d_in"][lvl], hi, lo, N, h, w)
            if lvl == 0:
                dW = dW[:, : self.in_nc].contiguous()
            grads[self.dkey[lvl] + ".weight"] = dW
            if lvl > 0 or need_dx:
                g_lrelu = layer.dgrad(hi, lo, N, h, w)
        dx = g_lrelu[..., : self.in_nc].permute(0, 3, 1, 2).contiguous() if need_dx else None
        return grads, dx

"""Training step and loop: ``DeepLIIFModel.optimize_parameters`` and ``deepliif train`` on the sm_100a kernels.

Reference semantics kept (deepliif/models/DeepLIIF_model.py:205-467, cli.py:194-570):
  forward -> D step (GAN loss on detached fakes + reals, x0.5, weighted) -> G step (GAN + SmoothL1*lambda_L1, weighted)
  Adam(lr, betas=(beta1, 0.999)), linear-decay LambdaLR stepped per epoch, checkpoints ``{epoch}_net_{name}.pth``.
What is different underneath:
  * every network forward/backward runs through ``engine_train`` (tcgen05 forward, dgrad, wgrad; norm backward);
    PyTorch autograd only glues the tiny loss expressions (cat / BCE / MSE / SmoothL1 on 3-channel images and logits);
  * parameters and gradients of all generators (resp. all discriminators) live in ONE flat fp32 bucket each:
    the optimizer is a single fused-Adam launch per bucket and data parallelism is one NCCL all-reduce per bucket
    (the reference wraps each of its 18 nets in its own DistributedDataParallel reducer, networks.py:134);
  * rank 0's initial weights are broadcast (the reference re-initialises after the DDP wrap with per-rank seeds, so
    its replicas start from different weights — SURVEY.md §5; not reproduced);
  * the VGG perceptual term needs downloaded VGG19 weights and is outside the scope (north_star: L1 + GAN).
Limits: `--padding zero` (reflect-pad backward is not built); dropout masks come from our own counter-based generator,
not ATen's Philox stream (no implementation can reproduce those); seg generators of the cascade must be UNets (`--net-gs unet_*`, the reference default): ResNet generators do not return input gradients yet."""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import engine_train, ops
from .models import networks


# -------------------------------------------------------------------------------------------------------------------
# autograd bridge
# -------------------------------------------------------------------------------------------------------------------
class _NetFn(torch.autograd.Function):
    """y = net(x) with the forward tape kept by the training engine; backward returns dL/dx and dL/dparams."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.train_engine()
        out, tape = eng.forward_train(x.detach())
        ctx.module, ctx.eng, ctx.tape = module, eng, tape
        ctx.names = [n for n, _ in module.named_parameters()]
        ctx.x_needs = x.requires_grad
        return out

    @staticmethod
    def backward(ctx, dy):
        need_params = any(ctx.needs_input_grad[2:])
        eng, tape = ctx.eng, ctx.tape
        ctx.tape = ctx.eng = None            # ctx attributes are not released by autograd: drop the activations now
        if isinstance(eng, engine_train.NLayerDTrainEngine):
            grads, dx = eng.backward(tape, dy, need_dx=ctx.x_needs, param_grads=need_params)
        elif isinstance(eng, engine_train.UnetTrainEngine):
            grads, dx = eng.backward(tape, dy, need_dx=ctx.x_needs)
        else:
            grads, dx = eng.backward(tape, dy, need_dx=ctx.x_needs)
        tape.clear()
        out = [None, dx]
        for name, need in zip(ctx.names, ctx.needs_input_grad[2:]):
            out.append(grads.get(name) if (need and need_params) else None)
        return tuple(out)


def _train_engine(self):
    key = ("train",) + self._param_key()
    if getattr(self, "_tengine_key", None) != key:
        dev = next(self.parameters()).device
        sd = self.state_dict()
        if isinstance(self, networks.ResnetGenerator):
            self._tengine = engine_train.ResnetTrainEngine(sd, device=dev, precision=self.precision,
                                                          norm_mode="batch" if self.cfg["norm"] == "batch" else "sample",
                                                          **self.cfg)
        elif isinstance(self, networks.NLayerDiscriminator):
            self._tengine = engine_train.NLayerDTrainEngine(sd, device=dev, precision=self.precision,
                                                           norm_mode="batch" if self.cfg["norm"] == "batch" else "sample",
                                                           **self.cfg)
        elif isinstance(self, networks.UnetGenerator):
            self._tengine = engine_train.UnetTrainEngine(sd, device=dev, precision=self.precision,
                                                        norm_mode="batch" if self.cfg["norm"] == "batch" else "sample",
                                                        use_dropout=self.use_dropout, **self.cfg)
        else:
            raise NotImplementedError(f"training path for {type(self).__name__} is not built yet")
        self._tengine_key = key
    return self._tengine


def _forward_with_grad(self, input):
    if self.training and torch.is_grad_enabled():
        return _NetFn.apply(self, input.float(), *self.parameters())
    return networks._EngineBacked._inference_forward(self, input)


def install():
    """Give the engine-backed modules their training-mode forward (idempotent)."""
    if getattr(networks._EngineBacked, "_train_installed", False):
        return
    networks._EngineBacked._inference_forward = networks._EngineBacked.forward
    networks._EngineBacked.forward = _forward_with_grad
    networks._EngineBacked.train_engine = _train_engine
    networks._EngineBacked._train_installed = True


# -------------------------------------------------------------------------------------------------------------------
# flat buckets + fused Adam
# -------------------------------------------------------------------------------------------------------------------
class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics on one flat fp32 bucket (parameters and gradients are views into it)."""

    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8):
        params = [p for p in params]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        al = lambda k: (k + 63) // 64 * 64          # every tensor starts 256-byte aligned (kernels use 128-bit loads)
        n = sum(al(p.numel()) for p in params)
        dev = params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += al(k)
        self.t = 0
        # a process group of ANY size (also world 1: the 1-GPU NCCL self-test) routes the bucket through the collectives
        self.dist_on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if self.dist_on else 1
        self.hyper_dev = None          # set by GraphedStep: device float[4] the captured Adam launch reads

    def zero_grad(self, set_to_none=False):
        self.grad.zero_()

    def broadcast_from_rank0(self):
        if self.dist_on:
            dist.broadcast(self.flat, src=0)

    def all_reduce_grads(self):
        """One NCCL all-reduce (sum) of the whole gradient bucket; the 1/world factor is applied inside the Adam kernel.
        Enqueued on the current stream, so it is also recorded by a CUDA-graph capture of the step (GraphedStep)."""
        if self.dist_on:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)

    def hyper(self):
        """{lr, 1-beta1^t, sqrt(1-beta2^t), 1/world} of the *next* update (step t+1)."""
        g = self.param_groups[0]
        return ops.adam_hyper(g["lr"], g["betas"][0], g["betas"][1], self.t + 1, 1.0 / self.world)

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        if self.hyper_dev is not None:       # graph capture / replay: the host refreshes hyper_dev and counts the steps
            ops.adam_step_dev(self.flat, self.grad, self.m, self.v, self.hyper_dev, g["betas"][0], g["betas"][1], g["eps"])
        else:
            self.t += 1
            ops.adam_step(self.flat, self.grad, self.m, self.v, g["lr"], g["betas"][0], g["betas"][1], g["eps"], self.t,
                          1.0 / self.world)
        networks._EngineBacked.GLOBAL_VERSION += 1        # packed weights of every engine are now stale


# -------------------------------------------------------------------------------------------------------------------
# the step (DeepLIIF_model.py:205-467)
# -------------------------------------------------------------------------------------------------------------------
def _d_inputs(model, i, which):
    """Conditional-GAN discriminator input for seg head i: cat(condition, image) (DeepLIIF_model.py:252-262)."""
    S = model.mod_id_seg
    cond = model.real_A if i == 0 else getattr(model, f"real_B_{i}")
    return torch.cat((cond, getattr(model, f"{which}_B_{S}")), 1)


def _loss_D(model):
    opt, n, S = model.opt, model.opt.modalities_no, model.mod_id_seg
    total = torch.zeros((), device=model.device)
    for i, name in enumerate(model.model_names_d):
        D = model._net(name)
        fake = torch.cat((model.real_A, getattr(model, f"fake_B_{i + 1}").detach()), 1)
        real = torch.cat((model.real_A, getattr(model, f"real_B_{i + 1}")), 1)
        lf = model.criterionGAN_mod(D(fake), False)
        lr = model.criterionGAN_mod(D(real), True)
        setattr(model, f"loss_D_fake_{i + 1}", lf); setattr(model, f"loss_D_real_{i + 1}", lr)
        total = total + (lf + lr) * 0.5 * model.loss_D_weights[i]
    if model.seg_gen:
        pf = sum(model._net(nm)(_d_inputs(model, i, "fake").detach()) * model.seg_weights[i]
                 for i, nm in enumerate(model.model_names_ds))
        pr = sum(model._net(nm)(_d_inputs(model, i, "real")) * model.seg_weights[i]
                 for i, nm in enumerate(model.model_names_ds))
        lf, lr = model.criterionGAN_seg(pf, False), model.criterionGAN_seg(pr, True)
        setattr(model, f"loss_D_fake_{S}", lf); setattr(model, f"loss_D_real_{S}", lr)
        total = total + (lf + lr) * 0.5 * model.loss_D_weights[n]
    model.loss_D = total
    return total


def _loss_G(model):
    opt, n, S = model.opt, model.opt.modalities_no, model.mod_id_seg
    total = torch.zeros((), device=model.device)
    for i, name in enumerate(model.model_names_d):
        fake_B = getattr(model, f"fake_B_{i + 1}")
        lg = model.criterionGAN_mod(model._net(name)(torch.cat((model.real_A, fake_B), 1)), True)
        l1 = model.criterionSmoothL1(fake_B, getattr(model, f"real_B_{i + 1}")) * opt.lambda_L1
        setattr(model, f"loss_G_GAN_{i + 1}", lg); setattr(model, f"loss_G_L1_{i + 1}", l1)
        total = total + (lg + l1) * model.loss_G_weights[i]
    if model.seg_gen:
        pf = sum(model._net(nm)(_d_inputs(model, i, "fake")) * model.seg_weights[i] for i, nm in enumerate(model.model_names_ds))
        lg = model.criterionGAN_seg(pf, True)
        l1 = model.criterionSmoothL1(getattr(model, f"fake_B_{S}"), getattr(model, f"real_B_{S}")) * opt.lambda_L1
        setattr(model, f"loss_G_GAN_{S}", lg); setattr(model, f"loss_G_L1_{S}", l1)
        # the reference weights the seg term with loss_G_weights[i] of the stale loop index (= last modality)
        total = total + (lg + l1) * model.loss_G_weights[n - 1]
    model.loss_G = total
    return total


def deepliif_step(model):
    install()
    model.forward()
    d_nets = [model._net(nm) for nm in model.model_names_d + model.model_names_ds]
    model.set_requires_grad(d_nets, True)
    model.optimizer_D.zero_grad()
    _loss_D(model).backward()
    _sync_and_step(model.optimizer_D)
    model.set_requires_grad(d_nets, False)
    model.optimizer_G.zero_grad()
    _loss_G(model).backward()
    _sync_and_step(model.optimizer_G)


class GraphedStep:
    """One `optimize_parameters()` (forward, D step, G step, both Adam updates) captured in a CUDA graph and replayed per
    batch: the reference's default topology issues ~4400 launches per step at batch 1, a replay costs one (measured:
    126 -> 99 ms per step there, 243 -> 229 ms for the batch-8 flat-5 configuration).

    What varies between steps lives in device memory the graph reads: the input batch (static tensors refreshed by
    `copy_`), Adam's {lr, bias corrections} (FlatAdam.hyper_dev) and the dropout step counter (engine.DROP_EPOCH).
    The first `warmup` calls run eagerly (they also size every cache); the next full batch captures; later ones replay.
    A batch of another size (the ragged last one) runs eagerly against the same device-resident state."""
    SLOTS = 4          # pinned staging slots for the per-step state: the host may run this many steps ahead of the GPU

    def __init__(self, model, warmup=2):
        self.model, self.warmup, self.calls = model, warmup, 0
        self.graph = None
        self.opts = [model.optimizer_D, model.optimizer_G]
        if not all(isinstance(o, FlatAdam) for o in self.opts):
            raise NotImplementedError("GraphedStep needs the flat-bucket Adam optimizers (--optimizer adam)")
        n = 2 + 4 * len(self.opts)                                   # [dropout epoch (int64 as 2 floats) | 4 floats per Adam]
        self.host = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(self.SLOTS)]
        self.done = [None] * self.SLOTS
        self.state_dev = torch.zeros(n, dtype=torch.float32, device=model.device)
        self.epoch = 0
        self.static = None
        # Eager steps run on a side stream: autograd's AccumulateGrad nodes remember the stream they were created on, and
        # nodes born on the legacy default stream cannot be joined from a capturing stream.
        self.side = torch.cuda.Stream(device=model.device)

    def _eager(self, data):
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            self.model.set_input(data)
            self.model.optimize_parameters()
        cur.wait_stream(self.side)

    def _upload(self):
        self.epoch += 1
        k = self.epoch % self.SLOTS
        if self.done[k] is not None:
            self.done[k].synchronize()                              # slot still in flight: wait for that copy only
        h = self.host[k]
        h[:2].view(torch.int64)[0] = self.epoch
        for i, o in enumerate(self.opts):
            h[2 + 4 * i: 6 + 4 * i] = torch.tensor(o.hyper())
        self.state_dev.copy_(h, non_blocking=True)
        ev = torch.cuda.Event(); ev.record()
        self.done[k] = ev

    def _bind(self):
        from . import engine as _engine
        for i, o in enumerate(self.opts):
            o.hyper_dev = self.state_dev[2 + 4 * i: 6 + 4 * i]
        _engine.DROP_EPOCH[0] = self.state_dev[:2].view(torch.int64)

    @staticmethod
    def _shape_key(data):
        A = data["A"]
        return tuple(A[0].shape if isinstance(A, list) else A.shape)

    def __call__(self, data):
        model = self.model
        if self.calls < self.warmup:
            self._eager(data)
            self.calls += 1
            return
        dev = model.device
        A = [a.to(dev) for a in data["A"]] if isinstance(data["A"], list) else data["A"].to(dev)
        B = [b.to(dev) for b in data["B"]]
        paths = data.get("A_paths", [])
        self._bind()
        self._upload()
        if self.static is not None and self._shape_key(data) != self.static_key:
            self._eager({"A": A, "B": B, "A_paths": paths})         # ragged batch: eager, same device-resident state
        else:
            if self.static is None:
                self.static = {"A": [a.clone() for a in A] if isinstance(A, list) else A.clone(), "B": [b.clone() for b in B]}
                self.static_key = self._shape_key(data)
                model.set_input({**self.static, "A_paths": paths})
                torch.cuda.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    model.optimize_parameters()
                # the step's result tensors (losses, visuals) are static outputs of the graph: an eager step in between
                # rebinds the model attributes, so they are put back after every replay
                self.outputs = {k: v for k, v in vars(model).items()
                                if torch.is_tensor(v) and k.startswith(("loss_", "fake_B", "real_"))}
            else:
                if isinstance(A, list):
                    for d, s_ in zip(self.static["A"], A):
                        d.copy_(s_, non_blocking=True)
                else:
                    self.static["A"].copy_(A, non_blocking=True)
                for d, s_ in zip(self.static["B"], B):
                    d.copy_(s_, non_blocking=True)
            model.image_paths = paths
            self.graph.replay()
            for k, v in self.outputs.items():
                setattr(model, k, v)
        for o in self.opts:
            o.t += 1
        self.calls += 1


def _sync_and_step(optimizer):
    if isinstance(optimizer, FlatAdam):
        optimizer.all_reduce_grads()
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        # generic torch optimizers (--optimizer other than adam): average every gradient across the ranks, the job
        # DistributedDataParallel does in the reference (networks.py:131-136)
        world = dist.get_world_size()
        for group in optimizer.param_groups:
            for p_ in group["params"]:
                if p_.grad is not None:
                    dist.all_reduce(p_.grad)
                    p_.grad.div_(world)
    optimizer.step()
    if not isinstance(optimizer, FlatAdam):
        networks._EngineBacked.GLOBAL_VERSION += 1


# -------------------------------------------------------------------------------------------------------------------
# `deepliif train`
# -------------------------------------------------------------------------------------------------------------------
def default_weights(model, modalities_no, default):
    """cli.py:349-355: the 4-modality DeepLIIF model has hand-set defaults, every other size gets equal weights."""
    if model in ("DeepLIIF", "DeepLIIFKD") and modalities_no == 4:
        return list(default)
    if model in ("DeepLIIF", "DeepLIIFKD"):
        return [1 / (modalities_no + 1)] * (modalities_no + 1)
    return [1 / modalities_no] * modalities_no


def _weights_arg(v, model, n, default):
    if v is None or (isinstance(v, str) and len(v) == 0):
        return default_weights(model, n, default)
    return [float(x) for x in v.split(",")] if isinstance(v, str) else [float(x) for x in v]


def prepare_train_params(kw):
    """The option prologue of the reference's `deepliif train` (cli.py:213-380), same rules in the same order: seg_no from
    seg_gen, validation only with a seg generator, zero padding when a seed asks for determinism, input_no / scale_size
    read off the first training image, modalities names, background colours of empty tiles, one architecture per
    generator, seg / loss weights (defaults, comma-separated strings, sanity checks) -> the dict Options() is built from."""
    from PIL import Image
    from .util import infer_background_colors
    d = dict(kw)
    model, n = d["model"], d["modalities_no"]
    if model not in ("DeepLIIF",):
        raise NotImplementedError(f"model class {model} is outside the B200 hot-path scope (only DeepLIIF is built)")
    d["seg_no"] = 1 if d["seg_gen"] else 0
    if not d["seg_gen"]:
        d["with_val"] = False
    if d.get("optimizer", "adam") != "adam":
        print(f"Optimizer torch.optim.{d['optimizer']} is not tested. Be careful about the parameters of the optimizer.")
    gpu_ids = tuple(d.get("gpu_ids") or ())
    if gpu_ids and gpu_ids[0] == -1:
        raise SystemExit("deepliif_b200 has no CPU path: --gpu-ids -1 is not available")
    if d.get("seed") is not None:
        d["padding"] = "zero"
        print("padding type is forced to zero padding, because neither refection pad2d or replication pad2d has a "
              "deterministic implementation")
    dir_train = os.path.join(d["dataroot"], "train")
    fns = sorted(x for x in os.listdir(dir_train) if x.endswith(".png"))
    print(f"{len(fns)} images found")
    img = Image.open(os.path.join(dir_train, fns[0]))
    print("image shape:", img.size)
    num_img = img.size[0] / img.size[1]
    assert int(num_img) == num_img, f"img size {img.size[0]} / {img.size[1]} = {num_img} is not an integer"
    input_no = int(num_img) - n - d["seg_no"]
    assert input_no > 0, (f"inferred number of input images is {input_no} (modalities_no {n}, seg_no {d['seg_no']}); "
                          "should be greater than 0")
    names = d.get("modalities_names") or ""
    names = [x.strip() for x in names.split(",") if len(x) > 0] if isinstance(names, str) else list(names)
    assert len(names) == 0 or len(names) == input_no + n, \
        f"--modalities-names has {len(names)} entries ({names}), expecting 0 or {input_no + n} entries"
    d.update(input_no=input_no, modalities_names=names, scale_size=img.size[1], gpu_ids=gpu_ids, lambda_identity=0, pool_size=0)
    if d["seg_gen"]:
        colors = infer_background_colors(dir_train, sample_size=10, input_no=input_no, modalities_no=n, seg_no=d["seg_no"],
                                         tile_size=32, return_list=True)
        if colors is not None:
            d["background_colors"] = colors
    net_g = d["net_g"].split(",") if isinstance(d["net_g"], str) else list(d["net_g"])
    assert len(net_g) in (1, n), ("net_g should contain either 1 architecture for all translation generators or the same "
                                  f"number of architectures as the number of translation generators ({n})")
    net_gs = d["net_gs"].split(",") if isinstance(d["net_gs"], str) else list(d["net_gs"])
    assert len(net_gs) in (1, d["seg_no"]), ("net_gs should contain either 1 architecture for all segmentation generators "
                                             f"or the same number of architectures as the number of segmentation generators ({d['seg_no']})")
    d["net_g"] = net_g * n if len(net_g) == 1 else net_g
    d["net_gs"] = net_gs * (n + d["seg_no"]) if len(net_gs) == 1 else net_gs
    seg_w = _weights_arg(d.get("seg_weights"), model, n, [0.25, 0.15, 0.25, 0.1, 0.25])
    lw_g = _weights_arg(d.pop("loss_weights_g", None), model, n, [0.2] * 5)
    lw_d = _weights_arg(d.pop("loss_weights_d", None), model, n, [0.2] * 5)
    assert sum(seg_w) == 1, "seg weights should add up to 1"
    assert sum(lw_g) == 1, "loss weights g should add up to 1"
    assert sum(lw_d) == 1, "loss weights d should add up to 1"
    for name, w in (("seg weights", seg_w), ("loss weights g", lw_g), ("loss weights d", lw_d)):
        assert len(w) == n + 1, f"{name} should have the same number of elements as number of modalities to be generated"
    d.update(seg_weights=seg_w, loss_G_weights=lw_g, loss_D_weights=lw_d)
    return d


def build_options(params):
    from .options import Options
    p = dict(params)
    n = p["modalities_no"]
    targets = n + (1 if p["seg_gen"] else 0)
    p.setdefault("seg_no", 1 if p["seg_gen"] else 0)
    p["seg_weights"] = p.get("seg_weights") or [1 / (n + 1)] * (n + 1)
    p["loss_G_weights"] = p.pop("loss_weights_g", None) or p.get("loss_G_weights") or [1 / targets] * targets
    p["loss_D_weights"] = p.pop("loss_weights_d", None) or p.get("loss_D_weights") or [1 / targets] * targets
    p.setdefault("modalities_names", [f"mod{i}" for i in range(n + 1)])
    opt = Options(d_params=p, mode="train")
    opt.gpu_ids = list(p.get("gpu_ids") or [int(os.environ.get("LOCAL_RANK", "0"))])
    opt.netD = p.get("net_d", "n_layers")
    if opt.netD == "n_layers":
        opt.n_layers_D = 4
    return opt


def make_optimizers(model):
    """Replace the model's optimizers by flat-bucket fused Adam when --optimizer adam (the default)."""
    opt = model.opt
    if str(opt.optimizer).lower() != "adam":
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for nm in model.model_names:                 # same initial weights on every rank (DDP's constructor broadcast)
                for t in list(model._net(nm).parameters()) + list(model._net(nm).buffers()):
                    dist.broadcast(t.data, src=0)
            networks._EngineBacked.GLOBAL_VERSION += 1
        return
    g = [p for nm in model.model_names_g + model.model_names_gs for p in model._net(nm).parameters()]
    d = [p for nm in model.model_names_d + model.model_names_ds for p in model._net(nm).parameters()]
    model.optimizer_G = FlatAdam(g, lr=opt.lr_g, betas=(opt.beta1, 0.999))
    model.optimizer_D = FlatAdam(d, lr=opt.lr_d, betas=(opt.beta1, 0.999))
    model.optimizers = [model.optimizer_G, model.optimizer_D]
    model.optimizer_G.broadcast_from_rank0(); model.optimizer_D.broadcast_from_rank0()


def run_training(params):
    from .models import create_model
    from .options import print_options
    install()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("deepliif train needs a CUDA (sm_100a) device: there is no CPU fallback")
    # reference cli.py:247-256: under torchrun every rank uses gpu_ids[LOCAL_RANK] and then sees it as its only GPU
    ids = [g for g in (params.get("gpu_ids") or ()) if g is not None and int(g) >= 0]
    if "LOCAL_RANK" in os.environ and world > 1:
        dev_index = int(ids[local]) if len(ids) > local else local
    else:
        dev_index = int(ids[0]) if ids else local
    torch.cuda.set_device(dev_index)
    params = dict(params, gpu_ids=(dev_index,))
    local = dev_index
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev_index))
    if params.get("seed") is not None:
        torch.manual_seed(params["seed"]); np.random.seed(params["seed"])
    opt = build_options(params)
    if rank == 0:
        print_options(opt, save=True)
    import random
    from .data.aligned_dataset import AlignedDataset, DeviceBatches, collate_u8
    if params.get("seed") is not None:
        random.seed(params["seed"] + rank)            # crop / flip parameters (base_dataset.py:62-78 uses `random`)
    ds = AlignedDataset(opt, "train")                 # uint8 tiles; ToTensor + Normalize run on the device
    sampler = torch.utils.data.distributed.DistributedSampler(ds, world, rank, shuffle=True) if world > 1 else None
    loader = torch.utils.data.DataLoader(ds, batch_size=opt.batch_size, shuffle=sampler is None and not opt.serial_batches,
                                         sampler=sampler, num_workers=opt.num_threads, collate_fn=collate_u8, drop_last=False,
                                         pin_memory=True)
    dl = DeviceBatches(loader, torch.device("cuda", local), input_no=getattr(opt, "input_no", 1))
    model = create_model(opt)
    model.setup(opt)
    if getattr(opt, "precision", "bf16x3") != "bf16x3":           # --precision bf16: single-pass operands (not fp32-parity)
        for nm in model.model_names:
            model._unwrap(model._net(nm)).precision = opt.precision
    make_optimizers(model)
    from .models import networks as nw
    model.schedulers = [nw.get_scheduler(o, opt) for o in model.optimizers]
    model.train()
    stepper = GraphedStep(model) if params.get("cuda_graph") else None
    # ---- the reference's loop (cli.py:404-572): epochs epoch_count .. n_epochs + n_epochs_decay inclusive, counters in
    # images, `latest` every save_latest_freq images, an epoch checkpoint every save_epoch_freq epochs, lr step per epoch.
    # Validation (--with-val) and the visdom / html dashboards are not part of this package.
    if getattr(opt, "with_val", False) and rank == 0:
        print("--with-val: validation metrics need the reference's dashboard stack and are skipped here")
    total_iters = 0
    epoch_base = 0
    if getattr(opt, "continue_train", False):
        try:
            epoch_base = int(opt.epoch)
        except (TypeError, ValueError):
            epoch_base = 0
    log_name = os.path.join(opt.checkpoints_dir, opt.name, "loss_log.txt")
    if rank == 0:
        os.makedirs(os.path.dirname(log_name), exist_ok=True)
        with open(log_name, "a") as f:
            f.write("================ Training Loss (%s) ================\n" % time.strftime("%c"))
    last_epoch = opt.n_epochs + opt.n_epochs_decay
    for epoch in range(opt.epoch_count, last_epoch + 1):
        if sampler is not None and not opt.serial_batches:
            sampler.set_epoch(epoch)
        t0 = iter_data_time = time.time()
        epoch_iter, t_data = 0, 0.0
        for data in dl:
            iter_start_time = time.time()
            if total_iters % opt.print_freq == 0:
                t_data = iter_start_time - iter_data_time
            total_iters += opt.batch_size
            epoch_iter += opt.batch_size
            if stepper is not None:
                stepper(data)
            else:
                model.set_input(data)
                model.optimize_parameters()
            if rank == 0 and total_iters % opt.print_freq == 0:
                losses = model.get_current_losses()
                t_comp = (time.time() - iter_start_time) / opt.batch_size
                message = "(epoch: %d, iters: %d, time: %.3f, data: %.3f) " % (epoch, epoch_iter, t_comp, t_data)
                message += "".join("%s: %.3f " % kv for kv in losses.items())
                print(message, flush=True)
                with open(log_name, "a") as f:
                    f.write("%s\n" % message)
            if rank == 0 and total_iters % opt.save_latest_freq == 0:
                print("saving the latest model (epoch %d, total_iters %d)" % (epoch, total_iters))
                model.save_networks("iter_%d" % total_iters if opt.save_by_iter else "latest")
            iter_data_time = time.time()
            if getattr(opt, "debug", False) and epoch_iter >= opt.debug_data_size:
                print(f"debug mode, epoch {epoch} stopped at epoch iter {epoch_iter} (>= {opt.debug_data_size})")
                break
        if rank == 0 and epoch % opt.save_epoch_freq == 0 and not (getattr(opt, "continue_train", False) and epoch == 0):
            print("saving the model at the end of epoch %d, iters %d" % (epoch, total_iters))
            model.save_networks("latest")
            model.save_networks(epoch + epoch_base)
        if rank == 0:
            print("End of epoch %d / %d \t Time Taken: %d sec" % (epoch, last_epoch, time.time() - t0))
        model.update_learning_rate()
    if world > 1:
        dist.barrier()

"""GPU parity of the training path against torch autograd on the oracle networks (CPU, fp32):
network-level gradients (ResNet generator, PatchGAN incl. input gradient) and one full DeepLIIF optimisation step
(losses and post-step weights), BASELINE config-4 style: ResNet generators + PatchGAN, L1 + GAN, Adam."""
import copy
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nets

pytestmark = pytest.mark.gpu


def _rand(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * 2 - 1


def _leafify(sd):
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd.items()}


def _cmp(grads, ref_sd, tol_l2=2e-2):
    """Relative L2 error per gradient tensor.  A max-abs gate is ill-posed here: the (Leaky)ReLU derivative is a step,
    so a pre-activation within ~1e-4 of zero (the forward tolerance) flips one mask element, which moves single
    weight-gradient entries by a few % at these tiny test extents (K = a few hundred pixels) — in *any* two fp32
    implementations.  The L2 norm over the tensor is insensitive to those isolated entries; the max is printed."""
    worst, worst_max = 0.0, 0.0
    gmax = max(ref_sd[k].grad.abs().max().item() for k in grads)
    for k, g in grads.items():
        r = ref_sd[k].grad
        assert r is not None, k
        d = (g.cpu() - r)
        if r.abs().max().item() < 1e-4 * gmax:
            # analytically-zero gradients (a conv bias in front of an affine-free InstanceNorm): only rounding noise
            assert d.abs().max().item() < 1e-4 * gmax, (k, d.abs().max().item())
            continue
        rel = d.norm().item() / max(r.norm().item(), 1e-12)
        worst, worst_max = max(worst, rel), max(worst_max, d.abs().max().item() / max(r.abs().max().item(), 1e-12))
        assert rel < tol_l2, (k, rel)
    print(f"   (worst max-abs/scale over tensors: {worst_max:.2e})")
    return worst


@pytest.mark.parametrize("norm,norm_mode,padding", [("instance", "sample", "zero"), ("batch", "batch", "zero"),
                                                    ("batch", "batch", "reflect"), ("instance", "sample", "reflect")])
def test_resnet_generator_gradients(norm, norm_mode, padding):
    """Parameter gradients and the input gradient (the seg generators of the cascade sit behind the modality generators;
    define_G's default padding_type is 'reflect', networks.py:142-144)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine_train
    cfg = dict(n_blocks=2, norm=norm, use_dropout=False, padding_type=padding)
    sd = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 2, norm, False, padding), 7, "stress")
    x = _rand((2, 3, 64, 64), 70).requires_grad_(True)
    dY = _rand((2, 3, 64, 64), 71)
    leaf = _leafify(sd)
    y_ref = nets.resnet_forward(x, leaf, norm_mode=norm_mode, **cfg)
    (y_ref * dY).sum().backward()
    eng = engine_train.ResnetTrainEngine(sd, norm_mode=norm_mode, precision="bf16x3", **cfg)
    y, ctx = eng.forward_train(x.detach().cuda())
    assert (y.cpu() - y_ref.detach()).abs().max().item() < 1e-3
    grads, dx = eng.backward(ctx, dY.cuda(), need_dx=True)
    expected = {k for k, v in leaf.items() if isinstance(v, torch.Tensor) and v.requires_grad}
    assert set(grads) == expected, set(grads) ^ expected
    worst = _cmp(grads, leaf)
    rel_dx = (dx.cpu() - x.grad).norm().item() / x.grad.norm().item()
    print(f"resnet {norm}/{norm_mode}/{padding}: worst relative grad error {worst:.2e} over {len(grads)} tensors; "
          f"input gradient rel-L2 {rel_dx:.2e}")
    assert rel_dx < 2e-2


@pytest.mark.parametrize("norm,n_layers", [("instance", 3), ("batch", 3), ("batch", 4)])
def test_discriminator_gradients_and_input_gradient(norm, n_layers):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine_train
    sd = nets.make_state_dict(nets.nlayer_d_param_shapes(n_layers, 64, 6, norm), 8, "stress")
    x = _rand((2, 6, 128, 128), 80).requires_grad_(True)
    leaf = _leafify(sd)
    y_ref = nets.nlayer_d_forward(x, leaf, n_layers=n_layers, norm=norm, norm_mode="batch")
    dY = _rand(tuple(y_ref.shape), 81)
    (y_ref * dY).sum().backward()
    eng = engine_train.NLayerDTrainEngine(sd, n_layers=n_layers, norm=norm, norm_mode="batch" if norm == "batch" else "sample")
    y, ctx = eng.forward_train(x.detach().cuda())
    assert (y.cpu() - y_ref.detach()).abs().max().item() < 1e-3 * max(1.0, y_ref.abs().max().item())
    grads, dx = eng.backward(ctx, dY.cuda(), need_dx=True)
    worst = _cmp(grads, leaf)
    errx = (dx.cpu() - x.grad).norm().item() / x.grad.norm().item()
    print(f"D {norm}/n{n_layers}: worst relative-L2 param-grad error {worst:.2e}, input-grad rel-L2 error {errx:.2e}")
    assert errx < 5e-3
    # frozen-D path (backward_G): same dx, no parameter gradients
    y2, ctx2 = eng.forward_train(x.detach().cuda())
    g2, dx2 = eng.backward(ctx2, dY.cuda(), need_dx=True, param_grads=False)
    assert g2 == {} and (dx2 - dx).abs().max().item() < 1e-6


def _oracle_step(sds_g, sds_d, A, Bs, cfg, lr=2e-4, beta1=0.5, lambda_l1=100.0, norm="instance"):
    """One optimize_parameters() of the reference semantics on the oracle nets (torch autograd, CPU)."""
    n = len(sds_g)
    G = [_leafify(sd) for sd in sds_g]
    D = [_leafify(sd) for sd in sds_d]
    params_g = [v for sd in G for v in sd.values() if isinstance(v, torch.Tensor) and v.requires_grad]
    params_d = [v for sd in D for v in sd.values() if isinstance(v, torch.Tensor) and v.requires_grad]
    og = torch.optim.Adam(params_g, lr=lr, betas=(beta1, 0.999))
    od = torch.optim.Adam(params_d, lr=lr, betas=(beta1, 0.999))
    nm = "batch" if norm == "batch" else "sample"
    fwdG = lambda sd, x: nets.resnet_forward(x, sd, norm_mode=nm, **cfg)
    fwdD = lambda sd, x: nets.nlayer_d_forward(x, sd, n_layers=4, norm=norm, norm_mode="batch")
    bce = torch.nn.BCEWithLogitsLoss(); sl1 = torch.nn.SmoothL1Loss()
    fakes = [fwdG(G[i], A) for i in range(n)]
    w = 1.0 / n
    lossD = 0
    losses = {}
    for i in range(n):
        pf = fwdD(D[i], torch.cat((A, fakes[i].detach()), 1)); pr = fwdD(D[i], torch.cat((A, Bs[i]), 1))
        lf, lr_ = bce(pf, torch.zeros_like(pf)), bce(pr, torch.ones_like(pr))
        losses[f"D_fake_{i + 1}"], losses[f"D_real_{i + 1}"] = lf.item(), lr_.item()
        lossD = lossD + (lf + lr_) * 0.5 * w
    od.zero_grad(); lossD.backward(); od.step()
    lossG = 0
    for i in range(n):
        pf = fwdD(D[i], torch.cat((A, fakes[i]), 1))
        lg, l1 = bce(pf, torch.ones_like(pf)), sl1(fakes[i], Bs[i]) * lambda_l1
        losses[f"G_GAN_{i + 1}"], losses[f"G_L1_{i + 1}"] = lg.item(), l1.item()
        lossG = lossG + (lg + l1) * w
    for p in params_d:
        p.requires_grad_(False)
    og.zero_grad(); lossG.backward(); og.step()
    return G, D, losses


@pytest.mark.parametrize("norm", ["instance", "batch"])
def test_one_optimisation_step_matches_oracle(norm, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import training
    from deepliif_b200.cli import TRAIN_DEFAULTS
    from deepliif_b200.models import create_model
    p = dict(TRAIN_DEFAULTS, dataroot=str(tmp_path), checkpoints_dir=str(tmp_path), name="t", gpu_ids=(0,), modalities_no=2,
             seg_gen=False, norm=norm, no_dropout=True, padding="zero", net_g="resnet_3blocks", batch_size=2)
    opt = training.build_options(p)
    torch.manual_seed(0)
    model = create_model(opt)
    sds_g = [copy.deepcopy({k: v.detach().cpu() for k, v in model._net(f"G{i + 1}").module.state_dict().items()}) for i in range(2)]
    sds_d = [copy.deepcopy({k: v.detach().cpu() for k, v in model._net(f"D{i + 1}").module.state_dict().items()}) for i in range(2)]
    training.make_optimizers(model)
    model.train()
    A = _rand((2, 3, 64, 64), 90); Bs = [_rand((2, 3, 64, 64), 91 + i) for i in range(2)]
    model.set_input({"A": A, "B": Bs, "A_paths": ["a", "b"]})
    model.optimize_parameters()
    torch.cuda.synchronize()
    got = model.get_current_losses()
    cfg = dict(n_blocks=3, norm=norm, use_dropout=False, padding_type="zero")
    torch.set_num_threads(min(32, os.cpu_count()))
    G, D, want = _oracle_step(sds_g, sds_d, A, Bs, cfg, norm=norm)
    for k, v in want.items():
        print(f"loss {k}: ours {got[k]:.6f} oracle {v:.6f}")
        assert abs(got[k] - v) <= 2e-3 * max(1.0, abs(v)), k
    # post-step weights.  Adam's first step is lr*g/(|g|+eps) ~ lr*sign(g): a weight whose true gradient is ~0 (biases
    # cancelled by a norm layer, dead ReLU paths) gets +-lr from rounding noise in ANY implementation, so the gate is
    # (a) the update vectors point the same way and (b) sign disagreements stay a small minority.
    lr = 2e-4
    for i in range(2):
        for tag, ref_sd, w0, name in (("G", G[i], sds_g[i], f"G{i + 1}"), ("D", D[i], sds_d[i], f"D{i + 1}")):
            ours = model._net(name).module.state_dict()
            n_bad = n_all = 0
            dot = na = nb = 0.0
            for k, r in ref_sd.items():
                if not (isinstance(r, torch.Tensor) and r.dtype.is_floating_point) or "running" in k:
                    continue
                uo, ur = ours[k].cpu() - w0[k], r.detach() - w0[k]
                d = (uo - ur).abs()
                n_all += d.numel(); n_bad += int((d > 0.5 * lr).sum())
                dot += float((uo * ur).sum()); na += float((uo * uo).sum()); nb += float((ur * ur).sum())
            frac, cos = n_bad / n_all, dot / (na * nb) ** 0.5
            print(f"{tag}{i + 1}: Adam-step cosine {cos:.4f}; weights off by > lr/2: {100 * frac:.3f}% of {n_all}")
            assert cos > 0.95 and frac < 0.06


@pytest.mark.parametrize("norm,norm_mode,nd,hw", [("instance", "sample", 5, 64), ("batch", "batch", 6, 128), ("batch", "batch", 9, 512)])
def test_unet_generator_gradients_and_input_gradient(norm, norm_mode, nd, hw):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine_train
    sd = nets.make_state_dict(nets.unet_param_shapes(nd, 64, 3, 3, norm), 9, "stress")
    n = 1 if hw == 512 else 2
    x = _rand((n, 3, hw, hw), 95).requires_grad_(True)
    dY = _rand((n, 3, hw, hw), 96)
    leaf = _leafify(sd)
    torch.set_num_threads(min(32, os.cpu_count()))
    y_ref = nets.unet_forward(x, leaf, num_downs=nd, norm=norm, norm_mode=norm_mode)
    (y_ref * dY).sum().backward()
    eng = engine_train.UnetTrainEngine(sd, num_downs=nd, norm=norm, norm_mode=norm_mode)
    y, ctx = eng.forward_train(x.detach().cuda())
    assert (y.cpu() - y_ref.detach()).abs().max().item() < 1e-3
    grads, dx = eng.backward(ctx, dY.cuda(), need_dx=True)
    expected = {k for k, v in leaf.items() if isinstance(v, torch.Tensor) and v.requires_grad}
    assert set(grads) == expected, set(grads) ^ expected
    worst = _cmp(grads, leaf, tol_l2=3e-2)
    errx = (dx.cpu() - x.grad).norm().item() / x.grad.norm().item()
    print(f"unet{nd} {norm}/{norm_mode} @{hw}: worst rel-L2 param-grad error {worst:.2e}, input-grad rel-L2 {errx:.2e}")
    assert errx < 1e-2


def test_default_cascade_step_runs_and_matches_oracle_losses(tmp_path):
    """The reference's default topology (ResNet modality generators, UNet seg generators behind them, n_layers=4
    PatchGANs, seg_gen=True) at reduced size: one optimize_parameters(); the D/G losses of the first step only
    depend on the forward pass, so they must match the oracle cascade."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import training
    from deepliif_b200.cli import TRAIN_DEFAULTS
    from deepliif_b200.models import create_model
    p = dict(TRAIN_DEFAULTS, dataroot=str(tmp_path), checkpoints_dir=str(tmp_path), name="c", gpu_ids=(0,), modalities_no=2,
             seg_gen=True, norm="batch", no_dropout=True, padding="zero", net_g="resnet_2blocks", net_gs="unet_128", batch_size=2)
    opt = training.build_options(p)
    torch.manual_seed(1)
    model = create_model(opt)
    sd = {nm: {k: v.detach().cpu().clone() for k, v in model._net(nm).module.state_dict().items()} for nm in model.model_names}
    training.make_optimizers(model)
    model.train()
    A = _rand((2, 3, 128, 128), 60); Bs = [_rand((2, 3, 128, 128), 61 + i) for i in range(3)]
    model.set_input({"A": A, "B": Bs, "A_paths": []})
    model.optimize_parameters()
    torch.cuda.synchronize()
    got = model.get_current_losses()
    # oracle forward + first-step losses (BatchNorm in training mode: pooled statistics)
    cfg = dict(n_blocks=2, norm="batch", use_dropout=False, padding_type="zero", norm_mode="batch")
    with torch.no_grad():
        fk = [nets.resnet_forward(A, sd[f"G{i + 1}"], **cfg) for i in range(2)]
        us = lambda t, s: nets.unet_forward(t, s, num_downs=7, norm="batch", norm_mode="batch")
        parts = [us(A, sd["GS0"]), us(fk[0], sd["GS1"]), us(fk[1], sd["GS2"])]
        w = opt.seg_weights
        seg = sum(p_ * w_ for p_, w_ in zip(parts, w))
        D = lambda t, s: nets.nlayer_d_forward(t, s, n_layers=4, norm="batch", norm_mode="batch")
        bce = torch.nn.BCEWithLogitsLoss(); mse = torch.nn.MSELoss(); sl1 = torch.nn.SmoothL1Loss()
        want = {}
        for i in range(2):
            pf = D(torch.cat((A, fk[i]), 1), sd[f"D{i + 1}"]); pr = D(torch.cat((A, Bs[i]), 1), sd[f"D{i + 1}"])
            want[f"D_fake_{i + 1}"] = bce(pf, torch.zeros_like(pf)).item(); want[f"D_real_{i + 1}"] = bce(pr, torch.ones_like(pr)).item()
            want[f"G_L1_{i + 1}"] = (sl1(fk[i], Bs[i]) * 100).item()
        conds = [A, Bs[0], Bs[1]]
        pf = sum(D(torch.cat((c, seg), 1), sd[f"DS{i}"]) * w[i] for i, c in enumerate(conds))
        pr = sum(D(torch.cat((c, Bs[2]), 1), sd[f"DS{i}"]) * w[i] for i, c in enumerate(conds))
        want["D_fake_S"] = mse(pf, torch.zeros_like(pf)).item(); want["D_real_S"] = mse(pr, torch.ones_like(pr)).item()
        want["G_L1_S"] = (sl1(seg, Bs[2]) * 100).item()
    for k, v in want.items():
        print(f"loss {k}: ours {got[k]:.6f} oracle {v:.6f}")
        assert abs(got[k] - v) <= 2e-3 * max(1.0, abs(v)), k
    assert all(torch.isfinite(p_).all() for nm in model.model_names for p_ in model._net(nm).parameters())


def dropout_mask_numpy(seed, n, p=0.5):
    """numpy twin of dlb::dropout_scale (csrc/rng.cuh): multiplier per element index."""
    idx = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(32)).astype(np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.where(u >= p, np.float32(1.0 / (1.0 - p)), np.float32(0.0))


def test_dropout_forward_backward_share_the_mask():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import ops
    N, H, W, C, seed = 2, 8, 8, 64, 1234567
    y = _rand((N, H, W, C), 1).cuda()
    sc, sh, mean, rstd = ops.norm_stats(y, None, None, False, want_stats=True)
    f32, _, _ = ops.norm_apply(y, sc, sh, ops.ACT_RELU, want_f32=True, want_split=False, drop_p=0.5, drop_seed=seed)
    base, _, _ = ops.norm_apply(y, sc, sh, ops.ACT_RELU, want_f32=True, want_split=False)
    m = torch.from_numpy(dropout_mask_numpy(seed, N * H * W * C)).view(N, H, W, C)
    assert torch.equal(f32.cpu(), base.cpu() * m)
    assert 0.45 < float((m > 0).float().mean()) < 0.55
    dout = _rand((N, H, W, C), 2).cuda()
    g1, _, _ = ops.norm_bwd(dout, y, sc, sh, mean, rstd, ops.ACT_RELU, want_f32=True, want_split=False, drop_p=0.5, drop_seed=seed)
    g2, _, _ = ops.norm_bwd(dout * m.cuda(), y, sc, sh, mean, rstd, ops.ACT_RELU, want_f32=True, want_split=False)
    assert (g1 - g2).abs().max().item() < 1e-6


def test_unet_training_applies_dropout_on_the_inner_blocks():
    """UnetGenerator(use_dropout=True) (reference networks.py:536, 604-605): Dropout(0.5) closes the num_downs-5 inner
    ngf*8 blocks in training.  The train engine draws a fresh mask per forward (outputs differ), restoring the torch seed
    restores the mask, the backward reuses the forward's mask (finite gradients for every parameter), and a generator
    without dropout has no drop levels."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine_train
    nd = 7                                                        # two dropout levels: 4 and 5
    sd = nets.make_state_dict(nets.unet_param_shapes(nd, 64, 3, 3, "batch"), 13, "stress")
    eng = engine_train.UnetTrainEngine(sd, num_downs=nd, norm="batch", norm_mode="batch", use_dropout=True)
    assert eng.drop_levels == {4, 5}
    assert engine_train.UnetTrainEngine(sd, num_downs=nd, norm="batch", norm_mode="batch").drop_levels == set()
    x = _rand((2, 3, 128, 128), 5).cuda()
    torch.manual_seed(11); y1, c1 = eng.forward_train(x)
    y2, _ = eng.forward_train(x)
    torch.manual_seed(11); y3, _ = eng.forward_train(x)
    assert (y1 - y2).abs().max().item() > 1e-4 and torch.equal(y1, y3)
    y0, _ = engine_train.UnetTrainEngine(sd, num_downs=nd, norm="batch", norm_mode="batch").forward_train(x)
    assert (y1 - y0).abs().max().item() > 1e-4
    g, _ = eng.backward(c1, _rand((2, 3, 128, 128), 6).cuda())
    expected = {k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    assert set(g) == expected and all(torch.isfinite(v).all() for v in g.values())
    # the module wrapper forwards the flag (define_G(use_dropout=True) in training)
    from deepliif_b200 import training
    from deepliif_b200.models import networks
    training.install()
    net = networks.define_G(3, 3, 64, "unet_128", "batch", True, gpu_ids=[0]).train()
    torch.manual_seed(3); a = net(x)
    b = net(x)
    assert (a - b).abs().max().item() > 1e-5
    net.eval()
    with torch.no_grad():
        assert torch.equal(net(x), net(x))


def test_resnet_training_with_dropout_runs_and_is_seeded():
    """use_dropout=True (the reference default, `not opt.no_dropout`): state_dict indices shift by the Dropout module;
    two forward passes differ unless the torch seed is restored."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from deepliif_b200 import engine_train
    cfg = dict(n_blocks=2, norm="batch", use_dropout=True, padding_type="zero")
    sd = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 2, "batch", True, "zero"), 3, "stress")
    eng = engine_train.ResnetTrainEngine(sd, norm_mode="batch", precision="bf16x3", **cfg)
    x = _rand((2, 3, 64, 64), 5).cuda()
    torch.manual_seed(7); y1, c1 = eng.forward_train(x)
    y2, _ = eng.forward_train(x)
    torch.manual_seed(7); y3, c3 = eng.forward_train(x)
    assert (y1 - y2).abs().max().item() > 1e-3 and torch.equal(y1, y3)
    g, _ = eng.backward(c1, _rand((2, 3, 64, 64), 6).cuda())
    expected = {k for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    assert set(g) == expected and all(torch.isfinite(v).all() for v in g.values())


def test_device_batches_prefetch_matches_host_transform(tmp_path):
    """uint8 batches -> pinned -> async H2D -> dlb_u8_to_f32 on a side stream == the reference's ToTensor + Normalize
    of the same tiles (oracle.pixel.transform), for every batch including the ragged last one."""
    import random
    import numpy as np
    from PIL import Image
    from oracle import pixel
    from oracle.gen_golden import dataset_opt, dataset_rows
    from deepliif_b200.data.aligned_dataset import AlignedDataset, DeviceBatches, collate_u8
    (tmp_path / "train").mkdir()
    for i, row in enumerate(dataset_rows(n=5)):
        Image.fromarray(row).save(tmp_path / "train" / f"s{i}.png")
    ds = AlignedDataset(dataset_opt(str(tmp_path), "resize_and_crop", 48, 32, False))
    random.seed(5)
    want = [ds[i][0].numpy() for i in range(5)]
    random.seed(5)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, num_workers=0, collate_fn=collate_u8)
    got = list(DeviceBatches(loader, torch.device("cuda", 0)))
    assert [b["A"].shape[0] for b in got] == [2, 2, 1] and all(len(b["B"]) == 5 for b in got)
    k = 0
    for b in got:
        for n in range(b["A"].shape[0]):
            planes = [b["A"][n]] + [t[n] for t in b["B"]]
            for j, pl in enumerate(planes):
                assert np.array_equal(pl.cpu().numpy()[None], pixel.transform(want[k][j]))
            k += 1


def test_padding_backward_kernels_match_autograd():
    """dlb_reflect_fold / dlb_stem_window_bwd against torch autograd of F.pad (fp32, pure sums: tight tolerance)."""
    from deepliif_b200 import ops
    from deepliif_b200.ops import PAD_REFLECT, PAD_ZERO
    for (N, H, W, C, p) in [(2, 9, 7, 8, 1), (1, 5, 6, 4, 3), (2, 16, 16, 64, 3), (1, 4, 4, 4, 3), (1, 6, 5, 4, 0)]:
        x = _rand((N, C, H, W), 1).requires_grad_(True)
        g = _rand((N, C, H + 2 * p, W + 2 * p), 2)
        add = _rand((N, H, W, C), 3)
        (F.pad(x, (p, p, p, p), mode="reflect") * g).sum().backward()
        got = ops.reflect_fold(g.permute(0, 2, 3, 1).contiguous().cuda(), p, add=add.cuda())
        want = x.grad.permute(0, 2, 3, 1) + add
        assert (got.cpu() - want).abs().max().item() < 1e-5, (N, H, W, C, p)
    for mode, tmode in ((PAD_ZERO, "constant"), (PAD_REFLECT, "reflect")):
        for (N, C, H, W) in [(2, 3, 12, 10), (1, 1, 8, 8), (1, 8, 5, 9)]:
            S, p = 7, 3
            x = _rand((N, C, H, W), 4).requires_grad_(True)
            dxw = _rand((N, H + 2 * p, W, 64), 5)
            xp = F.pad(x, (p, p, p, p), mode=tmode)
            # Xw[n, hp, w, s*8 + c] = xpad[n, c, hp, w + s]
            xw = torch.zeros((N, H + 2 * p, W, 64))
            for s_ in range(S):
                xw[..., s_ * 8:s_ * 8 + C] = xp[:, :, :, s_:s_ + W].permute(0, 2, 3, 1)
            (xw * dxw).sum().backward()
            got = ops.stem_window_bwd(dxw.cuda(), C, p, S, mode)
            assert (got.cpu() - x.grad).abs().max().item() < 1e-5, (mode, N, C, H, W)


def _tiny_model(tmp_path, no_dropout, seg_gen=False):
    from deepliif_b200 import training
    from deepliif_b200.cli import TRAIN_DEFAULTS
    from deepliif_b200.models import create_model
    p = dict(TRAIN_DEFAULTS, dataroot=str(tmp_path), checkpoints_dir=str(tmp_path), name="t", gpu_ids=(0,), modalities_no=2,
             seg_gen=seg_gen, norm="batch", no_dropout=no_dropout, padding="zero", net_g="resnet_2blocks", net_gs="unet_128",
             batch_size=2)
    opt = training.build_options(p)
    torch.manual_seed(0)
    model = create_model(opt)
    training.make_optimizers(model)
    model.train()
    return model


def test_cuda_graph_step_is_bit_identical_to_eager(tmp_path):
    """GraphedStep (capture once, replay per batch; Adam scalars and the batch read from device memory) must walk the
    same weights as the eager loop, bit for bit, including a ragged batch in the middle and an lr change."""
    from deepliif_b200 import training
    batches = [{"A": _rand((2, 3, 64, 64), 300 + i), "B": [_rand((2, 3, 64, 64), 400 + 10 * i + j) for j in range(2)],
                "A_paths": []} for i in range(7)]
    batches[4] = {"A": batches[4]["A"][:1], "B": [b[:1] for b in batches[4]["B"]], "A_paths": []}      # ragged
    flats = []
    for graphed in (False, True):
        model = _tiny_model(tmp_path, no_dropout=True)
        stepper = training.GraphedStep(model, warmup=2) if graphed else None
        for i, b in enumerate(batches):
            if i == 5:
                for o in (model.optimizer_G, model.optimizer_D):
                    o.param_groups[0]["lr"] = 1e-4
            if graphed:
                stepper(b)
            else:
                model.set_input(b); model.optimize_parameters()
        torch.cuda.synchronize()
        assert model.optimizer_G.t == len(batches) and model.optimizer_D.t == len(batches)
        flats.append((model.optimizer_G.flat.clone(), model.optimizer_D.flat.clone(), model.get_current_losses()))
    assert torch.equal(flats[0][0], flats[1][0]) and torch.equal(flats[0][1], flats[1][1])
    assert flats[0][2] == flats[1][2]


def test_cuda_graph_step_redraws_dropout_masks(tmp_path):
    """With dropout the host-drawn seeds are baked into the graph; the device step counter must still give every replay
    its own masks (same batch replayed twice -> different generator losses), forward and backward sharing them."""
    from deepliif_b200 import training
    model = _tiny_model(tmp_path, no_dropout=False, seg_gen=True)
    stepper = training.GraphedStep(model, warmup=1)
    b = {"A": _rand((2, 3, 128, 128), 500), "B": [_rand((2, 3, 128, 128), 501 + j) for j in range(3)], "A_paths": []}
    seen = []
    for _ in range(4):
        stepper(b)
        torch.cuda.synchronize()
        losses = model.get_current_losses()
        assert all(np.isfinite(v) for v in losses.values())
        seen.append(model.fake_B_1.detach().float().sum().item())
    assert stepper.graph is not None and len(set(seen[1:])) == 3


def test_optimize_parameters_matches_the_reference_model_golden(tmp_path):
    """tests/golden/train_step.npz holds the reference's own DeepLIIFModel.optimize_parameters() (CPU, seeded weights in
    all ten networks, seeded batch, VGG term patched out).  The same weights and batch through this package's model:
    all twelve losses — including the generator GAN terms, which are evaluated after the discriminator update — and
    the direction of the first Adam step on sampled tensors."""
    from deepliif_b200 import training
    from deepliif_b200.models import create_model
    from oracle.gen_golden import train_batch, train_params, train_state_dicts
    from test_train_step_golden_cpu import GOLD, reference_losses
    p = train_params(str(tmp_path)); p["gpu_ids"] = (0,)
    opt = training.build_options(p)
    torch.manual_seed(0)
    model = create_model(opt)
    sds = train_state_dicts()
    for name, sd in sds.items():
        model._net(name).module.load_state_dict(sd)
    training.make_optimizers(model)
    model.train()
    model.set_input(train_batch())
    model.optimize_parameters()
    torch.cuda.synchronize()
    got, ref = model.get_current_losses(), reference_losses()
    assert set(got) == set(ref)
    for k, v in ref.items():
        print(f"loss {k}: ours {got[k]:.6f} reference {v:.6f}")
        assert abs(got[k] - v) <= 2e-3 * max(1.0, abs(v)), k
    from deepliif_b200 import ops
    lr = {"G": opt.lr_g, "D": opt.lr_d}
    for name, key in (("G1", "model.1.weight"), ("GS0", "model.model.0.weight"), ("D1", "model.0.weight"), ("DS2", "model.0.weight")):
        w0 = sds[name][key]
        w_ref = torch.from_numpy(GOLD[f"{name}__{key}"])
        g_ref = torch.from_numpy(GOLD[f"{name}__{key}__grad"])
        w_our = model._net(name).module.state_dict()[key].cpu()
        # (1) optimizer parity, exact: the fused Adam kernel applied to the REFERENCE's gradient lands on the reference's
        #     post-step weights (torch.optim.Adam, lr 2e-4, betas (0.5, 0.999), eps 1e-8, step 1) to fp32 rounding
        p_ = w0.clone().cuda().contiguous().view(-1); m_ = torch.zeros_like(p_); v_ = torch.zeros_like(p_)
        ops.adam_step(p_, g_ref.cuda().contiguous().view(-1), m_, v_, lr[name[0]], opt.beta1, 0.999, 1e-8, 1)
        d_opt = (p_.cpu().view_as(w_ref) - w_ref).abs().max().item()
        # (2) post-step weights of the whole step against the reference.  Adam's first update is lr * g / (|g| + eps): where
        #     |g| is far above the gradient's rounding noise both implementations move by the same +-lr, so the weights agree
        #     to fp32 rounding; entries whose reference gradient is within the noise band may take the other sign.
        rms = g_ref.pow(2).mean().sqrt()
        d_w = (w_our - w_ref).abs()
        for thr in (0.05, 0.1, 0.25, 0.5):            # how fast the disagreements die out with the gradient magnitude
            m_ = g_ref.abs() >= thr * rms
            print(f"  {name}.{key}: |g| >= {thr} rms: {int((d_w[m_] > 2e-7).sum())} of {int(m_.sum())} post-step weights differ")
        solid = g_ref.abs() >= 0.25 * rms
        n_bad = int((d_w[solid] > 2e-7).sum())
        flips = float((d_w > 0.5 * lr[name[0]]).float().mean())
        uo, ur = w_our - w0, w_ref - w0
        cos = float((uo * ur).sum() / (uo.norm() * ur.norm()))
        print(f"{name}.{key}: adam(ref grad) vs ref weights max|d| {d_opt:.2e}; post-step weights: {int(solid.sum())} of "
              f"{solid.numel()} entries with |g| >= 0.25 rms -> {n_bad} differ by > 2e-7 (max {d_w[solid].max().item():.2e}); "
              f"sign flips overall {100 * flips:.2f}%; step cosine {cos:.4f}")
        assert d_opt <= 1e-8
        assert n_bad <= 0.002 * int(solid.sum())
        assert flips < 0.03 and cos > 0.95
    # BatchNorm2d running statistics after the step (what save_networks writes into the .pth files, base_model.py:190-212)
    for name, key in (("G1", "model.2"), ("G1", "model.5"), ("GS0", "model.model.1.model.2"), ("D1", "model.3"), ("DS2", "model.6")):
        sd_ = model._net(name).module.state_dict()
        for suffix, tol in (("running_mean", 2e-4), ("running_var", 2e-4)):
            ref_ = torch.from_numpy(GOLD[f"{name}__{key}.{suffix}"])
            got_ = sd_[f"{key}.{suffix}"].cpu()
            err = (got_ - ref_).abs().max().item()
            print(f"{name}.{key}.{suffix}: max|d| vs reference {err:.2e} (scale {ref_.abs().max().item():.3f})")
            assert err <= tol * max(1.0, ref_.abs().max().item())
        assert int(sd_[f"{key}.num_batches_tracked"]) == int(GOLD[f"{name}__{key}.num_batches_tracked"])

"""Small end-to-end exercise for compute-sanitizer: every kernel family once at tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nets
from deepliif_b200 import engine, engine_train, ops

def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed); return torch.rand(shape, generator=g) * 2 - 1

sd = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 1, "batch", False, "zero"), 3, "stress")
x = rnd((1, 3, 64, 64), 1).cuda()
eng = engine_train.ResnetTrainEngine(sd, n_blocks=1, norm="batch", use_dropout=False, padding_type="zero", norm_mode="batch")
y, ctx = eng.forward_train(x)
g, dx0 = eng.backward(ctx, rnd((1, 3, 64, 64), 2).cuda(), need_dx=True)
sdq = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 1, "batch", True, "reflect"), 8, "stress")
engq = engine_train.ResnetTrainEngine(sdq, n_blocks=1, norm="batch", use_dropout=True, padding_type="reflect", norm_mode="batch")
yq, cq = engq.forward_train(x)
gq, dxq = engq.backward(cq, rnd((1, 3, 64, 64), 9).cuda(), need_dx=True)
sdu = nets.make_state_dict(nets.unet_param_shapes(5, 64, 3, 3, "batch"), 4, "stress")
u = engine_train.UnetTrainEngine(sdu, num_downs=5, norm="batch", norm_mode="batch")
yu, cu = u.forward_train(rnd((1, 3, 32, 32), 3).cuda())
gu, dx = u.backward(cu, rnd((1, 3, 32, 32), 4).cuda())
sdd = nets.make_state_dict(nets.nlayer_d_param_shapes(3, 64, 6, "batch"), 5, "stress")
d = engine_train.NLayerDTrainEngine(sdd, n_layers=3, norm="batch", norm_mode="batch")
yd, cd = d.forward_train(rnd((1, 6, 64, 64), 5).cuda())
gd, dxd = d.backward(cd, rnd(tuple(yd.shape), 6).cuda())
sdr = nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 1, "instance", False, "reflect"), 6, "stress")
yr = engine.ResnetEngine(sdr, n_blocks=1, norm="instance", padding_type="reflect").forward(x)
# round 2: the fused-operand inference path — fused stem (patch staged in smem), halo-strip trunk convs with the skip add
# and the fp32 write-back (zero and reflect borders), and, forced on, the vertical-strip head and ConvTranspose phases
os.environ.update(DLB_FUSE_UP="1", DLB_FUSE_HEAD="1")
for pad_t, sd_ in (("reflect", sdr), ("zero", nets.make_state_dict(nets.resnet_param_shapes(3, 3, 64, 2, "instance", False, "zero"), 7, "stress"))):
    nb = 1 if pad_t == "reflect" else 2
    yf = engine.ResnetEngine(sd_, n_blocks=nb, norm="instance", padding_type=pad_t, fused=True, fuse_residual=True).forward(rnd((2, 3, 64, 96), 11).cuda())
    yn = engine.ResnetEngine(sd_, n_blocks=nb, norm="instance", padding_type=pad_t, fused=False).forward(rnd((2, 3, 64, 96), 11).cuda())
    print("fused vs unfused", pad_t, float((yf - yn).abs().max()))
# per-tap fused conversion (stride 2) and the dual-source halo strip (UNet up-convolution)
xr = rnd((1, 32, 32, 128), 12).cuda(); xr2 = rnd((1, 32, 32, 128), 13).cuda()
dd = ops.conv_desc(1, 32, 32, [128], 64, 4, 4, 2, 1, False, 0)
wh, wl = ops.pack_weights_tc(dd, (rnd((64, 128, 4, 4), 14) * 0.05).cuda())
ops.conv_tc_fused(dd, [dict(x=xr, act=ops.ACT_LRELU02)], wh, wl)
du = ops.conv_desc(1, 32, 32, [128, 128], 64, 4, 4, 2, 1, True, 0)
wh, wl = ops.pack_weights_tc(du, (rnd((256, 64, 4, 4), 15) * 0.05).cuda())
ops.conv_tc_fused(du, [dict(x=xr, act=ops.ACT_RELU), dict(x=xr2, act=ops.ACT_RELU)], wh, wl)
_, u8, mask = ops.seg_finish([yr], [1.0])
t = ops.u8_to_f32(u8)
# cell post-processing (union-find labelling, statistics, classification, boundaries)
from oracle import cells
from deepliif_b200 import postprocessing
orig, seg, marker = cells.synth_case(70, 93, 3)
ov, rf, sc = postprocessing.compute_final_results(orig, seg, marker, "40x", marker_thresh="default")
ov2, rf2, sc2 = postprocessing.compute_final_results(orig, seg, None, "40x", od_thresh_lower=20, od_thresh_upper=330, size_thresh=None)
torch.cuda.synchronize()
print("cells", sc, sc2)
print("sanitize run ok", float(y.abs().sum()), float(yu.abs().sum()), float(yd.abs().sum()), int(mask.sum()))

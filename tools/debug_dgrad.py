import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nets
from deepliif_b200 import engine_train, ops

def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed); return torch.rand(shape, generator=g) * 2 - 1

norm, n_layers = "instance", 3
sd = nets.make_state_dict(nets.nlayer_d_param_shapes(n_layers, 64, 6, norm), 8, "stress")
x = rnd((2, 6, 128, 128), 80).requires_grad_(True)
leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
taps = {}
y_ref = nets.nlayer_d_forward(x, leaf, n_layers=n_layers, norm=norm, norm_mode="batch", taps=taps)
for t in taps.values():
    t.retain_grad()
dY = rnd(tuple(y_ref.shape), 81)
(y_ref * dY).sum().backward()
eng = engine_train.NLayerDTrainEngine(sd, n_layers=n_layers, norm=norm, norm_mode="sample")
y, ctx = eng.forward_train(x.detach().cuda())
print("fwd err", (y.cpu() - y_ref.detach()).abs().max().item())
tape, N, (h, w) = ctx["tape"], ctx["N"], ctx["hw"]
dzh, dzl = ops.head_bwd_pack(dY.cuda().contiguous(), 1)
dout = eng.last64.dgrad(dzh, dzl, N, h, w, 1)
keys = list(taps.keys())   # conv keys in forward order: model.0, model.2, model.5, model.8, model.11
def cmp(name, ours_nhwc, ref_nchw):
    o = ours_nhwc.permute(0, 3, 1, 2).cpu()
    d = (o - ref_nchw).abs()
    print(f"{name}: max|d| {d.max().item():.3e} of scale {ref_nchw.abs().max().item():.3e}; shape {tuple(o.shape)}; "
          f"err by row {[round(v,4) for v in d.amax(dim=(0,1,3)).tolist()[:4]]}..{[round(v,4) for v in d.amax(dim=(0,1,3)).tolist()[-3:]]} "
          f"by col ..{[round(v,4) for v in d.amax(dim=(0,1,2)).tolist()[-3:]]}")
cmp("d a(model.8) [input grad of last conv]", dout, taps["model.8"].grad)
grads = {}
for i in range(len(tape) - 1, 0, -1):
    rec = tape[i]
    f32, hi, lo = ops.norm_bwd(dout, rec.y, rec.sc, rec.sh, rec.mean, rec.rstd, rec.act, pooled=eng._pooled(), want_f32=True, want_split=True)
    dw = rec.layer.wgrad(rec.x, hi, lo, *rec.dims, rec.pad)
    ref_w = leaf[rec.wkey + ".weight"].grad
    print(f"  wgrad {rec.wkey}: rel err {(dw.cpu() - ref_w).abs().max().item() / ref_w.abs().max().item():.3e}")
    # recompute wgrad from the fp32 dy through torch to separate norm_bwd from wgrad errors
    dout = rec.layer.dgrad(hi, lo, *rec.dims, rec.pad)
    prev_key = keys[i - 1]
    cmp(f"d a({prev_key})", dout, taps[prev_key].grad)

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import nets
from deepliif_b200 import engine_train, ops

def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed); return torch.rand(shape, generator=g) * 2 - 1
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()
norm, n_layers = "instance", 3
sd = nets.make_state_dict(nets.nlayer_d_param_shapes(n_layers, 64, 6, norm), 8, "stress")
x = rnd((2, 6, 128, 128), 80).requires_grad_(True)
leaf = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point else v) for k, v in sd.items()}
taps = {}
y_ref = nets.nlayer_d_forward(x, leaf, n_layers=n_layers, norm=norm, norm_mode="batch", taps=taps)
for t in taps.values(): t.retain_grad()
dY = rnd(tuple(y_ref.shape), 81)
(y_ref * dY).sum().backward()
eng = engine_train.NLayerDTrainEngine(sd, n_layers=n_layers, norm=norm, norm_mode="sample")
y, ctx = eng.forward_train(x.detach().cuda())
tape, N, (h, w) = ctx["tape"], ctx["N"], ctx["hw"]
dzh, dzl = ops.head_bwd_pack(dY.cuda().contiguous(), 1)
dout = eng.last64.dgrad(dzh, dzl, N, h, w, 1)
rec = tape[-1]
f32, hi, lo = ops.norm_bwd(dout, rec.y, rec.sc, rec.sh, rec.mean, rec.rstd, rec.act, pooled=False, want_f32=True, want_split=True)
torch.cuda.synchronize()
print("f32 vs hi+lo:", (f32 - (hi.float() + lo.float())).abs().max().item())
# reference dy from OUR y and dout through torch (CPU fp64)
yc = nchw(rec.y.cpu()).double().requires_grad_(True)
a = F.leaky_relu(F.instance_norm(yc, None, None, None, None, True, 0.0, 1e-5), 0.2)
a.backward(nchw(dout.cpu()).double())
print("norm_bwd in situ: max err", (nchw(f32.cpu()).double() - yc.grad).abs().max().item(), "scale", yc.grad.abs().max().item())
# the forward operand planes vs the oracle's activation
xin = nchw((rec.x.hi.float() + rec.x.lo.float()).cpu())
print("operand planes vs oracle a(model.5):", (xin - taps["model.5"].detach()).abs().max().item())
print("y8 vs oracle conv output: (via stats) skip")
# wgrad / dgrad through torch from our tensors
wt = leaf["model.8.weight"].detach().double().requires_grad_(True)
xi = xin.double().requires_grad_(True)
yy = F.conv2d(xi, wt, None, stride=1, padding=1)
yy.backward(nchw(f32.cpu()).double())
ref_w = leaf["model.8.weight"].grad
print("torch wgrad from our tensors vs oracle:", (wt.grad - ref_w).abs().max().item() / ref_w.abs().max().item())
dw = rec.layer.wgrad(rec.x, hi, lo, *rec.dims, rec.pad)
print("our wgrad vs torch-from-our-tensors:", (dw.cpu().double() - wt.grad).abs().max().item() / wt.grad.abs().max().item())
print("our wgrad vs oracle:", (dw.cpu() - ref_w).abs().max().item() / ref_w.abs().max().item())
dx = rec.layer.dgrad(hi, lo, *rec.dims, rec.pad)
print("our dgrad vs torch-from-our-tensors:", (nchw(dx.cpu()).double() - xi.grad).abs().max().item() / xi.grad.abs().max().item())
print("torch dgrad vs oracle d a(model.5):", (xi.grad - taps["model.5"].grad).abs().max().item() / taps["model.5"].grad.abs().max().item())

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
x = rnd((2, 6, 128, 128), 80)
taps = {}
with torch.no_grad():
    y_ref = nets.nlayer_d_forward(x, sd, n_layers=n_layers, norm=norm, norm_mode="batch", taps=taps)
    a5 = taps["model.5"]
    y8 = F.conv2d(a5, sd["model.8.weight"], sd["model.8.bias"], stride=1, padding=1)
eng = engine_train.NLayerDTrainEngine(sd, n_layers=n_layers, norm=norm, norm_mode="sample")
y, ctx = eng.forward_train(x.cuda())
rec = ctx["tape"][-1]
ours = nchw(rec.y.cpu())
d = (ours - y8).abs()
print("y8 ours vs oracle: max", d.max().item(), "scale", y8.abs().max().item())
print("per-row max err", [round(v, 5) for v in d.amax(dim=(0, 1, 3)).tolist()])
print("per-col max err", [round(v, 5) for v in d.amax(dim=(0, 1, 2)).tolist()])
print("per-n", d.amax(dim=(1, 2, 3)).tolist())
a8 = nchw(ops.norm_apply(rec.y, rec.sc, rec.sh, 2, want_f32=True, want_split=False)[0].cpu())
print("a8 ours vs oracle", (a8 - taps["model.8"]).abs().max().item())

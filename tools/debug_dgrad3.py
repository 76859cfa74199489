import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nets
from deepliif_b200 import engine_train, ops

def rnd(shape, seed):
    g = torch.Generator().manual_seed(seed); return torch.rand(shape, generator=g) * 2 - 1
norm, n_layers = "instance", 3
sd = nets.make_state_dict(nets.nlayer_d_param_shapes(n_layers, 64, 6, norm), 8, "stress")
x = rnd((2, 6, 128, 128), 80)
eng = engine_train.NLayerDTrainEngine(sd, n_layers=n_layers, norm=norm, norm_mode="sample")
y, ctx = eng.forward_train(x.cuda())
for rec in ctx["tape"][1:]:
    sc2, sh2, m2, r2 = ops.norm_stats(rec.y, None, None, False, want_stats=True)
    print(rec.wkey, tuple(rec.y.shape), "sc", (rec.sc - sc2).abs().max().item(), "sh", (rec.sh - sh2).abs().max().item(),
          "mean", (rec.mean - m2).abs().max().item(), "rstd", (rec.rstd - r2).abs().max().item(),
          "| mean mag", m2.abs().max().item(), "rstd mag", r2.abs().max().item())

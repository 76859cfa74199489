import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from deepliif_b200 import ops

def rnd(shape, seed, s=1.0):
    g = torch.Generator().manual_seed(seed); return (torch.rand(shape, generator=g) * 2 - 1) * s
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()
def split16(x):
    hi = x.to(torch.bfloat16); return hi, (x - hi.float()).to(torch.bfloat16)

for (N, H, W, C) in [(2, 15, 15, 512), (1, 15, 15, 512), (2, 16, 16, 512), (2, 15, 15, 64)]:
    y = (rnd((N, C, H, W), 1) * 2 + rnd((1, C, 1, 1), 2)).double().requires_grad_(True)
    dout = rnd((N, C, H, W), 5).double()
    a = F.leaky_relu(F.instance_norm(y, None, None, None, None, True, 0.0, 1e-5), 0.2)
    a.backward(dout)
    yd = nhwc(y.detach().float()).cuda()
    sc, sh, mean, rstd = ops.norm_stats(yd, None, None, False, want_stats=True)
    f32, hi, lo = ops.norm_bwd(nhwc(dout.float()).cuda(), yd, sc, sh, mean, rstd, 2, want_f32=True, want_split=True)
    err = (nchw(f32.cpu()).double() - y.grad).abs()
    print(f"norm_bwd {(N,H,W,C)}: max err {err.max().item():.3e} scale {y.grad.abs().max().item():.3e} by n {err.amax(dim=(1,2,3)).tolist()}")

for (N, H) in [(1, 16), (2, 16), (3, 16)]:
    Cin, Cout, R = 256, 512, 4
    x = rnd((N, Cin, H, H), 41).double().requires_grad_(True)
    w = rnd((Cout, Cin, R, R), 42, 0.05).double().requires_grad_(True)
    yv = F.conv2d(x, w, None, stride=1, padding=1)
    dy = rnd(tuple(yv.shape), 43).double()
    yv.backward(dy)
    d = ops.conv_desc(N, H, H, [Cin], Cout, R, R, 1, 1, False, 0)
    xh, xl = split16(nhwc(x.detach().float())); dh, dl = split16(nhwc(dy.float()))
    dw = ops.conv_wgrad(d, xh.cuda(), xl.cuda(), dh.cuda(), dl.cuda())
    print(f"wgrad N={N}: rel err {(dw.cpu().double() - w.grad).abs().max().item() / w.grad.abs().max().item():.3e}")
    oh = yv.shape[2]
    dd = ops.conv_desc(N, oh, oh, [Cout], Cin, R, R, 1, 1, True, 0)
    w_hi, w_lo = ops.pack_weights_tc(dd, w.detach().float().cuda(), ops.FMT_BF16, True)
    dx = ops.conv_tc(dd, [dh.cuda()], [dl.cuda()], w_hi, w_lo, None, ops.FMT_BF16, True)
    e = (nchw(dx.cpu()).double() - x.grad).abs()
    print(f"dgrad N={N}: rel err {e.max().item() / x.grad.abs().max().item():.3e} by n {e.amax(dim=(1,2,3)).tolist()}")

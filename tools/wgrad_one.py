"""One ResNet-block weight-gradient launch (N=8, 128x128, 256->256, 3x3) for an ncu capture."""
import sys, torch
sys.path.insert(0, "/root/repo")
from deepliif_b200 import ops
N, H, W, C = 8, 128, 128, 256
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda: (torch.randn((N, H, W, C), device="cuda", generator=g)).to(torch.bfloat16)
xh, xl, dh, dl = mk(), mk() * 0.01, mk(), mk() * 0.01
d = ops.conv_desc(N, H, W, [C], C, 3, 3, 1, 1, False, 0)
for _ in range(3):
    dw = ops.conv_wgrad(d, xh, xl, dh, dl)
torch.cuda.synchronize()
print("ok", float(dw.abs().mean()))

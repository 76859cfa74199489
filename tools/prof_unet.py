"""Per-op timing of one UNet-512 forward at micro-batch 8 (CUDA events around each op wrapper)."""
import sys, collections, torch
sys.path.insert(0, "/root/repo")
from deepliif_b200 import ops
from deepliif_b200.models import networks
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = networks.define_G(3, 3, 64, "unet_512", "batch", True, "normal", 0.02, []).to(dev).eval()
eng = net.engine()
x = torch.rand((8, 3, 512, 512), device=dev) * 2 - 1
for _ in range(3):
    y = eng.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = eng.forward(x)
e1.record(); torch.cuda.synchronize()
print(f"UNet-512 forward, batch 8: {e0.elapsed_time(e1) / 5:.3f} ms  ({48.44 * 8 * 3 / (e0.elapsed_time(e1) / 5):.0f} TF/s x3)")
recs = []
def wrap(name, fn, keyf):
    def f(*a, **k):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); r = fn(*a, **k); s1.record()
        recs.append((name, keyf(*a, **k), s0, s1))
        return r
    return f
dk = lambda d, *a, **k: (d.N, d.H, d.W, d.Cin[0] + (d.Cin[1] if d.nsrc > 1 else 0), d.Cout, d.R, d.stride, d.transposed)
ops.conv_tc = wrap("conv_tc", ops.conv_tc, dk)
ops.conv_direct = wrap("conv_direct", ops.conv_direct, dk)
ops.norm_apply = wrap("norm_apply", ops.norm_apply, lambda y, *a, **k: tuple(y.shape))
ops.norm_finalize = wrap("norm_finalize", ops.norm_finalize, lambda ws, N, HW, Cc, *a, **k: (N, HW, Cc))
ops.norm_stats = wrap("norm_stats", ops.norm_stats, lambda y, *a, **k: tuple(y.shape))
ops.head_finish = wrap("head_finish", ops.head_finish, lambda z, *a, **k: tuple(z.shape))
import deepliif_b200.engine as E
y = eng.forward(x); torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, key, s0, s1 in recs:
    agg.setdefault((name, key), [0.0, 0]); agg[(name, key)][0] += s0.elapsed_time(s1); agg[(name, key)][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"sum of op times {tot:.3f} ms")
for (name, key), (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f"{t:7.3f} ms {n:2d}x {name:14s} {key}")

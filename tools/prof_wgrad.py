"""Per-shape timing of conv_wgrad / conv_tc inside one training step (CUDA events around each call)."""
import sys, collections, torch
sys.path.insert(0, "/root/repo")
from deepliif_b200 import ops, training
from deepliif_b200.cli import TRAIN_DEFAULTS
from deepliif_b200.models import create_model
B, HW = 8, 512
dev = torch.device("cuda", 0)
p = dict(TRAIN_DEFAULTS, dataroot="/tmp", checkpoints_dir="/tmp/dlb_prof", name="p", gpu_ids=(0,), modalities_no=5, seg_gen=False,
         norm="instance", no_dropout=True, padding="zero", net_g="resnet_9blocks", net_d="basic", batch_size=B)
opt = training.build_options(p)
torch.manual_seed(0)
model = create_model(opt); training.make_optimizers(model); model.train()
g = torch.Generator().manual_seed(1)
batch = {"A": (torch.rand((B, 3, HW, HW), generator=g) * 2 - 1).to(dev),
         "B": [(torch.rand((B, 3, HW, HW), generator=g) * 2 - 1).to(dev) for _ in range(5)], "A_paths": []}
for _ in range(2):
    model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
recs = []
def wrap(name, fn):
    def f(d, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(d, *a, **k); e1.record()
        recs.append((name, (d.N, d.H, d.W, d.Cin[0], d.Cout, d.R, d.S, d.stride, d.pad, d.transposed), e0, e1))
        return r
    return f
ops.conv_wgrad = wrap("wgrad", ops.conv_wgrad)
ops.conv_tc = wrap("conv_tc", ops.conv_tc)
model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for name, key, e0, e1 in recs:
    agg[(name, key)][0] += e0.elapsed_time(e1); agg[(name, key)][1] += 1
for name in ("wgrad", "conv_tc"):
    tot = sum(v[0] for k, v in agg.items() if k[0] == name)
    print(f"== {name}: {tot:.1f} ms")
    for (nm, key), (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if nm != name: continue
        N, H, W, ci, co, R, S, st, pad, tr = key
        if tr: oh, ow = (H - 1) * st - 2 * pad + R + (1 if st == 2 else 0), (W - 1) * st - 2 * pad + S + (1 if st == 2 else 0); gf = 2 * N * H * W * ci * co * R * S / 1e9
        else: oh, ow = (H + 2 * pad - R) // st + 1, (W + 2 * pad - S) // st + 1; gf = 2 * N * oh * ow * ci * co * R * S / 1e9
        print(f"{t:8.2f} ms {n:4d}x  {t / n * 1e3:8.1f} us  {gf * 3 / (t / n):7.0f} TF/s(x3)  N{N} {H}x{W} {ci}->{co} {R}x{S} s{st} p{pad} {'T' if tr else ''}")

"""Kernel-time breakdown of one training step (BASELINE configs[3]) via torch.profiler (CUPTI sees every kernel)."""
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
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    model.set_input(batch); model.optimize_parameters()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
t_min, t_max = 1e30, 0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        import re
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", ev.name)
        name = (m.group(1) + (m.group(2) or ""))[:70] if m else ev.name[:60]
        tot[name][0] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        tot[name][1] += 1
        t_min = min(t_min, ev.time_range.start); t_max = max(t_max, ev.time_range.end)
busy = sum(v[0] for v in tot.values())
print(f"span {(t_max - t_min) / 1e3:.1f} ms, sum of kernel time {busy / 1e3:.1f} ms")
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{t / 1e3:9.2f} ms {n:6d}x {100 * t / busy:5.1f}%  {k}")

"""cProfile of the host side of one training step (where do the 240 ms of enqueue time go?)."""
import sys, cProfile, pstats, torch
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
for _ in range(3):
    model.set_input(batch); model.optimize_parameters()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(2):
    model.set_input(batch); model.optimize_parameters()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)

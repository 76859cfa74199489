"""One ResNet-9 generator forward at batch 8, 512x512 (the micro-batch of the headline step), issued eagerly on one stream:
the target of the per-launch ncu list (`ncu --metrics gpu__time_duration.sum ...`) kept under profiles/.

    python tools/one_forward.py [--fused 0|1] [--fuse-residual 0|1] [--reps 2]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--fuse-residual", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--net", default="resnet_9blocks")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--precision", default="bf16x3")
    a = ap.parse_args()
    from deepliif_b200.models import networks
    torch.manual_seed(0)
    g = networks.define_G(3, 3, 64, a.net, "batch", a.net.startswith("resnet"), "normal", 0.02, [], "zero")
    g.precision = a.precision
    g.fused, g.fuse_residual = bool(a.fused), bool(a.fuse_residual)
    os.environ["DLB_FUSED"] = str(a.fused)
    g.cuda().eval()
    x = (torch.rand((a.batch, 3, a.hw, a.hw), generator=torch.Generator().manual_seed(1)) * 2 - 1).cuda()
    eng = g.engine()
    for _ in range(a.reps):
        y = eng.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = eng.forward(x)
    e1.record()
    torch.cuda.synchronize()
    print("forward ms:", e0.elapsed_time(e1), "out", tuple(y.shape), float(y.abs().mean()))


if __name__ == "__main__":
    main()

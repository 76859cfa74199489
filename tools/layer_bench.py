"""Times the dedicated stem / head kernels alone at the headline micro-batch (8 x 512 x 512), CUDA events, L2 flushed
between repetitions; prints microseconds and the HBM rate of the algorithmic bytes.

    python tools/layer_bench.py [--which head|stem|both] [--batch 8] [--hw 512] [--reps 10]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps, flush):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="both")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--hw", type=int, default=512)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from deepliif_b200 import ops
    N, H = a.batch, a.hw
    g = torch.Generator().manual_seed(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if a.which in ("head", "both"):
        x = (torch.rand((N, H, H, 64), generator=g) * 2 - 1).cuda()
        w = ((torch.rand((3, 64, 7, 7), generator=g) * 2 - 1) * 0.04).cuda()
        sc = (torch.rand((N, 64), generator=g) + 0.5).cuda(); sh = (torch.rand((N, 64), generator=g) - 0.5).cuda()
        b = torch.zeros(3).cuda()
        wpk = ops.head_conv_pack(w)
        for mode in (ops.PAD_ZERO, ops.PAD_REFLECT):
            med, best = timed(lambda: ops.head_conv(x, sc, sh, ops.ACT_RELU, wpk, b, 3, mode), a.reps, flush)
            byt = x.numel() * 4 + N * 3 * H * H * 4
            print(f"head_conv pad_mode={mode}: median {med:.1f} us, best {best:.1f} us, {byt / med / 1e3:.0f} GB/s algorithmic")
    if a.which in ("stem", "both") and hasattr(ops, "stem_conv"):
        x = (torch.rand((N, 3, H, H), generator=g) * 2 - 1).cuda()
        w = ((torch.rand((64, 3, 7, 7), generator=g) * 2 - 1) * 0.1).cuda()
        wpk = ops.stem_conv_pack(w)
        for mode in (ops.PAD_ZERO, ops.PAD_REFLECT):
            ws = ops.stats_workspace(N, H * H, 64, x.device)
            med, best = timed(lambda: ops.stem_conv(x, wpk, None, 64, mode, stats_ws=ws), a.reps, flush)
            byt = x.numel() * 4 + N * 64 * H * H * 4
            print(f"stem_conv pad_mode={mode}: median {med:.1f} us, best {best:.1f} us, {byt / med / 1e3:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()

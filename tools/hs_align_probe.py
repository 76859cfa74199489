"""Does the tensor core fetch a halo-strip window slower when its 8-row groups are not 1024-byte aligned?

A 3x1 convolution in halo-strip mode has strip rows of exactly 8 pixels (group stride 1024 B, every tap window atom-
aligned); a 1x3 convolution has rows of 10 pixels (group stride 1280 B, windows start at 0 / 128 / 256 B).  Same channels,
same MMA count per tile (3 taps x 4 chunks x 12): the time per CTA tile isolates the operand-fetch effect."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepliif_b200 import ops  # noqa: E402


def run(R, S, N=8, H=128, W=128, C=256):
    g = torch.Generator().manual_seed(1)
    x = (torch.rand((N, H, W, C), generator=g) * 2 - 1).cuda()
    w = ((torch.rand((C, C, R, S), generator=g) * 2 - 1) * 0.05).cuda()
    d = ops.conv_desc(N, H, W, [C], C, R, S, 1, 1, False, 0)
    w_hi, w_lo = ops.pack_weights_tc(d, w, ops.FMT_BF16, True)
    src = [dict(x=x, act=ops.ACT_RELU)]
    oh, ow = ops.conv_out_shape(d)
    mode = ops.conv_tc_fused_mode(d, True, 0)
    for _ in range(3):
        ops.conv_tc_fused(d, src, w_hi, w_lo, None, ops.FMT_BF16, True, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_tc_fused(d, src, w_hi, w_lo, None, ops.FMT_BF16, True, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tiles = N * ((oh + 15) // 16) * ((ow + 7) // 8)
    print(f"{R}x{S}: mode {mode}, out {oh}x{ow}, {tiles} tiles, {ms * 1e3:.1f} us per launch, {ms * 1e6 / tiles * 148:.0f} ns per tile-slot")


if __name__ == "__main__":
    run(3, 1)
    run(1, 3)
    run(3, 3)

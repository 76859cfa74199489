"""compute-sanitizer target for the round-2 (second half) kernels only, at tiny sizes: the row-streaming stem and head
(zero / reflect borders, partial strips, CTA ranges spanning strips and images) and the conv_tc epilogue through shared
memory (coalesced copy-out on a plane-fed launch with fused statistics, TMA box stores on the fused-operand CTA-pair
launch, merged ConvTranspose phases).

    compute-sanitizer --tool memcheck python tools/sanitize_stream.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepliif_b200 import ops


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) * 2 - 1) * scale).cuda()


for (N, C, H, W, mode) in ((2, 3, 19, 130, ops.PAD_ZERO), (1, 4, 12, 40, ops.PAD_REFLECT)):
    x = rnd((N, C, H, W), 1)
    wpk = ops.stem_conv_pack(rnd((64, C, 7, 7), 2, 0.1))
    ws = ops.stats_workspace(N, H * W, 64, x.device)
    y = ops.stem_conv(x, wpk, rnd((64,), 3, 0.1), 64, mode, stats_ws=ws)
    sc, sh = ops.norm_finalize(ws, N, H * W, 64, None, None, False)
    wh = ops.head_conv_pack(rnd((3, 64, 7, 7), 4, 0.04))
    out = ops.head_conv(y, sc, sh, ops.ACT_RELU, wh, rnd((3,), 5, 0.1), 3, mode)
    print("stem/head", tuple(out.shape), float(out.abs().mean()))

# plane-fed conv with fused statistics (epilogue: transpose buffer + coalesced copy-out), stride 2 and merged ConvTranspose
x = rnd((2, 32, 32, 128), 6)
for (cout, k, st, pad, tr, op) in ((64, 3, 1, 1, False, 0), (128, 3, 2, 1, False, 0), (64, 3, 2, 1, True, 1)):
    d = ops.conv_desc(2, 32, 32, [128], cout, k, k, st, pad, tr, op)
    w = rnd((128, cout, k, k) if tr else (cout, 128, k, k), 7, 0.05)
    w_hi, w_lo = ops.pack_weights_tc(d, w, ops.FMT_BF16, True)
    xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
    oh, ow = ops.conv_out_shape(d)
    ws = ops.stats_workspace(2, oh * ow, cout, x.device)
    y = ops.conv_tc(d, [xh], [xl], w_hi, w_lo, None, ops.FMT_BF16, True, 0, stats_ws=ws)
    s1, _ = ops.norm_finalize(ws, 2, oh * ow, cout, None, None, False)
    print("conv_tc", cout, st, tr, float(y.abs().mean()), float(s1.abs().mean()))

# fused-operand CTA-pair launch (epilogue: TMA box stores), with the residual write-back
raw = rnd((2, 32, 32, 256), 8)
d = ops.conv_desc(2, 32, 32, [256], 256, 3, 3, 1, 1, False, 0)
w_hi, w_lo = ops.pack_weights_tc(d, rnd((256, 256, 3, 3), 9, 0.03), ops.FMT_BF16, True)
ws = ops.stats_workspace(2, 32 * 32, 256, raw.device)
keep = torch.empty_like(raw)
y = ops.conv_tc_fused(d, [dict(x=raw, scale=rnd((2, 256), 10) * 0.3 + 1, shift=rnd((2, 256), 11, 0.3), act=ops.ACT_RELU, out=keep)],
                      w_hi, w_lo, None, ops.FMT_BF16, True, 0, stats_ws=ws)
torch.cuda.synchronize()
print("conv_tc_fused", float(y.abs().mean()), float(keep.abs().mean()))
print("sanitize_stream ok")

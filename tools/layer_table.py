"""Per-layer table of one ResNet-9 generator forward (micro-batch 8, 512 x 512) from an ncu launch list: duration, algorithmic
FLOPs and HBM bytes, the rates they imply and the fraction of the measured peaks (MEASURED_PEAKS.json: sustained bf16 matmul
rate, copy bandwidth).  bf16x3 executes 3 MMAs per algorithmic MAC, so a conv's ceiling on the FLOP axis is 1/3.

    python tools/layer_table.py profiles/r02_launches_raw.csv [--first 31] > profiles/r02_layer_table.md
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, HW = 8, 512


def main():
    path = sys.argv[1]
    first = int(sys.argv[sys.argv.index("--first") + 1]) if "--first" in sys.argv else 31
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit"))
    rows = []
    for row in r:
        v = float(row[vi].replace(",", ""))
        v = v / 1e3 if row[ui] == "ns" else (v * 1e3 if row[ui] == "ms" else v)
        rows.append((row[ki].split("(")[0].split("::")[-1], v))
    rows = rows[first:]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    pf, pb = peaks.get("bf16_tflops_sustained", 1431.0), peaks.get("hbm_gbs", 6572.2)
    px = lambda s: N * (HW // s) ** 2
    conv = lambda s_out, cin, cout, taps: 2.0 * px(s_out) * cin * cout * taps / 1e9       # GFLOP
    f32 = lambda s, c: px(s) * c * 4 / 1e6                                               # MB
    # the forward's launches in order: (label, kernel-name prefix, count, GFLOP, MB)
    plan = [("stem 7x7 3->64 (+stats)", "stem_conv", 1, conv(1, 3, 64, 49), f32(1, 3) + f32(1, 64)),
            ("norm+ReLU+split of the stem output", "norm_apply", 1, 0, 2 * f32(1, 64)),
            ("down1 3x3 s2 64->128", "conv_tc", 1, conv(2, 64, 128, 9), f32(1, 64) + f32(2, 128)),
            ("norm+ReLU+split", "norm_apply", 1, 0, 2 * f32(2, 128)),
            ("down2 3x3 s2 128->256", "conv_tc", 1, conv(4, 128, 256, 9), f32(2, 128) + f32(4, 256)),
            ("18 ResNet-block convs 3x3 256->256 (fused operand)", "conv_tc", 18, 18 * conv(4, 256, 256, 9), 18 * 2 * f32(4, 256) + 9 * 2 * f32(4, 256)),
            ("norm+skip add+split of the trunk output", "norm_apply", 1, 0, 2 * f32(4, 256)),
            ("up1 ConvT 3x3 s2 256->128 (4 phase launches)", "conv_tc", 4, conv(2, 256, 128, 9 / 4), f32(4, 256) + f32(2, 128)),
            ("norm+ReLU+split", "norm_apply", 1, 0, 2 * f32(2, 128)),
            ("up2 ConvT 3x3 s2 128->64 (merged phases)", "conv_tc", 1, conv(1, 128, 64, 9 / 4), f32(2, 128) + f32(1, 64)),
            ("head norm+ReLU+7x7 64->3+tanh", "head_conv", 1, conv(1, 64, 3, 49), f32(1, 64) + f32(1, 3))]
    it = iter([x for x in rows if not x[0].startswith(("stats_", "array", "MeanOps", "direct_copy", "pack_w", "head_pack", "stem_pack"))])
    fin = sum(v for n, v in rows[:rows.index(next(x for x in rows if x[0].startswith("head_conv"))) + 1] if n.startswith("stats_"))
    print("| layer (8 tiles of 512 x 512) | us | GFLOP | MB | TFLOP/s | of bf16 peak | GB/s | of HBM peak |")
    print("|---|---|---|---|---|---|---|---|")
    total = 0.0
    for label, pref, cnt, gf, mb in plan:
        t = 0.0
        for _ in range(cnt):
            n, v = next(it)
            assert n.startswith(pref), (label, n)
            t += v
        total += t
        tf = gf / t * 1e3 if t else 0                 # GFLOP / us = PFLOP/s -> TFLOP/s
        gbs = mb / t * 1e3                             # MB / us = TB/s -> GB/s
        print(f"| {label} | {t:.0f} | {gf:.1f} | {mb:.0f} | {tf:.0f} | {tf / pf:.3f} | {gbs:.0f} | {gbs / pb:.2f} |")
    print(f"| 38 x stats_finalize | {fin:.0f} | | | | | | |")
    print(f"| **forward** | **{total + fin:.0f}** | | | | | | |")
    print()
    print(f"Peaks: {pf:.0f} TFLOP/s sustained bf16 matmul, {pb:.0f} GB/s copy (MEASURED_PEAKS.json).  Durations are ncu "
          "`gpu__time_duration.sum` per launch (serialised, cold caches, kernels run one at a time at ~1.6 GHz while the sustained "
          "matmul peak is measured under the power cap of a long run: the trunk's 0.34 is 0.29 of the burst peak "
          f"{peaks.get('bf16_tflops', 1698.8):.0f} TFLOP/s; bf16x3 ceiling 1/3); the trunk row counts each block's skip tensor once.")


if __name__ == "__main__":
    main()

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per launch (in order) and per kernel name.

    python tools/launch_shares.py gpurun_out/launches.csv [--skip N] [--take M] [--per-launch]
"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    take = int(sys.argv[sys.argv.index("--take") + 1]) if "--take" in sys.argv else 10 ** 9
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, gi, vi, ui, bi = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Metric Value", "Metric Unit", "Block Size"))
    for row in r:
        v = float(row[vi].replace(",", ""))
        if row[ui] == "ns":
            v /= 1e3
        elif row[ui] == "ms":
            v *= 1e3
        name = row[ki].split("(")[0].split("::")[-1]
        rows.append((name, row[gi], row[bi], v))
    rows = rows[skip:skip + take]
    if "--per-launch" in sys.argv:
        for i, (n, g, b, v) in enumerate(rows):
            print(f"{i + skip},{n[:40]},{g},{b},{v:.1f}")
        return
    agg = OrderedDict()
    for n, g, b, v in rows:
        k = (n, g, b)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel,grid,block,launches,total_us,avg_us,share_pct")
    for (n, g, b), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:48]},{g},{b},{c},{t:.1f},{t / c:.1f},{100 * t / tot:.1f}")
    print(f"# total {tot:.1f} us over {len(rows)} launches")


if __name__ == "__main__":
    main()

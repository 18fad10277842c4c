#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        name = row["Kernel Name"]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(c for c, _ in agg.values())} launches, {tot/1000:.3f} ms total device time")
    print(f"{'share':>7} {'count':>7} {'avg_us':>10}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t/tot*100:6.2f}% {c:7d} {t/c:10.2f}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)

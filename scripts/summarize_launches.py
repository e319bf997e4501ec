#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the captured window)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = name.replace("void ", "").replace("vck::<unnamed>::", "")[:64]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-3)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    print(f"{'ms':>10} {'share':>6} {'n':>6}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{v[1]:10.3f} {100 * v[1] / tot:5.1f}% {v[0]:6d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])

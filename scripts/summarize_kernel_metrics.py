#!/usr/bin/env python
"""Per-kernel table from an `ncu --metrics <list> --csv` launch list (one CSV row per launch and metric): launches, device time,
DRAM bytes -> achieved GB/s and fraction of the measured copy bandwidth (MEASURED_PEAKS.json), tensor-pipe activity, occupancy.

    python scripts/summarize_kernel_metrics.py gpurun_out/kernels.csv [peak_GBs] [--split-by-read] > profiles/rNN_per_kernel_ncu.txt

Times under ncu are cold-cache and serialised; bytes and pipe activity per launch are what the table is for."""
import collections
import csv
import json
import os
import re
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "ms": 1e3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("vck::<unnamed>::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
    if name.startswith("at::"):
        return ("torch " + re.sub(r"\(.*", "", name))[:48]
    return re.sub(r"\(.*", "", name)[:60]


def main(path, peak_gbs, split_by_read=False):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = row["ID"]
        d = launches.setdefault(key, {"name": short(row["Kernel Name"]), "grid": row.get("Grid Size", ""), "block": row.get("Block Size", "")})
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        m, u = row["Metric Name"], row["Metric Unit"]
        if m.startswith("gpu__time_duration"):
            v *= UNIT.get(u, 1e-3)
        elif m.startswith("dram__bytes"):
            v *= BYTES.get(u, 1.0)
        d[m] = v
    agg = collections.OrderedDict()
    for d in launches.values():
        if split_by_read:  # kernels that serve several problem sizes (the decode GEMV): one line per amount of DRAM read
            d["name"] += f" [{d.get('dram__bytes_read.sum', 0.0) / 1e6:.0f} MB read]"
        a = agg.setdefault(d["name"], collections.defaultdict(float))
        a["n"] += 1
        for k, v in d.items():
            if isinstance(v, float):
                a[k] += v
                a["w:" + k] += v * d.get("gpu__time_duration.sum", 0.0)
        a["grid"] = d["grid"]
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print(f"{len(launches)} launches, {tot / 1e3:.3f} ms serialised under ncu; HBM peak {peak_gbs:.0f} GB/s (measured copy bandwidth)")
    hdr = f"{'kernel':60} {'n':>4} {'ms':>7} {'share':>6} {'us/launch':>9} {'MB rd':>8} {'MB wr':>8} {'GB/s':>7} {'of HBM':>6} {'tc%':>5} {'mma%':>5} {'warps%':>6} {'issue%':>6} {'regs':>4}"
    print(hdr)

    def wavg(a, key):
        t = a["gpu__time_duration.sum"]
        return a["w:" + key] / t if t > 0 and ("w:" + key) in a else float("nan")

    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = a["gpu__time_duration.sum"]
        rd, wr = a.get("dram__bytes_read.sum", 0.0), a.get("dram__bytes_write.sum", 0.0)
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0.0
        regs = a.get("launch__registers_per_thread", 0.0) / a["n"]
        print(f"{name:60} {int(a['n']):4d} {t / 1e3:7.3f} {100 * t / tot:5.1f}% {t / a['n']:9.1f} {rd / a['n'] / 1e6:8.2f} {wr / a['n'] / 1e6:8.2f} "
              f"{gbs:7.0f} {gbs / peak_gbs:6.2f} {wavg(a, 'sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} "
              f"{wavg(a, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):5.1f} "
              f"{wavg(a, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{wavg(a, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} {regs:4.0f}")


if __name__ == "__main__":
    peak = None
    split = "--split-by-read" in sys.argv
    if split:
        sys.argv.remove("--split-by-read")
    if len(sys.argv) > 2:
        peak = float(sys.argv[2])
    else:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6555.0
    main(sys.argv[1], peak, split)

#!/usr/bin/env python
"""One block per kernel from `ncu -i X.ncu-rep --page raw --csv`: duration, DRAM traffic, pipe activity, occupancy, and the warp
stall reasons (cycles per issued instruction) -- the numbers quoted in DESIGN.md / profiles/README.md.

    ncu -i gpurun_out/r02ab_step_fwd.ncu-rep --page raw --csv > /tmp/fwd.csv && python scripts/ncu_full_summary.py /tmp/fwd.csv
"""
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 pipe active % (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "mma.sync tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), CTAs/SM"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in idx.items() if re.match(r"smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$", h)
                  or re.match(r"smsp__average_warp_latency_issue_stalled_.*\.ratio$", h)]
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("unnamed>::", "")
        print(f"== launch {r[idx['ID']]}: {name}")
        for k, label in KEYS:
            if k in idx and r[idx[k]] != "":
                print(f"   {label:48} {r[idx[k]]} {units[idx[k]]}")
        st = []
        for h, i in stall_cols:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            st.append((v, re.sub(r".*issue_stalled_(.*?)(_per_issue_active)?\.ratio", r"\1", h)))
        st.sort(reverse=True)
        if st:
            print("   stall reasons (warp cycles per issued instruction): " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python
"""Static SASS size of every GEMM kernel variant and of its epilogue chunk loop (the innermost backward branch that encloses the
tcgen05.ld = LDTM), from `cuobjdump -sass` of the built library -- runs on the build machine, no GPU.

    python scripts/sass_epilogue_count.py [videocad_b200/libvideocad_b200.so]

The epilogue warps of the 2-SM kernel are issue-bound on the heavy variants (profiles/r02ab_ncu_full_in_step_summary.txt): the loop
body executes once per (epilogue warp, 16-column chunk), i.e. body / 16 is an upper bound of the thread instructions per output
element (both sides of runtime branches are counted).  Used to check epilogue changes before spending GPU time on them."""
import os
import re
import subprocess
import sys


def main(so):
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    funcs, cur, ins = {}, None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                funcs[cur] = ins
            cur, ins = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if cur:
        funcs[cur] = ins
    rows = []
    for name, ins in funcs.items():
        m = re.search(r"(gemm_tc_pair_kernel|gemm_tc_kernel)IL[ij](\d+)EL[ij](\d+)E(?:Li(\d+)E)?", name)
        if not m:
            continue
        tag = f"{m.group(1)}<{m.group(2)}, {m.group(3)}{', ' + m.group(4) if m.group(4) else ''}>"
        ldtm = [a for a, t in ins if t.startswith("LDTM")]
        body = None
        for a, t in ins:
            b = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
            if b:
                tgt = int(b.group(1), 16)
                if tgt < a and any(tgt <= x <= a for x in ldtm):
                    n = (a - tgt) // 16
                    body = n if body is None or n < body else body
        rows.append((tag, len(ins), body))
    print(f"{'kernel variant':44} {'SASS total':>10} {'epilogue chunk loop':>20}")
    for tag, total, body in sorted(rows):
        print(f"{tag:44} {total:10d} {body if body is not None else '-':>20}")


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "videocad_b200", "libvideocad_b200.so"))

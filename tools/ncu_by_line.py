#!/usr/bin/env python
"""Attribute an ncu source-page capture to CUDA source lines.

    python tools/ncu_by_line.py report.ncu-rep object.o kernel_mangled_name [top_n] [ncu_kernel_name_substring]

ncu's CSV source page is per SASS instruction; `nvdisasm -g` of the same cubin carries the
line table.  Both list the kernel's instructions in address order, so they are joined by index.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    lines, cur, active = [], None, False
    for ln in dis:
        if ln.startswith("//--------------------- .text."):
            active = kernel in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", ln):
            lines.append(cur)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    blocks, cur_b = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur_b = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur_b)
        elif cur_b is not None and r and r[0] == "Address":
            cur_b["hdr"] = r
        elif cur_b is not None and cur_b["hdr"] and len(r) == len(cur_b["hdr"]):
            cur_b["rows"].append(r)
    want = sys.argv[5] if len(sys.argv) > 5 else kernel.split("EN")[0].lstrip("_Z0123456789N").replace("3xrs", "")
    def norm(name):  # ncu prints template arguments as "(int)1, (bool)0" and prefixes "void xrs::"
        return re.sub(r"\((?:int|bool|long|unsigned int)\)|\s+|xrs::|^void", "", name)

    match = [blk for blk in blocks if norm(want) in norm(blk["name"])]
    if not match:
        sys.exit(f"no kernel matching {want!r} in the report; kernels: {sorted({blk['name'] for blk in blocks})}")
    b = match[0]
    h = b["hdr"]
    i_ex = h.index("Instructions Executed")
    # the sampling column's name differs between ncu versions / report sections
    cands = [k for k, name in enumerate(h) if name.strip() in ("# Samples", "Samples", "Warp Stall Sampling (All Samples)")
             or name.strip().startswith("# Samples")]
    i_s = cands[0] if cands else None
    if len(b["rows"]) != len(lines):
        print(f"warning: {len(b['rows'])} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for k, r in enumerate(b["rows"]):
        key = lines[k] if k < len(lines) else None
        n, s = int(float(r[i_ex] or 0)), (int(float(r[i_s] or 0)) if i_s is not None else 0)
        agg[key][0] += n
        agg[key][1] += s
        tot_i += n
        tot_s += s
    print(b["name"], "| warp instructions", tot_i, "| stall samples", tot_s)
    for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {str(key):34s} inst {100.0 * n / max(tot_i, 1):6.2f}%   samples {100.0 * s / max(tot_s, 1):6.2f}%")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump: warp instructions executed per SASS opcode.

    python tools/ncu_source_hot.py report.ncu-rep [kernel-index]
"""
import csv
import subprocess
import sys
from collections import Counter


def opcode(src: str) -> str:
    toks = src.split()
    if not toks:
        return "?"
    tok = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    parts = tok.split(".")
    if parts[0] in ("MUFU", "LDG", "STG", "F2I", "I2F", "F2F") and len(parts) > 1:
        return ".".join(parts[:2])
    return parts[0]


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(txt.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    b = blocks[which]
    h = b["hdr"]
    i_src, i_ex, i_samp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, samples = Counter(), Counter()
    tot = 0
    for r in b["rows"]:
        n = int(float(r[i_ex] or 0))
        op = opcode(r[i_src])
        ops[op] += n
        samples[op] += int(float(r[i_samp] or 0))
        tot += n
    print(b["name"], "| SASS lines", len(b["rows"]), "| warp instructions", tot)
    for op, n in ops.most_common(30):
        print(f"  {op:14s} {n:14d} {100.0 * n / tot:6.2f}%   stall samples {samples[op]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""DRAM traffic per launch of the bench's kernels from an `ncu --set full` report, as the JSON that
bench.py reads for `roofline.traffic` (profiles/kernel_traffic.json).

    python tools/ncu_traffic.py two_step=<report.ncu-rep> [fused=<report.ncu-rep>] > profiles/kernel_traffic.json

Kernel names are reduced to the ones libxrs's own profiler uses (k2_gather_staged<bilinear>, ...).
Several launches of one kernel are averaged.
"""
import csv
import io
import json
import re
import subprocess
import sys

METHODS = {"0": "nearest", "1": "bilinear", "2": "triangular"}


def short_name(full: str) -> str:
    m = re.search(r"(k2_gather_staged|k2_gather_direct)<\w+, \(?(?:int\))?(\d)", full)
    if m:
        return f"{m.group(1)}<{METHODS.get(m.group(2), m.group(2))}>" if m.group(1) == "k2_gather_staged" else m.group(1)
    m = re.search(r"k2_gather_dual<\w+, \(?(?:int\))?(\d)", full)
    if m:
        return f"k2_gather_dual<nearest+{METHODS.get(m.group(1), m.group(1))}>"
    m = re.search(r"(k3_reproject)<\w+, \w+, \(?(?:int\))?(\d), \(?(?:bool\))?(\d)", full)
    if m:
        return f"k3_reproject{'_sep' if m.group(3) == '1' else ''}<{METHODS.get(m.group(2), m.group(2))}>"
    name = full.split("(")[0].split("<")[0].strip()
    return name.split("::")[-1].replace("void ", "")


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, metric, want_unit):
        v, u = float(r[col[metric]].replace(",", "")), units[col[metric]].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1,
                 "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
        return v * scale.get(u, 1)

    out = {}
    for r in data:
        name = short_name(r[col["Kernel Name"]])
        e = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += val(r, "dram__bytes_read.sum", "byte") + val(r, "dram__bytes_write.sum", "byte")
        e["us"] += val(r, "gpu__time_duration.sum", "us")
    return {k: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "duration_us_under_ncu": v["us"] / v["launches"],
                "launches": v["launches"]} for k, v in out.items()}


def main():
    out = {"source": "ncu --set full --clock-control none of `python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu "
                     "--no-configs --no-graph` on one B200 (tools/r2_pass.sh)"}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        out[key] = read(rep)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

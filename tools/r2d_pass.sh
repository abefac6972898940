#!/bin/bash
# Round 2, third session: end-to-end A/B of the two-method gather in the band-chunk pipeline on ONE box
# (PCIe rates differ from box to box by several per cent, so only same-box numbers compare).
#   gpurun --timeout 300 -- 'bash tools/r2d_pass.sh r2d'
set -u
TAG=${1:-r2d}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-configs"
timeout 100 $B --no-dual > "$OUT/bench_per_method.json" 2> "$OUT/bench_per_method.err"; echo "per-method rc=$?" | tee -a "$OUT/status.txt"
timeout 100 $B > "$OUT/bench_dual.json" 2> "$OUT/bench_dual.err"; echo "dual rc=$?" | tee -a "$OUT/status.txt"
XRS_PIPE_OUT_SLOTS=2 timeout 100 $B > "$OUT/bench_dual_2slots.json" 2> "$OUT/bench_dual_2slots.err"; echo "dual, 2 output slots rc=$?" | tee -a "$OUT/status.txt"
python - "$OUT" <<'PY' | tee -a "$OUT/status.txt"
import json, sys
for f in ("bench_per_method", "bench_dual", "bench_dual_2slots"):
    try:
        d = json.loads(open(f"{sys.argv[1]}/{f}.json").read().strip().splitlines()[-1])
        e = d["e2e"]
        print(f, "device ms/step", round(d["ms_per_step"], 3), "e2e ms/step", round(e["ms_per_step"], 2), "e2e value", round(e["value"]), d["parity_check"]["bilinear_bit_exact"])
    except Exception as ex:
        print(f, "unreadable:", ex)
PY

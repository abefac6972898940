#!/bin/bash
# Round 2, third session, last GPU call: the whole GPU suite and smoke() on the final tree, then the
# k1_resolve occupancy variants (XRS_K1R_MINBLOCKS = 5 / 6) against the product library on the same box.
#   gpurun --timeout 240 -- 'bash tools/r2f_pass.sh r2f'
set -u
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
tail -3 "$OUT/pytest_gpu.log"
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"
tail -2 "$OUT/smoke.log"
for v in main k1r5 k1r6; do
    if [ "$v" = main ]; then unset XRS_LIB; else export XRS_LIB=$PWD/xcube_resampling_b200/libxrs_$v.so; fi
    timeout 60 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-configs > "$OUT/bench_$v.json" 2> "$OUT/bench_$v.err"
    echo "== $v bench rc=$?" | tee -a "$OUT/status.txt"
    python - "$OUT/bench_$v.json" <<'PY' | tee -a "$OUT/status.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks = {k["kernel"]: k["ms_per_launch"] for k in d["roofline"]["kernels"]}
    print("   step ms", round(d["ms_per_step"], 4), "k1_resolve ms", round(ks.get("k1_resolve", 0), 4), "ij bit-exact", d["parity_check"]["ij_bit_exact"])
except Exception as e:
    print("   unreadable:", e)
PY
    if [ "$v" != main ]; then
        timeout 60 python -m pytest tests/test_rectify_gpu.py -x -q -m gpu > "$OUT/pytest_$v.log" 2>&1; echo "   pytest rectify ($v) rc=$?" | tee -a "$OUT/status.txt"
    fi
done
unset XRS_LIB
cat "$OUT/status.txt"

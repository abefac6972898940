#!/bin/bash
# Round 2, third session: ONE short gpurun call for the two-method gather (xrs_gather_ij2).  Ordered by
# importance, every result written as soon as it exists (the box time left was under ten minutes):
#   gpurun --timeout 480 -- 'bash tools/r2c_pass.sh r2c'
set -u
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
B="python bench.py"
timeout 200 python -m pytest tests/test_gather_dual_gpu.py -q -m gpu -x > "$OUT/pytest_dual.log" 2>&1; echo "pytest dual rc=$?" | tee -a "$OUT/status.txt"
tail -3 "$OUT/pytest_dual.log"
timeout 120 $B --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs > "$OUT/bench_dual_device.json" 2> "$OUT/bench_dual_device.err"; echo "bench dual (device leg) rc=$?" | tee -a "$OUT/status.txt"
timeout 300 python -m pytest tests/test_multigpu_gpu.py tests/test_rectify_gpu.py tests/test_spatial_gpu.py -q -m gpu -x > "$OUT/pytest_rectify.log" 2>&1; echo "pytest rectify+multigpu+spatial rc=$?" | tee -a "$OUT/status.txt"
tail -3 "$OUT/pytest_rectify.log"
timeout 300 $B --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench (full line) rc=$?" | tee -a "$OUT/status.txt"
SHORT="$B --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-graph"
timeout 240 ncu --clock-control none --set full --import-source on -k "regex:k2_gather_dual" -s 1 -c 1 -f -o "$OUT/dual" $SHORT > "$OUT/ncu_dual.log" 2>&1
if [ -f "$OUT/dual.ncu-rep" ]; then
    python tools/ncu_summary.py "$OUT/dual.ncu-rep" "$OUT/k2_dual_ncu_full.txt" > /dev/null 2>> "$OUT/ncu_dual.log"
    python tools/ncu_traffic.py dual="$OUT/dual.ncu-rep" > "$OUT/kernel_traffic_dual.json" 2>> "$OUT/ncu_dual.log"
    python tools/ncu_source_hot.py "$OUT/dual.ncu-rep" > "$OUT/k2_dual_opcodes.txt" 2>> "$OUT/ncu_dual.log"
    rm -f "$OUT/dual.ncu-rep"
fi
timeout 120 $B --steps 10 --warmup 3 --no-dual --no-e2e --no-cpu --no-configs > "$OUT/bench_per_method_device.json" 2> "$OUT/bench_per_method_device.err"; echo "bench per-method rc=$?" | tee -a "$OUT/status.txt"
timeout 200 ncu --clock-control none --metrics gpu__time_duration.sum -c 200 --csv --log-file "$OUT/bench_launches.csv" \
    $B --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > "$OUT/ncu_launches.log" 2>&1
timeout 400 python -m pytest tests/test_zz_full_size_gpu.py tests/test_reproject_gpu.py -q -m gpu -x > "$OUT/pytest_rest.log" 2>&1; echo "pytest full-size+reproject rc=$?" | tee -a "$OUT/status.txt"
python - "$OUT" <<'PY' | tee -a "$OUT/status.txt"
import json, sys
for f in ("bench_dual_device", "bench_per_method_device"):
    try:
        d = json.loads(open(f"{sys.argv[1]}/{f}.json").read().strip().splitlines()[-1])
        k = d["roofline"]["kernels"][0]
        print(f, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "top", k["kernel"], round(k["ms_per_launch"], 3), "frac", round(k["frac"], 3), "parity", d["parity_check"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
cat "$OUT/status.txt"

#!/bin/bash
# Round-2 measurement pass on one B200: GPU tests, smoke, the bench line (with configs and the CPU
# leg), the reference arm, the ncu launch list and --set full captures summarised on the box
# (the .ncu-rep files are too large to travel back).
#   gpurun --timeout 1500 -- 'bash tools/r2_pass.sh <tag>'
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
B="python bench.py"
NCU="ncu --clock-control none"

timeout 600 python -m pytest tests -q -m gpu --maxfail=20 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee -a "$OUT/status.txt"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/status.txt"
timeout 300 $B --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "ref rc=$?" | tee -a "$OUT/status.txt"
timeout 400 $B --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/status.txt"
timeout 200 $B --steps 10 --warmup 3 --fused --no-e2e --no-cpu --no-configs > "$OUT/bench_fused.json" 2> "$OUT/bench_fused.err"

timeout 100 python tools/microbench/scan_slab_bench.py > "$OUT/scan_slab_n8.json" 2> "$OUT/scan_slab_n8.err"

SHORT="$B --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs --no-graph"
timeout 200 $SHORT > "$OUT/short_plain.log" 2>&1 || { echo "short bench failed"; exit 1; }
# launch list (per-launch gpu__time_duration.sum; cold-cache, serialised: compare shares)
timeout 300 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$OUT/bench_launches.csv" \
    $B --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > "$OUT/ncu_launches.log" 2>&1
# --set full of the step's kernels (second step of the short run)
timeout 600 $NCU --set full --import-source on -k "regex:k0_|k1_|k2_" -s 9 -c 9 -f -o "$OUT/step" $SHORT > "$OUT/ncu_step.log" 2>&1
if [ -f "$OUT/step.ncu-rep" ]; then
    python tools/ncu_summary.py "$OUT/step.ncu-rep" "$OUT/step_ncu_full.txt" > /dev/null 2>> "$OUT/ncu_step.log"
    python tools/ncu_traffic.py two_step="$OUT/step.ncu-rep" > "$OUT/kernel_traffic.json" 2>> "$OUT/ncu_step.log"
    python tools/ncu_by_line.py "$OUT/step.ncu-rep" xcube_resampling_b200/_obj/rectify_ij.o _ZN3xrs10k1_scatterENS_6IjGeomE 40 "k1_scatter(" > "$OUT/k1_scatter_by_line.txt" 2>> "$OUT/ncu_step.log"
    python tools/ncu_by_line.py "$OUT/step.ncu-rep" xcube_resampling_b200/_obj/gather.o _ZN3xrs16k2_gather_stagedIfLi1ELb0EEEvNS_12StagedParamsIT_EEilllllNS_8IjSourceEllS2_ 30 "k2_gather_staged<float, 1, 0>" > "$OUT/k2_bilinear_by_line.txt" 2>> "$OUT/ncu_step.log"
    rm -f "$OUT/step.ncu-rep"
fi
ls -la "$OUT"
cat "$OUT/status.txt"
